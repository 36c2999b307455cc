// attention_tc.cu — spatial attention on the 5th-generation tensor cores (reference
// TimesformerSelfAttention.forward, models/modeling_timesformer_siglip.py:688-717: per frame and
// head, softmax(Q K^T / 8) V over the S <= 208 tokens of the frame; 196 at 224x224).
//
// One persistent CTA per SM walks (frame, head) items.  Per item the frame's Q, K and V head slices
// (S x 64 each) are pulled straight out of the fused QKV activation by TMA — a 4-D tensor map
// addresses token n of frame (b,t) at row (b*S + n)*T + t, so the residual stream's (b,n,t) order is
// read in place (the reference permutes and copies, :962-971) — into 128B-swizzled shared memory,
// double-buffered across items.  Roles:
//
//   warp 0        TMA producer  (Q as two 128-row M tiles, K and V as one SK-row tile; rows >= S are
//                                zero-filled by TMA's out-of-bounds handling)
//   warp 1        MMA issuer    S = Q K^T  : tcgen05.mma M=128, N=SK, K=64,  A,B from smem (K-major)
//                               O = P V    : tcgen05.mma M=128, N=64,  K=SK, A = P from TENSOR MEMORY,
//                                            B = V from smem as an MN-major operand (no transpose copy)
//   warps 4-7     softmax + output of M tile 0 (TMEM region 0), thread == query row
//   warps 8-11    softmax + output of M tile 1 (TMEM region 1)
//
// A TMEM region is 256 columns: S (fp32, SK columns) is overwritten in place by P (bf16/fp16 pairs,
// SK/2 columns) during the second softmax pass, and the O accumulator (64 columns at +128) reuses the
// then-dead upper half of S.  The two regions let the S/PV MMAs of one M tile run under the softmax
// of the other.  exp2 runs on MUFU in the log2 domain with the 1/8 scale folded in.
#include <cuda.h>
#include <math.h>

#include <type_traits>

#include "sf_kernels.h"
#include "sf_ptx.cuh"
#include "sf_tma.h"

namespace sf {
namespace {

constexpr int kHd = 64;
constexpr int kMaxKeys = 208;                       // keys per frame, padded to a multiple of 16
constexpr int kQTileBytes = 128 * 128;              // 128 query rows x 64 x 2 B
constexpr int kKVTileBytes = kMaxKeys * 128;        // 26 KB
constexpr int kStageBytes = 2 * kQTileBytes + 2 * kKVTileBytes;   // 84 KB per item
constexpr int kRegionCols = 256;
constexpr int kOCol = 128;                          // O accumulator inside a region
constexpr int kSmemBytes = 2 * kStageBytes + 256 + 1024;
constexpr float kLog2e = 1.4426950408889634f;

struct SpatialTcArgs {
  void* out;
  long out_ld;
  int heads, S, SK, T_inner, ntiles, items;
  float scale_log2;
};

template <typename T> struct Fmt;
template <> struct Fmt<__half> { static constexpr int value = 0; };
template <> struct Fmt<__nv_bfloat16> { static constexpr int value = 1; };

// 2^x for x <= 0 on the FMA / ALU pipes (no MUFU): x = n + f with n = round(x) taken from the low mantissa
// bits of x + 1.5 * 2^23, 2^f by a cubic on [-0.5, 0.5] (max relative error 1.9e-4, an order of magnitude
// below the rounding of P to a 16-bit operand), 2^n added into the exponent field.  x is clamped at -120
// (2^-120 is 0 for every purpose here) so the exponent arithmetic cannot wrap.
// Measured on B200 (round 2, cfg2 shapes, tools/exp_poly.sh): 0 of 4 -> 60.4 us, 1 of 4 -> 60.8, 2 of 4 -> 63.4,
// 3 of 4 -> 67.9: the exponential pass is NOT MUFU-bound (it waits on the tcgen05.ld / tcgen05.st round trips of
// its 16-column chunks), so the offload stays off; kept as a compile-time experiment (-DSF_EXP2_POLY_PER4=n).
#ifndef SF_EXP2_POLY_PER4
#define SF_EXP2_POLY_PER4 0
#endif
constexpr int kExp2PolyPer4 = SF_EXP2_POLY_PER4;     // of every 4 consecutive elements, how many use exp2_fma
__device__ __forceinline__ float exp2_fma(float x) {
  x = fmaxf(x, -120.0f);
  const float magic = 12582912.0f;                    // 1.5 * 2^23
  const float t = x + magic;
  const float f = x - (t - magic);
  float p = 0.05587553605437279f;
  p = fmaf(p, f, 0.24229462444782257f);
  p = fmaf(p, f, 0.6931272745132446f);
  p = fmaf(p, f, 0.999948263168335f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <typename T>
__global__ void __launch_bounds__(384, 1)
spatial_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                       const SpatialTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(smem + 2 * kStageBytes);
  uint64_t* kv_empty = kv_full + 2;
  uint64_t* s_full = kv_empty + 2;
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* region_free = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(region_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);        // the four softmax warps of the region
      mbar_init(&o_full[i], 1);
      mbar_init(&region_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * kRegionCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch_dependents();

  const int D = a.heads * kHd;
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      int i = 0;
      for (int item = blockIdx.x; item < a.items; item += gridDim.x, ++i) {
        const int st = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        const int frame = item / a.heads, h = item % a.heads;
        const int b = a.T_inner > 1 ? frame / a.T_inner : frame;
        const int t = a.T_inner > 1 ? frame % a.T_inner : 0;
        mbar_wait(&kv_empty[st], ph ^ 1);
        uint8_t* base = smem + st * kStageBytes;
        mbar_arrive_expect_tx(&kv_full[st], static_cast<uint32_t>(a.ntiles * kQTileBytes + 2 * a.SK * 128));
        tma_load_4d(base, &tmQ, &kv_full[st], h * kHd, t, 0, b);
        if (a.ntiles == 2) tma_load_4d(base + kQTileBytes, &tmQ, &kv_full[st], h * kHd, t, 128, b);
        tma_load_4d(base + 2 * kQTileBytes, &tmKV, &kv_full[st], D + h * kHd, t, 0, b);
        tma_load_4d(base + 2 * kQTileBytes + kKVTileBytes, &tmKV, &kv_full[st], 2 * D + h * kHd, t, 0, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one_sync()) {
      const uint32_t idesc_s = umma_idesc_f16(128, a.SK, Fmt<T>::value, 0);
      const uint32_t idesc_pv = umma_idesc_f16(128, kHd, Fmt<T>::value, 1);
      const int ksteps_pv = a.SK / 16;
      int i = 0;
      for (int item = blockIdx.x; item < a.items; item += gridDim.x, ++i) {
        const int st = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        const uint32_t par = i & 1;
        const uint32_t sbase = smem_u32(smem + st * kStageBytes);
        mbar_wait(&kv_full[st], ph);
        tc_fence_after();
        const uint64_t dk = umma_desc_sw128_kmajor(sbase + 2 * kQTileBytes);
        for (int r = 0; r < a.ntiles; ++r) {
          mbar_wait(&region_free[r], par ^ 1);
          tc_fence_after();
          const uint64_t dq = umma_desc_sw128_kmajor(sbase + r * kQTileBytes);
          const uint32_t d_s = tmem_base + r * kRegionCols;
#pragma unroll
          for (int k = 0; k < kHd / 16; ++k) umma_f16(d_s, dq + 2 * k, dk + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&s_full[r]);
        }
        const uint64_t dv = umma_desc_sw128_mnmajor(sbase + 2 * kQTileBytes + kKVTileBytes);
        for (int r = 0; r < a.ntiles; ++r) {
          mbar_wait(&p_full[r], par);
          tc_fence_after();
          const uint32_t d_o = tmem_base + r * kRegionCols + kOCol;
          const uint32_t a_p = tmem_base + r * kRegionCols;
          for (int k = 0; k < ksteps_pv; ++k)
            umma_f16_ts(d_o, a_p + k * 8, dv + static_cast<uint64_t>(k) * (2048 >> 4), idesc_pv, k > 0 ? 1u : 0u);
          umma_commit(&o_full[r]);
        }
        umma_commit(&kv_empty[st]);   // every MMA that reads this stage's smem has retired
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + output warps
    const int r = (warp - 4) >> 2;            // M tile / TMEM region
    const int quarter = warp & 3;
    if (r < a.ntiles) {
      const int n = r * 128 + quarter * 32 + lane;          // token (query row) of this thread
      const bool warp_valid = (r * 128 + quarter * 32) < a.S;
      const bool row_valid = n < a.S;
      const uint32_t t_s = tmem_base + r * kRegionCols + (static_cast<uint32_t>(quarter * 32) << 16);
      const int nfull = a.S >> 4;               // chunks of 16 keys without padding
      const int nchunks = a.SK >> 4;
      const bool wide = ((reinterpret_cast<uintptr_t>(a.out) & 31) == 0) && (a.out_ld % 16 == 0);
      T* outp = reinterpret_cast<T*>(a.out);
      int i = 0;
      for (int item = blockIdx.x; item < a.items; item += gridDim.x, ++i) {
        const uint32_t par = i & 1;
        const int frame = item / a.heads, h = item % a.heads;
        long row;
        if (a.T_inner > 1) row = (static_cast<long>(frame / a.T_inner) * a.S + n) * a.T_inner + frame % a.T_inner;
        else row = static_cast<long>(frame) * a.S + n;
        float l = 0.f;
        mbar_wait(&s_full[r], par);
        tc_fence_after();
        if (warp_valid) {
          // pass 1: row maximum of the raw scores
          float mx = -INFINITY;
          uint32_t v[2][16];
          tmem_ld_32x32b_x16(t_s, v[0]);
          for (int c = 0; c < nchunks; c += 2) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int cc = c + hh;
              if (cc < nchunks) {
                tmem_ld_wait();
                if (cc + 1 < nchunks) tmem_ld_32x32b_x16(t_s + (cc + 1) * 16, v[hh ^ 1]);
                if (cc < nfull) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(v[hh][j]));
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (cc * 16 + j < a.S) mx = fmaxf(mx, __uint_as_float(v[hh][j]));
                }
              }
            }
          }
          // pass 2: p = 2^((s - max) * scale * log2 e), row sum, P (16-bit pairs) over S in place
          const float nm = -mx * a.scale_log2;
          tmem_ld_32x32b_x16(t_s, v[0]);
          for (int c = 0; c < nchunks; c += 2) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int cc = c + hh;
              if (cc < nchunks) {
                tmem_ld_wait();
                if (cc + 1 < nchunks) tmem_ld_32x32b_x16(t_s + (cc + 1) * 16, v[hh ^ 1]);
                float pj[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float x = fmaf(__uint_as_float(v[hh][j]), a.scale_log2, nm);
                  // the exponential pass is bound by the MUFU pipe (16 ex2 per clock and SM): kExp2Poly of every
                  // 16 elements take the FMA pipe instead (exp2_fma below), the rest MUFU.EX2
                  if ((j % 4) < kExp2PolyPer4) {
                    pj[j] = exp2_fma(x);
                  } else {
                    float e;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
                    pj[j] = e;
                  }
                }
                if (cc >= nfull) {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (cc * 16 + j >= a.S) pj[j] = 0.f;
                }
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  l += pj[2 * j] + pj[2 * j + 1];
                  pk[j] = Pack2<T>::pack(pj[2 * j], pj[2 * j + 1]);
                }
                tmem_st_32x32b_x8(t_s + cc * 8, pk);
              }
            }
          }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[r]);

        // output: O / l for this thread's token, 64 contiguous elements of its row
        mbar_wait(&o_full[r], par);
        tc_fence_after();
        if (warp_valid) {
          const float inv = 1.0f / l;
          T* orow = outp + row * a.out_ld + h * kHd;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(t_s + kOCol + half * 32, o);
            tmem_ld_wait();
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              pk[j] = Pack2<T>::pack(__uint_as_float(o[2 * j]) * inv, __uint_as_float(o[2 * j + 1]) * inv);
            if (row_valid) {
              if (wide) {
#pragma unroll
                for (int q = 0; q < 2; ++q)
                  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + half * 32 + q * 16),
                               "r"(pk[q * 8]), "r"(pk[q * 8 + 1]), "r"(pk[q * 8 + 2]), "r"(pk[q * 8 + 3]),
                               "r"(pk[q * 8 + 4]), "r"(pk[q * 8 + 5]), "r"(pk[q * 8 + 6]), "r"(pk[q * 8 + 7])
                               : "memory");
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  *reinterpret_cast<uint4*>(orow + half * 32 + q * 8) =
                      make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&region_free[r]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kRegionCols);
  }
}

// 4-D view of the fused QKV activation: (column, t, n, b) with token n of frame (b,t) at row
// (b*S + n)*T + t; T = 1 describes contiguous frames.  Box = 64 columns x 1 x box_rows tokens x 1.
int make_frame_map(CUtensorMap* map, int dtype, const void* base, int ld, int ncols, int T, int S, int Bf,
                   int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -3;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(ncols), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(S),
                        static_cast<cuuint64_t>(Bf)};
  cuuint64_t gstride[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(T) * ld * 2,
                           static_cast<cuuint64_t>(S) * T * ld * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(kHd), 1, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapDataType dt = dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(map, dt, 4, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(frame map) failed (%d): ld=%d T=%d S=%d B=%d box=%d", (int)r, ld, T, S, Bf, box_rows);
    return -3;
  }
  return 0;
}

}  // namespace

bool spatial_attention_tc_supported(int ld_qkv, int S) {
  return S >= 1 && S <= kMaxKeys && (ld_qkv % 8 == 0);
}

int spatial_attention_tc(cudaStream_t stream, int dtype, const void* qkv, int ld_qkv, void* out, int ld_out,
                         int frames, int heads, int S, int T_inner, float scale) {
  const int T = T_inner > 1 ? T_inner : 1;
  const int Bf = frames / T;
  const int SK = (S + 15) & ~15;
  CUtensorMap tmQ, tmKV;
  int rc = make_frame_map(&tmQ, dtype, qkv, ld_qkv, 3 * heads * kHd, T, S, Bf, 128);
  if (rc) return rc;
  rc = make_frame_map(&tmKV, dtype, qkv, ld_qkv, 3 * heads * kHd, T, S, Bf, SK);
  if (rc) return rc;
  SpatialTcArgs a;
  a.out = out; a.out_ld = ld_out; a.heads = heads; a.S = S; a.SK = SK; a.T_inner = T_inner;
  a.ntiles = S > 128 ? 2 : 1;
  a.items = frames * heads;
  a.scale_log2 = scale * kLog2e;
  const int grid = a.items < num_sms() ? a.items : num_sms();
  cudaError_t e;
  {
    ProfScope ps(stream, kProfSpatialAttn, 4.0 * frames * heads * static_cast<double>(S) * S * kHd,
                 2.0 * frames * heads * kHd * 4.0 * S);
    LaunchCfg lc(dim3(static_cast<unsigned>(grid)), dim3(384), kSmemBytes, stream);
    if (dtype == kBF16) {
      static bool attr = false;
      if (!attr) {
        cudaFuncSetAttribute(spatial_attn_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        attr = true;
      }
      e = cudaLaunchKernelEx(&lc.cfg, spatial_attn_tc_kernel<__nv_bfloat16>, tmQ, tmKV, a);
    } else {
      static bool attr = false;
      if (!attr) {
        cudaFuncSetAttribute(spatial_attn_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        attr = true;
      }
      e = cudaLaunchKernelEx(&lc.cfg, spatial_attn_tc_kernel<__half>, tmQ, tmKV, a);
    }
  }
  count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("spatial_attention_tc launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

}  // namespace sf
