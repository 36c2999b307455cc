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
//
// This is the GENERAL kernel (any S <= 208).  224x224 frames (S = 196) take spatial_attn_row_kernel further
// down: whole score rows in registers, 43 us instead of 60 us at cfg2 shapes.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <stdio.h>

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
constexpr int kExp2PolyPer4 = SF_EXP2_POLY_PER4;
// Packing two non-negative fp32 probabilities into a bf16 pair.  F2FP (cvt.rn.bf16x2.f32) shares the quarter-rate XU
// pipe with MUFU.EX2; for bf16 the upper halves of the two words taken by one PRMT (ALU pipe) are the truncated values.
// With the exponent argument shifted by log2(1 + 2^-9) truncation errs by (-2^-9, +2^-9) relative, centred like
// round-to-nearest, and the row sum is taken from the same shifted fp32 values, so the normalisation is consistent.
#ifndef SF_PACK_PRMT
#define SF_PACK_PRMT 1
#endif
template <typename T> struct PPack {
  static constexpr float kShift = 0.f;
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) { return Pack2<T>::pack(lo, hi); }
};
#if SF_PACK_PRMT
template <> struct PPack<__nv_bfloat16> {
  static constexpr float kShift = 0.0028150156f;       // log2(1 + 2^-9)
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
  }
};
#endif     // of every 4 consecutive elements, how many use exp2_fma
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

// ---------------------------------------------------------------------------------------------------------------
// Round-2 kernel for 224x224 frames (193 <= S <= 200): every softmax thread holds its WHOLE score row in registers.
//
// What the round-1 kernel above is bound by (clock64 timelines, profiles/r2_spatial_attention_timeline.md): not the
// MUFU pipe, not tensor-memory bandwidth (a microbenchmark reads 400+ B/clk per SM), but the per-warp chain of
// 16-column tcgen05.ld -> compute -> tcgen05.st round trips, twice over the row, with two warps per scheduler.
// Spreading a row over 2-3 warps (tried: 58.8 / 66.7 us) only adds barriers.  Here, as in FlashAttention-4, a
// softmax warpgroup takes registers from the control warpgroup (setmaxnreg: 232 against 40), loads the 200 score
// columns of its row with four tcgen05.ld, and computes the maximum, the exponentials, the row sum and the packed P
// branch-free out of registers: one read of S, no cross-warp exchange, one dependent TMEM round trip per row.
//
//   warp 0 / 3  TMA producers: K | V into a ring of three stages (up to two items ahead), Q into one buffer per M
//               tile (free again once the tile's S has retired)
//   warp 1 / 2  MMA issuer of M tile 0 / 1: S -> (softmax) -> PV per item, each on its own barriers
//   warps 4-7   softmax + read-out of M tile 0 (TMEM region 0: S/P at columns [0, 208), O apart at [416, 480) so that
//               S of the next item only waits for PV, not for the read-out)
//   warps 8-11  the same for M tile 1 (region 1: S/P at [208, 416), O in the dead score columns [336, 400))
// The two tiles free-run; `skew` delays tile 1's first S so that its exponential phase falls into tile 0's
// load / maximum / read-out phases (the MUFU pipe is the one resource both need at full rate).
constexpr int kDefaultSkew = 2000;                              // cycles; sweep on B200 (three K|V stages): 0 -> 44.8 us, 2000 -> 43.2, 3500 -> 45.5, 5000 -> 46.1
constexpr int kRowOcts = 25;                                    // 8-column groups held per thread (200 columns)
constexpr int kReg1Col = kMaxKeys;                              // region 1 starts right behind region 0's scores
constexpr int kO0Col = 2 * kMaxKeys;                            // O of region 0
constexpr int kSoftmaxRegs = 232, kControlRegs = 40;   // 40 + 2 x 232 = 3 x 168: exactly the registers the CTA was launched with (more would block setmaxnreg.inc forever)
constexpr int kHalfKSteps = 7;                                  // PV k-steps (16 keys each) issued after the first half of P
constexpr int kRowKvStages = 3;                                 // K | V operand stages (52 KB each)
constexpr int kRowKvStageBytes = 2 * kKVTileBytes;
constexpr int kSmemBytesRow = kRowKvStages * kRowKvStageBytes + 2 * kQTileBytes + 1024 + 8 * 4096 + 1024;   // + one 32 x 128 B output tile per softmax warp

#ifdef SF_ATTN_TIMELINE
#define SF_TL(...) __VA_ARGS__
#else
#define SF_TL(...)
#endif

struct SpatialRowArgs {
  SpatialTcArgs a;
  int skew;
};

template <typename T>
__global__ void __launch_bounds__(384, 1)
spatial_attn_row_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const __grid_constant__ CUtensorMap tmO, const SpatialRowArgs ra) {
  const SpatialTcArgs& a = ra.a;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  // shared memory: kRowKvStages x (K | V) stages, one Q buffer per M tile, barriers, one output tile per softmax warp
  uint8_t* q_smem = smem + kRowKvStages * kRowKvStageBytes;
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(q_smem + 2 * kQTileBytes);
  uint64_t* kv_empty = kv_full + kRowKvStages;
  uint64_t* q_full = kv_empty + kRowKvStages;
  uint64_t* q_empty = q_full + 2;
  uint64_t* s_full = q_empty + 2;
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* region_free = o_full + 2;
  uint64_t* p_half = region_free + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_half + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < kRowKvStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);      // one commit per tile's issuer
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);        // the four softmax warps of the tile
      mbar_init(&p_half[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&region_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch_dependents();

  const int D = a.heads * kHd;
  const int nitems = (a.items - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1;
  if (warp < 4) {
    setmaxnreg_dec<kControlRegs>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer, K / V: a ring of three stages
      // runs up to two items ahead, so the two M tiles can stay half an item apart without waiting for operands
      if (elect_one_sync()) {
        int st = 0;
        uint32_t ph = 0;
        for (int i = 0; i < nitems; ++i) {
          const int item = blockIdx.x + i * gridDim.x;
          const int frame = item / a.heads, h = item % a.heads;
          const int b = a.T_inner > 1 ? frame / a.T_inner : frame;
          const int t = a.T_inner > 1 ? frame % a.T_inner : 0;
          mbar_wait(&kv_empty[st], ph ^ 1);
          uint8_t* base = smem + st * kRowKvStageBytes;
          mbar_arrive_expect_tx(&kv_full[st], static_cast<uint32_t>(2 * a.SK * 128));
          tma_load_4d(base, &tmKV, &kv_full[st], D + h * kHd, t, 0, b);
          tma_load_4d(base + kKVTileBytes, &tmKV, &kv_full[st], 2 * D + h * kHd, t, 0, b);
          if (++st == kRowKvStages) { st = 0; ph ^= 1; }
        }
      }
    } else if (warp == 3) {
      // ---------------------------------------------------------------- TMA producer, Q: one buffer per M tile, free
      // again as soon as the tile's S = Q K^T has retired
      if (elect_one_sync()) {
        for (int i = 0; i < nitems; ++i) {
          const int item = blockIdx.x + i * gridDim.x;
          const int frame = item / a.heads, h = item % a.heads;
          const int b = a.T_inner > 1 ? frame / a.T_inner : frame;
          const int t = a.T_inner > 1 ? frame % a.T_inner : 0;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            mbar_wait(&q_empty[r], (i & 1) ^ 1);
            mbar_arrive_expect_tx(&q_full[r], kQTileBytes);
            tma_load_4d(q_smem + r * kQTileBytes, &tmQ, &q_full[r], h * kHd, t, r * 128, b);
          }
        }
      }
    } else {
      // ---------------------------------------------------------------- MMA issuer of tile r
      const int r = warp - 1;
      if (elect_one_sync()) {
        const uint32_t idesc_s = umma_idesc_f16(128, a.SK, Fmt<T>::value, 0);
        const uint32_t idesc_pv = umma_idesc_f16(128, kHd, Fmt<T>::value, 1);
        const int ksteps_pv = a.SK / 16;
        const uint32_t d_s = tmem_base + (r ? kReg1Col : 0);
        const uint32_t d_o = tmem_base + (r ? kReg1Col + kOCol : kO0Col);
        int st = 0;
        uint32_t kv_ph = 0;
        for (int i = 0; i < nitems; ++i) {
          const uint32_t par = i & 1;
          const uint32_t sbase = smem_u32(smem + st * kRowKvStageBytes);
          mbar_wait(&kv_full[st], kv_ph);
          mbar_wait(&q_full[r], par);
          if (r == 1 && i == 0 && ra.skew > 0) {      // de-phase the two tiles once the first operands have landed
            const long long t0 = clock64();
            while (clock64() - t0 < ra.skew) {}
          }
          // tile 0: P of the previous item is dead once its PV has retired; tile 1: its O sits inside the score
          // columns, so the read-out has to be over as well
          if (r == 0) mbar_wait(&o_full[0], par ^ 1);
          else mbar_wait(&region_free[1], par ^ 1);
          tc_fence_after();
          const uint64_t dq = umma_desc_sw128_kmajor(smem_u32(q_smem + r * kQTileBytes));
          const uint64_t dk = umma_desc_sw128_kmajor(sbase);
#pragma unroll
          for (int k = 0; k < kHd / 16; ++k) umma_f16(d_s, dq + 2 * k, dk + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&s_full[r]);
          umma_commit(&q_empty[r]);     // the Q buffer is free for the next item
          const uint64_t dv = umma_desc_sw128_mnmajor(sbase + kKVTileBytes);
          mbar_wait(&p_half[r], par);                 // P of keys 0 .. 111 is in tensor memory
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kHalfKSteps; ++k)
            umma_f16_ts(d_o, d_s + k * 8, dv + static_cast<uint64_t>(k) * (2048 >> 4), idesc_pv, k > 0 ? 1u : 0u);
          mbar_wait(&p_full[r], par);
          tc_fence_after();
          for (int k = kHalfKSteps; k < ksteps_pv; ++k)
            umma_f16_ts(d_o, d_s + k * 8, dv + static_cast<uint64_t>(k) * (2048 >> 4), idesc_pv, 1u);
          umma_commit(&o_full[r]);
          umma_commit(&kv_empty[st]);   // this tile's MMAs that read the stage's smem have retired
          if (++st == kRowKvStages) { st = 0; kv_ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax + read-out of tile r, thread = row
    setmaxnreg_inc<kSoftmaxRegs>();
    const int r = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int n = r * 128 + quarter * 32 + lane;          // token (query row) of this thread
    const bool warp_valid = (r * 128 + quarter * 32) < a.S;
    const uint32_t t_s = tmem_base + (r ? kReg1Col : 0) + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_o = tmem_base + (r ? kReg1Col + kOCol : kO0Col) + (static_cast<uint32_t>(quarter * 32) << 16);
    SF_TL(__shared__ int tlog[2 * 12 * 7]; long long tbase = 0;)
    uint8_t* otile = q_smem + 2 * kQTileBytes + 1024 + (warp - 4) * 4096;   // 1024-byte aligned: TMA's 128B swizzle
    for (int i = 0; i < nitems; ++i) {
      const int item = blockIdx.x + i * gridDim.x;
      const uint32_t par = i & 1;
      float l = 0.f;
      SF_TL(long long tl0 = clock64(); long long tl1 = 0, tl2 = 0, tl3 = 0, tl4 = 0, tl5 = 0; if (i == 0) tbase = tl0;)
      mbar_wait(&s_full[r], par);
      SF_TL(tl1 = clock64();)
      tc_fence_after();
      if (warp_valid) {
        uint32_t v[kRowOcts * 8];
        tmem_ld_32x32b_x64(t_s, v);
        tmem_ld_32x32b_x64(t_s + 64, v + 64);
        tmem_ld_32x32b_x64(t_s + 128, v + 128);
        tmem_ld_32x32b_x8(t_s + 192, *reinterpret_cast<uint32_t (*)[8]>(v + 192));
        tmem_ld_wait();
        // keys S .. 199 (zero-filled K rows) out of the maximum and the sum
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (192 + j >= a.S) v[192 + j] = 0xff800000u;     // -inf
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int j = 0; j < kRowOcts * 8; j += 8) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
          m2 = fmaxf(m2, fmaxf(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])));
          m3 = fmaxf(m3, fmaxf(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])));
        }
        const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        SF_TL(tl2 = clock64();)
        const float nm = fmaf(-mx, a.scale_log2, PPack<T>::kShift);
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
        for (int o = 0; o < kRowOcts; ++o) {
          float pj[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = fmaf(__uint_as_float(v[o * 8 + j]), a.scale_log2, nm);
            if ((j % 4) < kExp2PolyPer4) {
              pj[j] = exp2_fma(x);
            } else {
              float e;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
              pj[j] = e;
            }
          }
          l0 += pj[0] + pj[1];
          l1 += pj[2] + pj[3];
          l2 += pj[4] + pj[5];
          l3 += pj[6] + pj[7];
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) pk[j] = PPack<T>::pack(pj[2 * j], pj[2 * j + 1]);
          tmem_st_32x32b_x4(t_s + o * 4, pk);
          if (o == 2 * kHalfKSteps - 1) {       // first half of P complete: the tensor pipe starts on PV under the rest
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_half[r]);
          }
        }
        if (a.SK > kRowOcts * 8) {         // keys 200 .. 207: V rows are zero, P must be finite
          const uint32_t z[4] = {0u, 0u, 0u, 0u};
          tmem_st_32x32b_x4(t_s + kRowOcts * 4, z);
        }
        l = (l0 + l1) + (l2 + l3);
        tmem_st_wait();
      }
      SF_TL(tl3 = clock64();)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!warp_valid) mbar_arrive(&p_half[r]);
        mbar_arrive(&p_full[r]);
      }

      // read-out: O / l, the 64 contiguous elements of this thread's token
      mbar_wait(&o_full[r], par);
      SF_TL(tl4 = clock64();)
      tc_fence_after();
      if (warp_valid) {
        uint32_t o[64];
        tmem_ld_32x32b_x64(t_o, o);
        const float inv = 1.0f / l;
        tmem_ld_wait();
        // O is in registers: hand the region back BEFORE the global stores (an mbarrier arrive is a release; behind the
        // stores it waited ~1.5 k cycles for them to drain)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&region_free[r]);
        SF_TL(tl5 = clock64();)
        // 32 rows x 128 B into this warp's shared-memory tile in TMA's 128-byte swizzle (16-byte chunk c of row lane at
        // chunk c ^ (lane & 7): conflict-free), then ONE bulk tensor store per warp: the 4-D map addresses token n of
        // frame (b,t) at row (b*S + n)*T + t and clips rows >= S.  (Per-thread stores cost ~2 k cycles of LSU time per
        // item, coalesced st.global through shared memory still ~1.5 k of store back-pressure in the softmax chain.)
        if (lane == 0) tma_store_wait_read<0>();       // the previous item's store has read the tile
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 w;
          w.x = Pack2<T>::pack(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
          w.y = Pack2<T>::pack(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
          w.z = Pack2<T>::pack(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
          w.w = Pack2<T>::pack(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
          *reinterpret_cast<uint4*>(otile + lane * 128 + ((q ^ (lane & 7)) << 4)) = w;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int frame = item / a.heads, h = item % a.heads;
          const int b = a.T_inner > 1 ? frame / a.T_inner : frame;
          const int t = a.T_inner > 1 ? frame % a.T_inner : 0;
          tma_store_4d(&tmO, otile, h * kHd, t, n, b);   // n of lane 0 = first token of this warp's 32 rows
          tma_store_commit();
        }
      }
      if (!warp_valid) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&region_free[r]);
      }
      SF_TL(if (blockIdx.x == 0 && lane == 0 && quarter == 0 && i < 12) { int* q = tlog + (r * 12 + i) * 7; q[0] = (int)(tl0 - tbase); q[1] = (int)(tl1 - tl0); q[2] = (int)(tl2 - tl0); q[3] = (int)(tl3 - tl0); q[4] = (int)(tl4 - tl0); q[5] = (int)(clock64() - tl0); q[6] = (int)(tl5 - tl0); })
    }
    if (lane == 0) tma_store_wait<0>();                // bulk stores out of this CTA's shared memory have completed
    SF_TL(if (blockIdx.x == 0 && lane == 0 && quarter == 0) for (int i = 0; i < nitems && i < 12; ++i) { int* q = tlog + (r * 12 + i) * 7; printf("tile %d item %2d: start %6d | s_full +%d max +%d exp+st +%d o_full +%d O loaded +%d out +%d\n", r, i, q[0], q[1], q[2], q[3], q[4], q[6], q[5]); })
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// 4-D view of the fused QKV activation: (column, t, n, b) with token n of frame (b,t) at row
// (b*S + n)*T + t; T = 1 describes contiguous frames.  Box = 64 columns x 1 x box_rows tokens x 1.
int make_frame_map(CUtensorMap* map, int dtype, const void* base, int ld, int ncols, int T, int S, int Bf,
                   int box_rows, bool store = false) {
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
                  CU_TENSOR_MAP_SWIZZLE_128B, store ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(frame map) failed (%d): ld=%d T=%d S=%d B=%d box=%d", (int)r, ld, T, S, Bf, box_rows);
    return -3;
  }
  return 0;
}

}  // namespace

int g_spatial_row_opt = -1;
void set_spatial_row(int on) { g_spatial_row_opt = on; }

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
  // 224x224 frames take the row-in-registers kernel (SF_SPATIAL_ROW=0 forces the general one); SF_SPATIAL_SKEW = start
  // offset of tile 1 in cycles.
  static const bool row_env = [] { const char* e = getenv("SF_SPATIAL_ROW"); return !(e && e[0] == '0'); }();
  const bool row_on = g_spatial_row_opt < 0 ? row_env : g_spatial_row_opt != 0;
  static const int skew = [] { const char* e = getenv("SF_SPATIAL_SKEW"); return e ? atoi(e) : kDefaultSkew; }();
  const bool row = row_on && S > (kRowOcts - 1) * 8 && S <= kRowOcts * 8 && ld_out % 8 == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  cudaError_t e;
  {
    ProfScope ps(stream, kProfSpatialAttn, 4.0 * frames * heads * static_cast<double>(S) * S * kHd,
                 2.0 * frames * heads * kHd * 4.0 * S);
    LaunchCfg lc(dim3(static_cast<unsigned>(grid)), dim3(384), row ? kSmemBytesRow : kSmemBytes, stream);
    static bool attr_set[4] = {false, false, false, false};
    auto set_attr = [&](auto kern, bool* attr) {
      if (!*attr) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, row ? kSmemBytesRow : kSmemBytes);
        *attr = true;
      }
    };
    if (row) {
      CUtensorMap tmO;
      rc = make_frame_map(&tmO, dtype, out, ld_out, heads * kHd, T, S, Bf, 32, true);
      if (rc) return rc;
      SpatialRowArgs ra;
      ra.a = a;
      ra.skew = skew;
      if (dtype == kBF16) {
        set_attr(spatial_attn_row_kernel<__nv_bfloat16>, &attr_set[0]);
        e = cudaLaunchKernelEx(&lc.cfg, spatial_attn_row_kernel<__nv_bfloat16>, tmQ, tmKV, tmO, ra);
      } else {
        set_attr(spatial_attn_row_kernel<__half>, &attr_set[1]);
        e = cudaLaunchKernelEx(&lc.cfg, spatial_attn_row_kernel<__half>, tmQ, tmKV, tmO, ra);
      }
    } else if (dtype == kBF16) {
      set_attr(spatial_attn_tc_kernel<__nv_bfloat16>, &attr_set[2]);
      e = cudaLaunchKernelEx(&lc.cfg, spatial_attn_tc_kernel<__nv_bfloat16>, tmQ, tmKV, a);
    } else {
      set_attr(spatial_attn_tc_kernel<__half>, &attr_set[3]);
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
