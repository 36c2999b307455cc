// gemm_tcgen05.cu — the workhorse of the StreamFormer encoder: every projection on the hot path
// (patch-embed conv-as-GEMM, temporal/spatial QKV, attention out-proj, temporal_dense, MLP fc1/fc2,
// pooling-head K/V/out/MLP; reference nn.Linear / nn.Conv2d call sites:
// models/modeling_timesformer_siglip.py:329-350, 513, 578, 629, 691, 728, 760, 811, 820, 830, 834,
// 895, 954, 1118-1124, 1135) runs through this one persistent, warp-specialised kernel:
//
//   warp 0        TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, mbarrier ring)
//   warp 1        MMA issuer     (tcgen05.mma kind::f16, fp32 accumulators in TMEM; with CG=2 one
//                                 256 x 256 x 16 instruction spans a CTA pair, cta_group::2)
//   warps 2..     epilogue       (tcgen05.ld -> bias / GELU / pos+time embed / gated residual ->
//                                 swizzled smem staging -> coalesced 16-byte global stores, optional
//                                 (b,t,n)<->(b,n,t) row permutation)
//
// TMEM holds two BN-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "sf_kernels.h"
#include "sf_ptx.cuh"

namespace sf {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;             // 64 x 2 B = one 128-byte swizzle row
constexpr int kUmmaK = 16;

// epilogue flavours (compile-time: keeps each kernel's code small and branch-free)
enum EpiMode : int { kEpiBias = 0, kEpiAct = 1, kEpiResidual = 2, kEpiEmbed = 3 };

struct GemmParams {
  int M, N, K;
  void* out;
  int ldo;
  GemmEpilogue epi;
};

// CG = CTAs cooperating on one tile (tcgen05 cta_group): 1 -> 128 x BN tile per CTA;
// 2 -> 256 x BN tile per CTA pair, each CTA staging its 128 A rows and BN/2 of the B rows.
// EW = epilogue warps (8 or 16).
template <int BN, int CG, int EW>
struct SmemLayout {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = (BN / CG) * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiBytesPerWarp = 2048 + 512;   // 32 rows x 64 B staging + bias slice (<=128 floats)
  static constexpr int kEpiBytes = EW * kEpiBytesPerWarp;
  static constexpr int kBudget = 232448 - 1024 - 256 - kEpiBytes;   // 227 KB minus slack, barriers, epilogue
  static constexpr int kStages = (kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes);
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kEpiOffset = kBarOffset + 256;
  static constexpr int kTotal = kEpiOffset + kEpiBytes + 1024;
  static_assert(kStages >= 3, "not enough shared memory for a 3-stage pipeline");
  static_assert(2 * kStages + 4 <= 32, "barrier area overflow");
};

template <typename T> struct UmmaFmt;
template <> struct UmmaFmt<__half> { static constexpr int value = 0; };
template <> struct UmmaFmt<__nv_bfloat16> { static constexpr int value = 1; };

// Exact-erf GELU (hidden_act="gelu", reference ACT2FN["gelu"], …siglip.py:814-817) with erf from
// Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16/fp16 output rounding):
//   erf(z) = sign(z) * (1 - (a1 t + a2 t^2 + a3 t^3 + a4 t^4 + a5 t^5) exp(-z^2)),  t = 1/(1 + p|z|)
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = x * 0.70710678118654752440f;
  const float az = fabsf(z);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, az, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(az * az * -1.4426950408889634f));
  const float erf_abs = fmaf(-poly, e, 1.0f);
  const float erf_z = copysignf(erf_abs, z);
  const float h = 0.5f * x;
  return fmaf(h, erf_z, h);
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * (x + k1 * x * x * x);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
  const float h = 0.5f * x;
  return fmaf(h, th, h);
}

// nearest-neighbour source index, identical arithmetic to F.interpolate(mode="nearest"):
// scale = float(in)/out ; src = min(floor(dst*scale), in-1)
__device__ __forceinline__ int time_index(int t_abs, int time_len, int time_total) {
  if (time_total <= time_len) return t_abs;
  float scale = static_cast<float>(time_len) / static_cast<float>(time_total);
  int s = static_cast<int>(floorf(static_cast<float>(t_abs) * scale));
  return s < time_len - 1 ? s : time_len - 1;
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 q;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(addr));
  return q;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 q;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(addr));
  return q;
}

template <typename T, int BN, int CG, int EPI, int EW>
__global__ void __launch_bounds__(64 + EW * 32, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  using L = SmemLayout<BN, CG, EW>;
  constexpr int kStages = L::kStages;
  constexpr int kTileM = kBM * CG;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;   // rank inside the CTA pair
  const bool leader = cta_rank == 0;
  const int worker = blockIdx.x / CG;          // one worker = one CTA (CG=1) or one CTA pair (CG=2)
  const int num_workers = gridDim.x / CG;

  const int m_tiles = (p.M + kTileM - 1) / kTileM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (p.K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], EW * CG);   // (leader's copy) epilogue warps of every CTA of the pair
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      tmem_alloc_cg2(tmem_slot, 2 * BN);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, 2 * BN);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (every CTA)
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        const int m0 = (tile / n_tiles) * kTileM + static_cast<int>(cta_rank) * kBM;
        const int n0 = (tile % n_tiles) * BN + static_cast<int>(cta_rank) * (BN / CG);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if constexpr (CG == 2) {
            // the leader's barrier counts the bytes of both CTAs' loads for this stage
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            tma_load_2d_cg2(sa, &tmA, &full_bar[stage], kb * kBK, m0);
            tma_load_2d_cg2(sb, &tmB, &full_bar[stage], kb * kBK, n0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * kBK, m0);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * kBK, n0);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_f16(kTileM, BN, UmmaFmt<T>::value);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
          const uint64_t da = umma_desc_sw128_kmajor(sa);
          const uint64_t db = umma_desc_sw128_kmajor(sb);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // advance 32 bytes (16 bf16) along K inside the swizzle atom: +2 in the >>4 address field
            if constexpr (CG == 2) umma_f16_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if constexpr (CG == 2) umma_commit_cg2_mc(&empty_bar[stage], 0x3); else umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if constexpr (CG == 2) umma_commit_cg2_mc(&tmem_full[acc], 0x3); else umma_commit(&tmem_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // Each warp owns 32 accumulator rows (its TMEM lane quarter) x kColsPerWarp columns and walks
    // them in 32-column chunks, software-pipelined (the tcgen05.ld of chunk c+1 is in flight while
    // chunk c is processed):  TMEM -> bias/act/embed/residual in the thread==row layout -> 2 KB
    // swizzled smem staging -> 16-byte global stores in an (8 rows x 64 B) coalesced layout.
    // Residual rows travel the same staging buffer in the opposite direction first.
    const int ew = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int colgrp = ew >> 2;            // which slice of the BN columns
    constexpr int kColsPerWarp = BN / (EW / 4);
    constexpr int kChunks = kColsPerWarp / 32;
    const uint32_t stage_u = smem_u32(smem + L::kEpiOffset + ew * L::kEpiBytesPerWarp);
    const uint32_t bias_u = stage_u + 2048;
    const GemmEpilogue& e = p.epi;
    const float gscale = e.gate ? tanhf(__ldg(e.gate)) : 1.0f;
    T* out = reinterpret_cast<T*>(p.out);
    const T* res = reinterpret_cast<const T*>(e.residual);
    const int crow0 = lane >> 2, cchunk = lane & 3;   // coalesced layout: rows crow0 + 8*i, 16-byte chunk cchunk
    auto stage_addr = [&](int row, int ch) -> uint32_t {
      return stage_u + static_cast<uint32_t>(row * 64 + ((ch ^ ((row >> 1) & 3)) << 4));
    };
    int it = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * kTileM + static_cast<int>(cta_rank) * kBM;
      const int n0 = (tile % n_tiles) * BN;
      const int m = m0 + quarter * 32 + lane;
      const bool row_ok = m < p.M;
      // row decomposition / permutation
      long r = m;
      int site = 0, frame = 0;
      if (e.row_map == kRowBTNtoBNT || EPI == kEpiEmbed) {
        // m = (b*T + t)*S + n
        site = m % e.S;
        const int bt = m / e.S;
        frame = bt % e.T;
        const int b = bt / e.T;
        if (e.row_map == kRowBTNtoBNT) r = (static_cast<long>(b) * e.S + site) * e.T + frame;
      } else if (e.row_map == kRowBNTtoBTN) {
        // m = (b*S + n)*T + t
        frame = m % e.T;
        const int bn = m / e.T;
        site = bn % e.S;
        const int b = bn / e.S;
        r = (static_cast<long>(b) * e.T + frame) * e.S + site;
      }
      const float* pos_row = nullptr;
      const float* time_row = nullptr;
      if constexpr (EPI == kEpiEmbed) {
        if (e.pos) pos_row = e.pos + static_cast<long>(site) * p.N;
        if (e.time_emb)
          time_row = e.time_emb + static_cast<long>(time_index(e.time_off + frame, e.time_len, e.time_total)) * p.N;
      }
      // output rows handled by this lane in the coalesced layout (-1 = out of range)
      const int r_own = row_ok ? static_cast<int>(r) : -1;
      int r_c[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) r_c[i] = __shfl_sync(0xffffffffu, r_own, crow0 + 8 * i);
      const int wcol0 = n0 + colgrp * kColsPerWarp;
      // this warp's bias slice -> smem (overlaps the MMAs of this tile)
      if (lane * 4 < kColsPerWarp) {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const int bc = wcol0 + lane * 4;
        if (e.bias && bc < p.N) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + bc));
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(bias_u + lane * 16), "f"(b4.x), "f"(b4.y),
                     "f"(b4.z), "f"(b4.w) : "memory");
      }
      __syncwarp();

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN +
                              colgrp * kColsPerWarp;
      uint32_t raw[2][32];
      tmem_ld_32x32b_x32(t_base, raw[0]);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int col0 = wcol0 + c * 32;
        const int ccol = col0 + cchunk * 8;
        const bool ccol_ok = ccol < p.N;
        uint4 rr[4];
        if constexpr (EPI == kEpiResidual) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            rr[i] = make_uint4(0u, 0u, 0u, 0u);
            if (r_c[i] >= 0 && ccol_ok)
              rr[i] = *reinterpret_cast<const uint4*>(res + static_cast<long>(r_c[i]) * e.ldr + ccol);
          }
        }
        float4 bb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bb[j] = lds128f(bias_u + (c * 32 + j * 4) * 4);
        tmem_ld_wait();                                     // chunk c is in registers
        if (c + 1 < kChunks) tmem_ld_32x32b_x32(t_base + (c + 1) * 32, raw[(c + 1) & 1]);
        if constexpr (EPI == kEpiResidual) {
#pragma unroll
          for (int i = 0; i < 4; ++i) sts128(stage_addr(crow0 + 8 * i, cchunk), rr[i].x, rr[i].y, rr[i].z, rr[i].w);
          __syncwarp();
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // 4 groups of 8 columns
          float vv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) vv[j] = __uint_as_float(raw[c & 1][g * 8 + j]);
          vv[0] += bb[2 * g].x; vv[1] += bb[2 * g].y; vv[2] += bb[2 * g].z; vv[3] += bb[2 * g].w;
          vv[4] += bb[2 * g + 1].x; vv[5] += bb[2 * g + 1].y; vv[6] += bb[2 * g + 1].z; vv[7] += bb[2 * g + 1].w;
          if constexpr (EPI == kEpiAct) {
            if (e.act == kActGeluErf) {
#pragma unroll
              for (int j = 0; j < 8; ++j) vv[j] = gelu_erf(vv[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) vv[j] = gelu_tanh(vv[j]);
            }
          }
          if constexpr (EPI == kEpiEmbed) {
            const int col = col0 + g * 8;
            if (pos_row && row_ok && col < p.N) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(pos_row + col));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(pos_row + col + 4));
              vv[0] += b0.x; vv[1] += b0.y; vv[2] += b0.z; vv[3] += b0.w;
              vv[4] += b1.x; vv[5] += b1.y; vv[6] += b1.z; vv[7] += b1.w;
            }
            if (time_row && row_ok && col < p.N) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(time_row + col));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(time_row + col + 4));
              vv[0] += b0.x; vv[1] += b0.y; vv[2] += b0.z; vv[3] += b0.w;
              vv[4] += b1.x; vv[5] += b1.y; vv[6] += b1.z; vv[7] += b1.w;
            }
          }
          const uint32_t my = stage_addr(lane, g);
          if constexpr (EPI == kEpiResidual) {
            const uint4 q = lds128(my);
            const float2 r0 = Pack2<T>::unpack(q.x), r1 = Pack2<T>::unpack(q.y);
            const float2 r2 = Pack2<T>::unpack(q.z), r3 = Pack2<T>::unpack(q.w);
            vv[0] = fmaf(gscale, vv[0], r0.x); vv[1] = fmaf(gscale, vv[1], r0.y);
            vv[2] = fmaf(gscale, vv[2], r1.x); vv[3] = fmaf(gscale, vv[3], r1.y);
            vv[4] = fmaf(gscale, vv[4], r2.x); vv[5] = fmaf(gscale, vv[5], r2.y);
            vv[6] = fmaf(gscale, vv[6], r3.x); vv[7] = fmaf(gscale, vv[7], r3.y);
          } else if constexpr (EPI == kEpiBias) {
            if (e.gate) {
#pragma unroll
              for (int j = 0; j < 8; ++j) vv[j] *= gscale;
            }
          }
          sts128(my, Pack2<T>::pack(vv[0], vv[1]), Pack2<T>::pack(vv[2], vv[3]), Pack2<T>::pack(vv[4], vv[5]),
                 Pack2<T>::pack(vv[6], vv[7]));
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 q = lds128(stage_addr(crow0 + 8 * i, cchunk));
          if (r_c[i] >= 0 && ccol_ok) *reinterpret_cast<uint4*>(out + static_cast<long>(r_c[i]) * p.ldo + ccol) = q;
        }
        __syncwarp();  // staging buffer is reused by the next chunk
      }
      // all TMEM reads of this accumulator are complete -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_remote(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
      }
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // peers may still signal our barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, 2 * BN); else tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
    set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(err));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// 2D K-major operand map: dims {K, rows}, box {64, box_rows}, 128B swizzle, zero OOB fill.
int make_operand_map(CUtensorMap* map, int dtype, const void* base, int rows, int K, int ld,
                     int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -3;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt =
      dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rows=%d K=%d ld=%d base=%p", (int)r, rows, K, ld,
              base);
    return -3;
  }
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T, int BN, int CG, int EPI, int EW>
int launch_gemm(cudaStream_t stream, int dtype, const void* A, int lda, const void* W, int ldw,
                const GemmParams& p) {
  using L = SmemLayout<BN, CG, EW>;
  CUtensorMap tmA, tmB;
  int rc = make_operand_map(&tmA, dtype, A, p.M, p.K, lda, kBM);
  if (rc) return rc;
  rc = make_operand_map(&tmB, dtype, W, p.N, p.K, ldw, BN / CG);
  if (rc) return rc;
  auto kernel = gemm_tcgen05_kernel<T, BN, CG, EPI, EW>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm smem=%d): %s", L::kTotal, cudaGetErrorString(e));
      return -2;
    }
    attr_set = true;
  }
  const int m_tiles = (p.M + kBM * CG - 1) / (kBM * CG);
  const int n_tiles = (p.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  const int max_workers = num_sms() / CG;
  const int workers = tiles < max_workers ? tiles : max_workers;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(workers * CG));
  cfg.blockDim = dim3(64 + EW * 32);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e;
  {
    ProfScope ps(stream, kProfGemm, 2.0 * p.M * p.N * p.K,
                 2.0 * (static_cast<double>(p.M) * p.K + static_cast<double>(p.N) * p.K +
                        static_cast<double>(p.M) * p.N * (p.epi.residual ? 2 : 1)));
    e = cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, p);
  }
  count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("gemm launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <typename T, int EPI>
int dispatch_shape(cudaStream_t stream, int dtype, const void* A, int lda, const void* W, int ldw,
                   const GemmParams& p) {
  // 256x256 tiles on CTA pairs (cta_group::2: half the B traffic per CTA) when they fill the machine,
  // 128x256 single-CTA tiles next, 128x128 for small problems (more CTAs in flight).
  // SF_GEMM_MODE=1 forces single-CTA tiles (debug / A-B comparison).
  static const int mode = env_int("SF_GEMM_MODE", 0);
  constexpr int EW = 8;
  const int m_tiles = (p.M + kBM - 1) / kBM;
  const int m_tiles2 = (p.M + 2 * kBM - 1) / (2 * kBM);
  const int n_tiles = (p.N + 255) / 256;
  const bool pair = mode != 1 && (p.N >= 256) && (m_tiles2 * n_tiles >= num_sms() / 2);
  const bool wide = (p.N >= 256) && (m_tiles * n_tiles >= num_sms());
  if (pair) return launch_gemm<T, 256, 2, EPI, EW>(stream, dtype, A, lda, W, ldw, p);
  if (wide) return launch_gemm<T, 256, 1, EPI, EW>(stream, dtype, A, lda, W, ldw, p);
  return launch_gemm<T, 128, 1, EPI, EW>(stream, dtype, A, lda, W, ldw, p);
}

template <typename T>
int dispatch_epilogue(cudaStream_t stream, int dtype, const void* A, int lda, const void* W, int ldw,
                      const GemmParams& p) {
  const GemmEpilogue& e = p.epi;
  if (e.pos || e.time_emb) {
    if (e.residual || e.act != kActNone) { set_error("gemm: embed epilogue cannot be combined with residual/activation"); return -1; }
    return dispatch_shape<T, kEpiEmbed>(stream, dtype, A, lda, W, ldw, p);
  }
  if (e.residual) {
    if (e.act != kActNone) { set_error("gemm: residual epilogue cannot be combined with an activation"); return -1; }
    return dispatch_shape<T, kEpiResidual>(stream, dtype, A, lda, W, ldw, p);
  }
  if (e.act != kActNone) {
    if (e.gate) { set_error("gemm: activation epilogue cannot be combined with a gate"); return -1; }
    return dispatch_shape<T, kEpiAct>(stream, dtype, A, lda, W, ldw, p);
  }
  return dispatch_shape<T, kEpiBias>(stream, dtype, A, lda, W, ldw, p);
}

}  // namespace

int gemm(cudaStream_t stream, int dtype, const void* A, int lda, const void* W, int ldw, void* out,
         int ldo, int M, int N, int K, const GemmEpilogue& epi) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((K % 8) || (N % 8) || (lda % 8) || (ldw % 8) || (ldo % 8) || (epi.residual && (epi.ldr % 8))) {
    set_error("gemm: K, N and leading dims must be multiples of 8 (M=%d N=%d K=%d lda=%d ldw=%d ldo=%d)",
              M, N, K, lda, ldw, ldo);
    return -1;
  }
  if (dtype != kBF16 && dtype != kF16) {
    set_error("gemm: dtype must be bf16 or f16");
    return -1;
  }
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.out = out; p.ldo = ldo; p.epi = epi;
  if (dtype == kBF16) return dispatch_epilogue<__nv_bfloat16>(stream, dtype, A, lda, W, ldw, p);
  return dispatch_epilogue<__half>(stream, dtype, A, lda, W, ldw, p);
}

}  // namespace sf
