// gemm_tcgen05.cu — the workhorse of the StreamFormer encoder: every projection on the hot path
// (patch-embed conv-as-GEMM, temporal/spatial QKV, attention out-proj, temporal_dense, MLP fc1/fc2,
// pooling-head K/V/out/MLP; reference nn.Linear / nn.Conv2d call sites:
// models/modeling_timesformer_siglip.py:329-350, 513, 578, 629, 691, 728, 760, 811, 820, 830, 834,
// 895, 954, 1118-1124, 1135) runs through this one persistent, warp-specialised kernel:
//
//   warp 0        TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, mbarrier ring)
//   warp 1        MMA issuer     (tcgen05.mma kind::f16, fp32 accumulators in TMEM; with CG=2 one
//                                 256 x 256 x 16 instruction spans a CTA pair, cta_group::2)
//   warp 2        aux            (one tile ahead: bias / LN column-sum slices and per-row LayerNorm
//                                 statistics into shared memory, so the epilogue never waits on HBM)
//   warps 4..     epilogue       (tcgen05.ld, thread == row -> folded LayerNorm / bias / GELU /
//                                 pos+time embed / gated residual -> 256-bit global stores, one full
//                                 sector per lane; optional (b,t,n)<->(b,n,t) row permutation; partial
//                                 row statistics of the output for the next folded LayerNorm)
//
// TMEM holds two BN-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "sf_kernels.h"
#include "sf_ptx.cuh"
#include "sf_tma.h"

namespace sf {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;             // 64 x 2 B = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kChainMaxPhases = 4;

// epilogue flavours (compile-time: keeps each kernel's code small and branch-free)
enum EpiMode : int { kEpiBias = 0, kEpiAct = 1, kEpiResidual = 2, kEpiEmbed = 3 };

struct GemmParams {
  int M, N, K;
  void* out;
  int ldo;
  // tile walk: a work item = one M tile x `group` consecutive N tiles, processed back to back by the
  // same worker; on the slot path the first `res_kb` K blocks of the item's A tile stay resident
  // in shared memory for the whole item (fetched once instead of once per N tile)
  int group;
  int res_kb;
  GemmEpilogue epi;
};

// Walks the output tiles of one worker: items are dealt round-robin, the N tiles of an item in order.
struct TileWalk {
  int n_tiles, group, groups, num_items, step, item, j;
  __device__ TileWalk(int M, int N, int tile_m, int bn, int group_, int worker, int workers) {
    const int m_tiles = (M + tile_m - 1) / tile_m;
    n_tiles = (N + bn - 1) / bn;
    group = group_ < 1 ? 1 : group_;
    groups = (n_tiles + group - 1) / group;
    num_items = m_tiles * groups;
    step = workers;
    item = worker;
    j = 0;
  }
  __device__ bool valid() const { return item < num_items; }
  __device__ int m_tile() const { return item / groups; }
  __device__ int n_tile() const { return (item % groups) * group + j; }
  __device__ bool first_in_item() const { return j == 0; }
  __device__ bool last_in_item() const { return j + 1 == group || n_tile() + 1 >= n_tiles; }
  __device__ void next() {
    if (last_in_item()) { item += step; j = 0; } else { ++j; }
  }
};

// CG = CTAs cooperating on one tile (tcgen05 cta_group): 1 -> 128 x BN tile per CTA;
// 2 -> 256 x BN tile per CTA pair, each CTA staging its 128 A rows and BN/2 of the B rows.
// EW = epilogue warps (8 or 16).
template <int BN, int CG, int EW, bool TS = false>
struct SmemLayout {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = (BN / CG) * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // per accumulator stage, filled by the aux warp one tile ahead of the epilogue warps:
  //   bias slice [BN] fp32 | LN column-sum slice [BN] fp32 | per-row (-mean, rstd) [128] float2
  static constexpr int kAuxBias = 0;
  static constexpr int kAuxColsum = BN * 4;
  static constexpr int kAuxRows = 2 * BN * 4;
  static constexpr int kAuxBytesPerStage = 2 * BN * 4 + kBM * 8;
  static constexpr int kAuxBytes = 2 * kAuxBytesPerStage;
  static constexpr int kBudget = 232448 - 1024 - 256 - kAuxBytes;   // 227 KB minus alignment slack, barriers, aux
  static constexpr int kStages = (kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes);
  // Slot path (CTA pairs, 256 x 256 tiles): the A part and the B part of a K block are both 16 KB,
  // so operand memory is kSlots uniform 16 KB slots.  The first res_kb slots hold resident A K
  // blocks of the current work item, the rest form the ring every other operand block streams through.
  static constexpr bool kSlotPath = (CG == 2 && BN == 256);
  static constexpr int kSlotBytes = 16384;
  // TS (TMA-store epilogue): every epilogue warp owns three staging buffers of 32 rows x one chunk
  // (32 or 16 columns) through which its output — and its residual, when there is one — moves
  // between registers and global memory as whole TMA boxes.
  static constexpr int kChunkCols = (EW == 16) ? 16 : 32;
  static constexpr int kStageBufBytes = 32 * kChunkCols * 2;
  static constexpr int kStagingBytes = TS ? EW * 3 * kStageBufBytes : 0;
  static constexpr int kSlots = TS ? 10 : 13;
  static constexpr int kMaxRes = kSlots - 4;        // leaves a ring of >= 4 slots
  static constexpr int kBarBytes = kSlotPath ? 1024 : 256;
  static constexpr int kBarOffset = kSlotPath ? kSlots * kSlotBytes : kStages * kStageBytes;
  static constexpr int kAuxOffset = kBarOffset + kBarBytes;
  static constexpr int kStagingOffset = kAuxOffset + kAuxBytes;
  static constexpr int kTotal = kStagingOffset + kStagingBytes + 1024;
  static_assert(!TS || kSlotPath, "the TMA-store epilogue exists on the slot path only");
  static_assert(kStages >= 3, "not enough shared memory for a 3-stage pipeline");
  static_assert(2 * kStages + 8 <= 31, "barrier area overflow");
  static_assert(!kSlotPath || (kABytes == kSlotBytes && kBBytes == kSlotBytes), "slot path needs 16 KB operand blocks");
  static_assert(!kSlotPath || (2 * kSlots + 2 * kMaxRes + 8 + 1 + 3 * EW) * 8 <= kBarBytes, "slot barrier area overflow");
  static_assert(kTotal <= 232448, "shared memory budget exceeded");
};

template <typename T> struct UmmaFmt;
template <> struct UmmaFmt<__half> { static constexpr int value = 0; };
template <> struct UmmaFmt<__nv_bfloat16> { static constexpr int value = 1; };

// erf GELU (hidden_act="gelu", reference ACT2FN["gelu"], …siglip.py:814-817) as x * Phi(x) with
// Phi(x) = 1 / (1 + 2^(x * P(x^2))), P the degree-5 minimax fit (in x^2) of -log2(e) * logit(Phi(x)) / x
// on |x| <= 4.5 (gelu(-4.5) = -1.5e-5, Phi(4.5) = 1 - 3.4e-6).  Relative error of gelu <= 1.8e-4 on
// the fitted range and absolute error <= 3.5e-5 everywhere, i.e. below the bf16 and fp16 output
// rounding; 12 instructions with 2 MUFU (ex2, rcp) instead of the 19 of an Abramowitz-Stegun erf —
// the fc1 epilogue is issue/MUFU-bound, not MMA-bound, otherwise.
__device__ __forceinline__ float gelu_erf(float x) {
  // only x^2 is clamped: beyond |x| = 4.5 the exponent keeps growing linearly in x with the slope
  // reached at the boundary, so Phi still tends to 1 (x > 0) and to 0 faster than 1/|x| (x < 0)
  const float t = fminf(x * x, 20.25f);
  float p = fmaf(5.838623416e-08f, t, -3.653467351e-06f);
  p = fmaf(p, t, 7.064459774e-05f);
  p = fmaf(p, t, 5.841390203e-04f);
  p = fmaf(p, t, -1.059873517e-01f);
  p = fmaf(p, t, -2.301437105e+00f);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * p));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return x * r;
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * (x + k1 * x * x * x);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
  const float h = 0.5f * x;
  return fmaf(h, th, h);
}

// ----------------------------------------------------------------------------- packed fp32 pairs (FFMA2 / FMUL2 / FADD2)
// sm_100 issues one instruction for two fp32 lanes held in an aligned register pair.  The epilogue warps are
// ISSUE-bound in the GELU epilogue (fc1: ~23 instructions per output element against a K = 768 main loop), so the
// bias / LayerNorm-fold / GELU-polynomial / residual arithmetic runs on pairs: the same IEEE operations in the same
// order per element (results are bit-identical to the scalar form), about 0.7 of the instructions.
#ifdef SF_GEMM_TIMELINE
#define SF_GTL(...) __VA_ARGS__
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#else
#define SF_GTL(...)
#endif
using f32x2 = unsigned long long;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// gelu_erf() above on a pair
__device__ __forceinline__ f32x2 gelu_erf2(f32x2 x) {
  float t0, t1;
  upk2(mul2(x, x), t0, t1);
  const f32x2 t = pk2(fminf(t0, 20.25f), fminf(t1, 20.25f));
  f32x2 p = fma2(pk2(5.838623416e-08f, 5.838623416e-08f), t, pk2(-3.653467351e-06f, -3.653467351e-06f));
  p = fma2(p, t, pk2(7.064459774e-05f, 7.064459774e-05f));
  p = fma2(p, t, pk2(5.841390203e-04f, 5.841390203e-04f));
  p = fma2(p, t, pk2(-1.059873517e-01f, -1.059873517e-01f));
  p = fma2(p, t, pk2(-2.301437105e+00f, -2.301437105e+00f));
  float a0, a1, e0, e1, r0, r1;
  upk2(mul2(x, p), a0, a1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  upk2(add2(pk2(e0, e1), pk2(1.0f, 1.0f)), a0, a1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(a0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(a1));
  return mul2(x, pk2(r0, r1));
}
// Eight consecutive accumulator columns of one row as four pairs, and the epilogue stages on them.
struct Epi8 {
  f32x2 v[4];
  __device__ __forceinline__ void load(const uint32_t* raw) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = pk2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
  }
  // v = rstd * (v + nmean * colsum) + bias      (folded LayerNorm)
  __device__ __forceinline__ void ln_fold(float nmean, float rstd, const float4& c0, const float4& c1, const float4& b0,
                                          const float4& b1) {
    const f32x2 nm = pk2(nmean, nmean), rs = pk2(rstd, rstd);
    v[0] = fma2(rs, fma2(nm, pk2(c0.x, c0.y), v[0]), pk2(b0.x, b0.y));
    v[1] = fma2(rs, fma2(nm, pk2(c0.z, c0.w), v[1]), pk2(b0.z, b0.w));
    v[2] = fma2(rs, fma2(nm, pk2(c1.x, c1.y), v[2]), pk2(b1.x, b1.y));
    v[3] = fma2(rs, fma2(nm, pk2(c1.z, c1.w), v[3]), pk2(b1.z, b1.w));
  }
  __device__ __forceinline__ void add_bias(const float4& b0, const float4& b1) {
    v[0] = add2(v[0], pk2(b0.x, b0.y)); v[1] = add2(v[1], pk2(b0.z, b0.w));
    v[2] = add2(v[2], pk2(b1.x, b1.y)); v[3] = add2(v[3], pk2(b1.z, b1.w));
  }
  __device__ __forceinline__ void add8(const float4& q0, const float4& q1) { add_bias(q0, q1); }
  __device__ __forceinline__ void gelu_erf_() {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = gelu_erf2(v[j]);
  }
  __device__ __forceinline__ void gelu_tanh_() {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a, b;
      upk2(v[j], a, b);
      v[j] = pk2(gelu_tanh(a), gelu_tanh(b));
    }
  }
  __device__ __forceinline__ void scale(float g) {
    const f32x2 gg = pk2(g, g);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = mul2(v[j], gg);
  }
  // v = g * v + residual   (residual: 8 values of the activation dtype in 4 words)
  template <typename T>
  __device__ __forceinline__ void residual(float g, const uint32_t* rq) {
    const f32x2 gg = pk2(g, g);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 r = Pack2<T>::unpack(rq[j]);
      v[j] = fma2(gg, v[j], pk2(r.x, r.y));
    }
  }
  template <typename T>
  __device__ __forceinline__ void pack(uint32_t* oq) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a, b;
      upk2(v[j], a, b);
      oq[j] = Pack2<T>::pack(a, b);
    }
  }
};

// nearest-neighbour source index, identical arithmetic to F.interpolate(mode="nearest"):
// scale = float(in)/out ; src = min(floor(dst*scale), in-1)
__device__ __forceinline__ int time_index(int t_abs, int time_len, int time_total) {
  if (time_total <= time_len) return t_abs;
  float scale = static_cast<float>(time_len) / static_cast<float>(time_total);
  int s = static_cast<int>(floorf(static_cast<float>(t_abs) * scale));
  return s < time_len - 1 ? s : time_len - 1;
}

// output row of GEMM row m under the (b,t,n) <-> (b,n,t) permutations
__device__ __forceinline__ long map_out_row(int m, int row_map, int Tn, int Sn) {
  if (row_map == kRowBTNtoBNT) {
    const int n = m % Sn, bt = m / Sn;
    return (static_cast<long>(bt / Tn) * Sn + n) * Tn + bt % Tn;
  }
  if (row_map == kRowBNTtoBTN) {
    const int t = m % Tn, bn = m / Tn;
    return (static_cast<long>(bn / Sn) * Tn + t) * Sn + bn % Sn;
  }
  return m;
}

__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 q;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(addr));
  return q;
}
__device__ __forceinline__ void sts128f(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void lds128u(uint32_t addr, uint32_t* v) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void sts128u(uint32_t addr, const uint32_t* v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
// 16 consecutive 2-byte elements of one row <-> 8 registers.  `wide`: the row is 32-byte aligned, so
// one 256-bit access (a full 32-byte sector per lane; sm_100 LDG/STG.256) moves them; otherwise two
// 128-bit accesses.  ncols = valid columns at p (a multiple of 8; < 16 only on a ragged N edge).
__device__ __forceinline__ void ldg_cols16(uint32_t (&v)[8], const void* p, int ncols, bool wide) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0u;
  if (wide && ncols >= 16) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
  } else {
    if (ncols >= 8) {
      const uint4 q = *reinterpret_cast<const uint4*>(p);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    if (ncols >= 16) {
      const uint4 q = *(reinterpret_cast<const uint4*>(p) + 1);
      v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
    }
  }
}
__device__ __forceinline__ void stg_cols16(void* p, const uint32_t (&v)[8], int ncols, bool wide) {
  if (wide && ncols >= 16) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
  } else {
    if (ncols >= 8) *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1], v[2], v[3]);
    if (ncols >= 16) *(reinterpret_cast<uint4*>(p) + 1) = make_uint4(v[4], v[5], v[6], v[7]);
  }
}

template <typename T, int BN, int CG, int EPI, int EW, bool TS>
__global__ void __launch_bounds__(128 + EW * 32, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                    const GemmParams p) {
  using L = SmemLayout<BN, CG, EW, TS>;
  SF_GTL(const unsigned long long tl_entry = gtimer();)
  constexpr int kStages = L::kStages;
  constexpr int kTileM = kBM * CG;
  constexpr bool kLnCapable = (EPI == kEpiBias || EPI == kEpiAct);
  constexpr bool kStatsCapable = (EPI == kEpiResidual || EPI == kEpiEmbed);
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  constexpr bool kSlotPath = L::kSlotPath;
  constexpr int kRingBars = kSlotPath ? L::kSlots : kStages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);   // stage ring / slot ring
  uint64_t* empty_bar = full_bar + kRingBars;
  uint64_t* res_full = empty_bar + kRingBars;                               // slot path: resident A blocks
  uint64_t* res_empty = res_full + (kSlotPath ? L::kMaxRes : 0);
  uint64_t* tmem_full = res_empty + (kSlotPath ? L::kMaxRes : 0);
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* aux_full = tmem_empty + 2;
  uint64_t* aux_empty = aux_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_empty + 2);
  uint64_t* res_bar = aux_empty + 3;     // TS + residual: three per epilogue warp (one per staging buffer)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;   // rank inside the CTA pair
  const bool leader = cta_rank == 0;
  const int worker = blockIdx.x / CG;          // one worker = one CTA (CG=1) or one CTA pair (CG=2)
  const int num_workers = gridDim.x / CG;

  const int k_blocks = (p.K + kBK - 1) / kBK;
  const int group = kSlotPath ? p.group : 1;
  // resident A K blocks (slot path only); the ring keeps the remaining slots
  const int res_kb = kSlotPath ? (p.res_kb < k_blocks ? p.res_kb : k_blocks) : 0;
  const int ring = L::kSlots - res_kb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kRingBars; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    if constexpr (kSlotPath) {
      for (int s = 0; s < L::kMaxRes; ++s) {
        mbar_init(&res_full[s], 1);
        mbar_init(&res_empty[s], 1);
      }
    }
    if constexpr (TS) {
      tma_prefetch_desc(&tmOut);
      if constexpr (EPI == kEpiResidual) {
        tma_prefetch_desc(&tmRes);
        for (int i = 0; i < 3 * EW; ++i) mbar_init(&res_bar[i], 1);
      }
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], EW * CG);   // (leader's copy) epilogue warps of every CTA of the pair
      mbar_init(&aux_full[a], 1);
      mbar_init(&aux_empty[a], EW);
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
  // Weight operand of this CTA's first tile, started ahead of the dependency wait (GemmEpilogue::w_static): the first
  // pre_b K blocks of W go into the stages / ring slots the main loop would put them in anyway.  Slot path: only K
  // blocks whose A part is resident (kb < res_kb) — for those the ring sees B blocks only, in this order.
  int pre_b = 0;
  if (p.epi.w_static) {
    TileWalk w0(p.M, p.N, kTileM, BN, group, worker, num_workers);
    if (w0.valid()) {
      if constexpr (kSlotPath) pre_b = res_kb < ring ? res_kb : ring;
      else pre_b = kStages < k_blocks ? kStages : k_blocks;
      if (warp == 0 && elect_one_sync()) {
        const int n0 = w0.n_tile() * BN + static_cast<int>(cta_rank) * (BN / CG);
        for (int kb = 0; kb < pre_b; ++kb) {
          if constexpr (kSlotPath) {
            if (leader) mbar_arrive_expect_tx(&full_bar[kb], 2 * L::kSlotBytes);
            tma_load_2d_cg2(smem + (res_kb + kb) * L::kSlotBytes, &tmB, &full_bar[kb], kb * kBK, n0);
          } else if constexpr (CG == 2) {
            if (leader) mbar_arrive_expect_tx(&full_bar[kb], 2 * L::kStageBytes);
            tma_load_2d_cg2(smem + kb * L::kStageBytes + L::kABytes, &tmB, &full_bar[kb], kb * kBK, n0);
          } else {
            mbar_arrive_expect_tx(&full_bar[kb], L::kStageBytes);
            tma_load_2d(smem + kb * L::kStageBytes + L::kABytes, &tmB, &full_bar[kb], kb * kBK, n0);
          }
        }
      }
      __syncwarp();
    }
  }
  // PDL: everything above overlapped the tail of the previous kernel; from here on we read its output
  SF_GTL(const unsigned long long tl_prologue = gtimer();)
  griddep_wait();
  SF_GTL(const unsigned long long tl_dep = gtimer();)
  griddep_launch_dependents();

  const GemmEpilogue& e = p.epi;
  const bool ln = kLnCapable && e.ln_stats != nullptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (every CTA)
    if constexpr (kSlotPath) {
      if (elect_one_sync()) {
        int slot = 0;              // ring position (slots res_kb .. kSlots-1)
        uint32_t phase = 0;
        int items_done = 0;
        uint8_t* ring_base = smem + res_kb * L::kSlotBytes;
        auto ring_load = [&](const CUtensorMap* tm, int c0, int c1) {
          mbar_wait(&empty_bar[slot], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[slot], 2 * L::kSlotBytes);
          tma_load_2d_cg2(ring_base + slot * L::kSlotBytes, tm, &full_bar[slot], c0, c1);
          if (++slot == ring) { slot = 0; phase ^= 1; }
        };
        bool first_tile = true;
        for (TileWalk w(p.M, p.N, kTileM, BN, group, worker, num_workers); w.valid(); w.next()) {
          const int m0 = w.m_tile() * kTileM + static_cast<int>(cta_rank) * kBM;
          const int n0 = w.n_tile() * BN + static_cast<int>(cta_rank) * (BN / CG);
          const bool first = w.first_in_item();
          for (int kb = 0; kb < k_blocks; ++kb) {
            if (kb < res_kb) {
              if (first) {   // A block fetched once per item, into its resident slot
                mbar_wait(&res_empty[kb], (items_done & 1) ^ 1);
                if (leader) mbar_arrive_expect_tx(&res_full[kb], 2 * L::kSlotBytes);
                tma_load_2d_cg2(smem + kb * L::kSlotBytes, &tmA, &res_full[kb], kb * kBK, m0);
              }
            } else {
              ring_load(&tmA, kb * kBK, m0);
            }
            if (first_tile && kb < pre_b) {          // this W block went out ahead of griddepcontrol.wait
              if (++slot == ring) { slot = 0; phase ^= 1; }
            } else {
              ring_load(&tmB, kb * kBK, n0);
            }
          }
          first_tile = false;
          if (w.last_in_item()) ++items_done;
        }
      }
    } else if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      bool first_tile = true;
      for (TileWalk w(p.M, p.N, kTileM, BN, 1, worker, num_workers); w.valid(); w.next(), first_tile = false) {
        const int m0 = w.m_tile() * kTileM + static_cast<int>(cta_rank) * kBM;
        const int n0 = w.n_tile() * BN + static_cast<int>(cta_rank) * (BN / CG);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          const bool w_sent = first_tile && kb < pre_b;    // expect_tx armed and W loaded ahead of the dependency wait
          if constexpr (CG == 2) {
            // the leader's barrier counts the bytes of both CTAs' loads for this stage
            if (leader && !w_sent) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            tma_load_2d_cg2(sa, &tmA, &full_bar[stage], kb * kBK, m0);
            if (!w_sent) tma_load_2d_cg2(sb, &tmB, &full_bar[stage], kb * kBK, n0);
          } else {
            if (!w_sent) mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * kBK, m0);
            if (!w_sent) tma_load_2d(sb, &tmB, &full_bar[stage], kb * kBK, n0);
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
    constexpr uint32_t idesc = umma_idesc_f16(kTileM, BN, UmmaFmt<T>::value);
    if constexpr (kSlotPath) {
      if (leader && elect_one_sync()) {
        int slot = 0;
        uint32_t phase = 0;
        int items_done = 0, it = 0;
        const uint32_t ring_u = smem_u32(smem + res_kb * L::kSlotBytes);
        for (TileWalk w(p.M, p.N, kTileM, BN, group, worker, num_workers); w.valid(); w.next(), ++it) {
          const int acc = it & 1;
          const uint32_t acc_phase = (it >> 1) & 1;
          const bool first = w.first_in_item(), last = w.last_in_item();
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kb = 0; kb < k_blocks; ++kb) {
            uint32_t sa;
            int a_slot = -1;
            if (kb < res_kb) {
              if (first) mbar_wait(&res_full[kb], items_done & 1);
              sa = smem_u32(smem + kb * L::kSlotBytes);
            } else {
              mbar_wait(&full_bar[slot], phase);
              a_slot = slot;
              sa = ring_u + slot * L::kSlotBytes;
              if (++slot == ring) { slot = 0; phase ^= 1; }
            }
            mbar_wait(&full_bar[slot], phase);
            const int b_slot = slot;
            const uint32_t sb = ring_u + slot * L::kSlotBytes;
            if (++slot == ring) { slot = 0; phase ^= 1; }
            tc_fence_after();
            const uint64_t da = umma_desc_sw128_kmajor(sa);
            const uint64_t db = umma_desc_sw128_kmajor(sb);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k)
              umma_f16_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            // hand the operand slots back (in both CTAs of the pair) when these MMAs retire
            if (a_slot >= 0) umma_commit_cg2_mc(&empty_bar[a_slot], 0x3);
            umma_commit_cg2_mc(&empty_bar[b_slot], 0x3);
            if (kb < res_kb && last) umma_commit_cg2_mc(&res_empty[kb], 0x3);
          }
          umma_commit_cg2_mc(&tmem_full[acc], 0x3);
          if (last) ++items_done;
        }
      }
    } else if (leader && elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      SF_GTL(unsigned long long tl_first = 0;)
      for (TileWalk w(p.M, p.N, kTileM, BN, 1, worker, num_workers); w.valid(); w.next(), ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          SF_GTL(if (kb == 0 && it == 0) tl_first = gtimer();)
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
        SF_GTL(if (it == 0 && blockIdx.x == 0 && p.M < 1024) printf("gemm N=%d K=%d epi=%d: prologue +%llu dep_wait +%llu first_stage +%llu mma issued +%llu ns\n", p.N, p.K, EPI, tl_prologue - tl_entry, tl_dep - tl_entry, tl_first - tl_entry, gtimer() - tl_entry);)
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ aux warp
    // Runs one tile ahead of the epilogue warps and stages everything their arithmetic needs that
    // is not the accumulator, so no global-memory latency sits on the epilogue's critical path:
    // the bias slice and the LN column-sum slice of the tile's BN columns, and for a folded
    // LayerNorm (-mean, rstd) of the tile's 128 rows, reduced from the producer's partials.
    const float inv_k = 1.0f / static_cast<float>(p.K);
    int it = 0;
    for (TileWalk w(p.M, p.N, kTileM, BN, group, worker, num_workers); w.valid(); w.next(), ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int m0 = w.m_tile() * kTileM + static_cast<int>(cta_rank) * kBM;
      const int n0 = w.n_tile() * BN;
      if constexpr (EPI == kEpiResidual && !TS) {
        // pull the tile's residual rows into L2 a whole tile ahead of the epilogue's loads
        const int left = p.N - n0;
        const uint32_t bytes = static_cast<uint32_t>((left < BN ? left : BN) * 2) & ~15u;
        const T* resb = reinterpret_cast<const T*>(e.residual);
#pragma unroll
        for (int rr = 0; rr < kBM / 32; ++rr) {
          const int m = m0 + rr * 32 + lane;
          if (m < p.M && bytes)
            prefetch_l2_bulk(resb + map_out_row(m, e.row_map, e.T, e.S) * e.ldr + n0, bytes);
        }
      }
      // Issue every global load of the tile first (bias / column-sum slices and all row-statistics
      // partials of the 128 rows), then wait for the stage: one memory round trip per tile instead
      // of one per row group — this warp must never be slower than the epilogue it feeds.
      constexpr int kVec = BN >= 128 ? BN / 4 / 32 : 1; // float4 per lane per table (2 for BN=256, 1 for 128; BN=64: lanes 0-15)
      float4 b4[kVec], c4[kVec];
#pragma unroll
      for (int v = 0; v < kVec; ++v) {
        const int col = n0 + (v * 32 + lane) * 4;
        b4[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        c4[v] = b4[v];
        if (col < p.N && (v * 32 + lane) * 4 < BN) {
          if (e.bias) b4[v] = __ldg(reinterpret_cast<const float4*>(e.bias + col));
          if (ln) c4[v] = __ldg(reinterpret_cast<const float4*>(e.ln_colsum + col));
        }
      }
      // row statistics: up to 12 partials per row (64-column producers) held in registers, two row
      // groups per round trip
      constexpr int kMaxParts = 12;
      float s1x[kBM / 32], s2x[kBM / 32];
      if constexpr (kLnCapable) {
        if (ln) {
#pragma unroll
          for (int half = 0; half < kBM / 64; ++half) {
            float2 t[2][kMaxParts];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int m = m0 + (half * 2 + rr) * 32 + lane;
#pragma unroll
              for (int q = 0; q < kMaxParts; ++q) {
                t[rr][q] = make_float2(0.f, 0.f);
                if (q < e.ln_parts && m < p.M) t[rr][q] = __ldg(e.ln_stats + static_cast<long>(q) * p.M + m);
              }
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int m = m0 + (half * 2 + rr) * 32 + lane;
              float s1 = 0.f, s2 = 0.f;
#pragma unroll
              for (int q = 0; q < kMaxParts; ++q) {
                s1 += t[rr][q].x;
                s2 += t[rr][q].y;
              }
              for (int q = kMaxParts; q < e.ln_parts; ++q) {   // (never with the shipped tile shapes)
                if (m < p.M) {
                  const float2 u = __ldg(e.ln_stats + static_cast<long>(q) * p.M + m);
                  s1 += u.x;
                  s2 += u.y;
                }
              }
              s1x[half * 2 + rr] = s1;
              s2x[half * 2 + rr] = s2;
            }
          }
        }
      }
      mbar_wait(&aux_empty[st], ph ^ 1);
      const uint32_t aux_u = smem_u32(smem + L::kAuxOffset + st * L::kAuxBytesPerStage);
#pragma unroll
      for (int v = 0; v < kVec; ++v) {
        if ((v * 32 + lane) * 4 < BN) {
          sts128f(aux_u + L::kAuxBias + (v * 32 + lane) * 16, b4[v]);
          if constexpr (kLnCapable) sts128f(aux_u + L::kAuxColsum + (v * 32 + lane) * 16, c4[v]);
        }
      }
      if constexpr (kLnCapable) {
        if (ln) {
#pragma unroll
          for (int rr = 0; rr < kBM / 32; ++rr) {
            const float mean = s1x[rr] * inv_k;
            const float var = fmaxf(fmaf(s2x[rr], inv_k, -mean * mean), 0.f);
            const float rstd = rsqrtf(var + e.ln_eps);
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(aux_u + L::kAuxRows + (rr * 32 + lane) * 8), "f"(-mean), "f"(rstd)
                         : "memory");
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&aux_full[st]);
    }
  } else if (warp >= 4 && TS) {
    // ------------------------------------------------------------------ epilogue warps, TMA-store path
    // Thread == accumulator row (TMEM lane), 32- (or 16-) column chunks as below, but nothing touches
    // global memory with per-thread accesses (a row-per-lane 32-byte store costs the LSU data pipe
    // ~45 wavefronts per KB and made the epilogue, not the tensor pipe, the bound of every K=768
    // GEMM): a chunk is written to a swizzled staging buffer (conflict-free STS.128) and leaves as
    // ONE TMA box store per warp; the residual arrives the same way, requested two chunks ahead
    // into the buffer the output will later leave from.  Three buffers per warp rotate.
    if constexpr (TS) {
    const int ew = warp - 4;
    const int quarter = warp & 3;
    const int colgrp = ew >> 2;
    constexpr int kColsPerWarp = BN / (EW / 4);
    constexpr int kCW = L::kChunkCols;
    constexpr int kChunks = kColsPerWarp / kCW;
    constexpr int kRowBytes = kCW * 2;
    constexpr int kBufBytes = L::kStageBufBytes;
    constexpr int kC16 = kRowBytes / 16;            // 16-byte pieces per staged row
    static_assert(kChunks % 2 == 0, "chunks are processed in double-buffered pairs");
    const bool want_stats = kStatsCapable && e.stats_out != nullptr;
    const float gscale = e.gate ? tanhf(__ldg(e.gate)) : 1.0f;
    uint8_t* stg = smem + L::kStagingOffset + ew * (3 * kBufBytes);
    const uint32_t stg_u = smem_u32(stg);
    uint64_t* rbar = res_bar + ew * 3;
    // CU_TENSOR_MAP_SWIZZLE_64B / _32B: 16-byte piece index ^= (row / (128 / row bytes)) mod pieces
    const uint32_t swz = (static_cast<uint32_t>(lane) / (128 / kRowBytes)) & (kC16 - 1);
    const uint32_t my_row_u = static_cast<uint32_t>(lane) * kRowBytes;
    uint32_t f = 0;        // chunks processed by this warp so far: buffer f % 3, barrier parity (f / 3) & 1
    // residual look-ahead: (tile, chunk) of the next residual box to request (lane 0 only)
    TileWalk wl(p.M, p.N, kTileM, BN, group, worker, num_workers);
    int lc = 0;
    uint32_t lf = 0;
    auto issue_res = [&]() {
      if (!wl.valid()) return;
      const int lrow = wl.m_tile() * kTileM + static_cast<int>(cta_rank) * kBM + quarter * 32;
      const int lcol = wl.n_tile() * BN + colgrp * kColsPerWarp + lc * kCW;
      const uint32_t b = lf % 3;
      mbar_arrive_expect_tx(&rbar[b], kBufBytes);
      tma_load_2d(stg + b * kBufBytes, &tmRes, &rbar[b], lcol, lrow);
      ++lf;
      if (++lc == kChunks) { lc = 0; wl.next(); }
    };
    if constexpr (EPI == kEpiResidual) {
      if (lane == 0) { issue_res(); issue_res(); }
    }
    int it = 0;
    for (TileWalk w(p.M, p.N, kTileM, BN, group, worker, num_workers); w.valid(); w.next(), ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = w.m_tile() * kTileM + static_cast<int>(cta_rank) * kBM;
      const int n0 = w.n_tile() * BN;
      const int row0 = m0 + quarter * 32;
      const int m = row0 + lane;
      const bool row_ok = m < p.M;
      const int wcol0 = n0 + colgrp * kColsPerWarp;
      const uint32_t aux_u = smem_u32(smem + L::kAuxOffset + acc * L::kAuxBytesPerStage);
      const uint32_t bias_u = aux_u + L::kAuxBias + colgrp * kColsPerWarp * 4;
      const uint32_t csum_u = aux_u + L::kAuxColsum + colgrp * kColsPerWarp * 4;
      mbar_wait(&aux_full[acc], acc_phase);
      float nmean = 0.f, rstd = 1.f;
      if constexpr (kLnCapable) {
        if (ln) {
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(nmean), "=f"(rstd)
                       : "r"(aux_u + L::kAuxRows + (quarter * 32 + lane) * 8));
        }
      }
      float st1 = 0.f, st2 = 0.f;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN +
                              colgrp * kColsPerWarp;
      uint32_t raw[2][kCW];
      auto tmem_fetch = [&](int c, uint32_t (&dst)[kCW]) {
        if constexpr (kCW == 32) tmem_ld_32x32b_x32(t_base + c * kCW, dst);
        else tmem_ld_32x32b_x16(t_base + c * kCW, dst);
      };
      tmem_fetch(0, raw[0]);
#pragma unroll 1
      for (int cp = 0; cp < kChunks / 2; ++cp) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = cp * 2 + h;
          const int col0 = wcol0 + c * kCW;
          tmem_ld_wait();                                     // chunk c is in registers
          if (c + 1 < kChunks) {
            tmem_fetch(c + 1, raw[h ^ 1]);
          } else {
            // last TMEM read of this accumulator is done: hand it back to the MMA warp now, the
            // rest of the tile's epilogue overlaps the next-but-one tile's main loop
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CG == 2) mbar_arrive_remote(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
            }
          }
          const uint32_t buf = f % 3;
          const uint32_t buf_u = stg_u + buf * kBufBytes;
          uint32_t rb[kCW / 2];
          if constexpr (EPI == kEpiResidual) {
            mbar_wait(&rbar[buf], (f / 3) & 1);
#pragma unroll
            for (int q = 0; q < kC16; ++q) lds128u(buf_u + my_row_u + ((static_cast<uint32_t>(q) ^ swz) << 4), &rb[q * 4]);
          }
          uint32_t ob[kCW / 2];
#pragma unroll
          for (int g = 0; g < kCW / 8; ++g) {  // groups of 8 columns
            Epi8 x8;                                 // packed pairs: see f32x2 above
            x8.load(&raw[h][g * 8]);
            const float4 b0 = lds128f(bias_u + (c * kCW + g * 8) * 4);
            const float4 b1 = lds128f(bias_u + (c * kCW + g * 8 + 4) * 4);
            bool did_ln = false;
            if constexpr (kLnCapable) {
              if (ln) {
                const float4 c0 = lds128f(csum_u + (c * kCW + g * 8) * 4);
                const float4 c1 = lds128f(csum_u + (c * kCW + g * 8 + 4) * 4);
                x8.ln_fold(nmean, rstd, c0, c1, b0, b1);
                did_ln = true;
              }
            }
            if (!did_ln) x8.add_bias(b0, b1);
            if constexpr (EPI == kEpiAct) {
              if (e.act == kActGeluErf) x8.gelu_erf_();
              else x8.gelu_tanh_();
            }
            if constexpr (EPI == kEpiResidual) {
              x8.template residual<T>(gscale, &rb[g * 4]);
            } else if constexpr (EPI == kEpiBias) {
              if (e.gate) x8.scale(gscale);
            }
            uint32_t* oq = &ob[g * 4];
            x8.template pack<T>(oq);
            if constexpr (kStatsCapable) {
              if (want_stats) {   // statistics of the values as stored (rounded), what the next GEMM multiplies
                const float2 q0 = Pack2<T>::unpack(oq[0]), q1 = Pack2<T>::unpack(oq[1]);
                const float2 q2 = Pack2<T>::unpack(oq[2]), q3 = Pack2<T>::unpack(oq[3]);
                st1 += ((q0.x + q0.y) + (q1.x + q1.y)) + ((q2.x + q2.y) + (q3.x + q3.y));
                st2 = fmaf(q0.x, q0.x, fmaf(q0.y, q0.y, fmaf(q1.x, q1.x, fmaf(q1.y, q1.y, st2))));
                st2 = fmaf(q2.x, q2.x, fmaf(q2.y, q2.y, fmaf(q3.x, q3.x, fmaf(q3.y, q3.y, st2))));
              }
            }
          }
          // the staging buffer changes hands: residual in -> output out
          if constexpr (EPI == kEpiResidual) {
            __syncwarp();                                   // every lane has read its residual row
          } else {
            if (lane == 0) tma_store_wait_read<2>();        // the store that last used this buffer has drained it
            __syncwarp();
          }
#pragma unroll
          for (int q = 0; q < kC16; ++q) sts128u(buf_u + my_row_u + ((static_cast<uint32_t>(q) ^ swz) << 4), &ob[q * 4]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (row0 < p.M) tma_store_2d(&tmOut, stg + buf * kBufBytes, col0, row0);
            tma_store_commit();
            if constexpr (EPI == kEpiResidual) {
              tma_store_wait_read<1>();                     // buffers of chunks <= f-1 are free again
              issue_res();                                  // residual of chunk f+2 -> buffer (f+2) % 3
            }
          }
          ++f;
        }
      }
      if constexpr (kStatsCapable) {
        if (want_stats && row_ok)
          e.stats_out[static_cast<long>(wcol0 / kColsPerWarp) * p.M + m] = make_float2(st1, st2);
      }
      // all reads of the aux stage are complete -> hand it back (the accumulator went back above)
      __syncwarp();
      if (lane == 0) mbar_arrive(&aux_empty[acc]);
    }
    if (lane == 0) tma_store_wait<0>();   // staged data must stay valid until the last store has read it
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue warps
    // Each warp owns 32 accumulator rows (its TMEM lane quarter, thread == row) x kColsPerWarp
    // columns and walks them in 32-column chunks, software-pipelined: the tcgen05.ld of chunk c+1
    // and the residual loads of chunk c+1 are in flight while chunk c is processed.  A thread's 32
    // outputs of a chunk are 64 contiguous bytes of its row: they leave as two 256-bit stores (one
    // full 32-byte sector each), so no shared-memory transpose is needed; the residual arrives the
    // same way.
    const int ew = warp - 4;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int colgrp = ew >> 2;            // which slice of the BN columns
    constexpr int kColsPerWarp = BN / (EW / 4);
    constexpr int kCW = (EW == 16) ? 16 : 32;       // chunk width: 16 warps have half the registers each
    constexpr int kChunks = kColsPerWarp / kCW;
    constexpr int kPieces = kCW / 16;               // 16-column (32-byte) pieces per chunk
    static_assert(kChunks % 2 == 0, "chunks are processed in double-buffered pairs");
    const bool want_stats = kStatsCapable && e.stats_out != nullptr;
    const float gscale = e.gate ? tanhf(__ldg(e.gate)) : 1.0f;
    T* out = reinterpret_cast<T*>(p.out);
    const T* res = reinterpret_cast<const T*>(e.residual);
    const bool wide_out = ((reinterpret_cast<uintptr_t>(p.out) & 31) == 0) && (p.ldo % 16 == 0);
    const bool wide_res = ((reinterpret_cast<uintptr_t>(e.residual) & 31) == 0) && (e.ldr % 16 == 0);
    int it = 0;
    for (TileWalk w(p.M, p.N, kTileM, BN, group, worker, num_workers); w.valid(); w.next(), ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = w.m_tile() * kTileM + static_cast<int>(cta_rank) * kBM;
      const int n0 = w.n_tile() * BN;
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
        if (e.time_emb) {
          int t_off = e.time_off, t_total = e.time_total;
          if (e.time_off_dev) {
            t_off = __ldg(e.time_off_dev);
            t_total = t_off + e.T > e.time_horizon ? t_off + e.T : e.time_horizon;
          }
          time_row = e.time_emb + static_cast<long>(time_index(t_off + frame, e.time_len, t_total)) * p.N;
        }
      }
      const int wcol0 = n0 + colgrp * kColsPerWarp;
      T* orow = out + r * p.ldo;
      const T* rrow = res + r * e.ldr;
      // valid columns (multiple of 8) of the 16-column piece starting at `col`
      auto ncols_at = [&](int col) -> int {
        if (!row_ok) return 0;
        const int left = p.N - col;
        return left >= 16 ? 16 : (left > 0 ? left : 0);
      };
      uint32_t rbuf[2][kPieces][8];
      if constexpr (EPI == kEpiResidual) {
#pragma unroll
        for (int q = 0; q < kPieces; ++q) ldg_cols16(rbuf[0][q], rrow + wcol0 + q * 16, ncols_at(wcol0 + q * 16), wide_res);
      }
      const uint32_t aux_u = smem_u32(smem + L::kAuxOffset + acc * L::kAuxBytesPerStage);
      const uint32_t bias_u = aux_u + L::kAuxBias + colgrp * kColsPerWarp * 4;
      const uint32_t csum_u = aux_u + L::kAuxColsum + colgrp * kColsPerWarp * 4;
      mbar_wait(&aux_full[acc], acc_phase);
      float nmean = 0.f, rstd = 1.f;
      if constexpr (kLnCapable) {
        if (ln) {
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(nmean), "=f"(rstd)
                       : "r"(aux_u + L::kAuxRows + (quarter * 32 + lane) * 8));
        }
      }
      float st1 = 0.f, st2 = 0.f;   // partial (sum, sumsq) of this row's output columns

      mbar_wait(&tmem_full[acc], acc_phase);
      SF_GTL(const unsigned long long tl_acc = gtimer();)
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN +
                              colgrp * kColsPerWarp;
      uint32_t raw[2][kCW];
      auto tmem_fetch = [&](int c, uint32_t (&dst)[kCW]) {
        if constexpr (kCW == 32) tmem_ld_32x32b_x32(t_base + c * kCW, dst);
        else tmem_ld_32x32b_x16(t_base + c * kCW, dst);
      };
      tmem_fetch(0, raw[0]);
      // pairs of chunks (the two register buffers); the pair loop itself is not unrolled to keep the
      // epilogue's code inside the instruction cache
#pragma unroll 1
      for (int cp = 0; cp < kChunks / 2; ++cp) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = cp * 2 + h;
          const int col0 = wcol0 + c * kCW;
          tmem_ld_wait();                                     // chunk c is in registers
          if (c + 1 < kChunks) {
            tmem_fetch(c + 1, raw[h ^ 1]);
            if constexpr (EPI == kEpiResidual) {
#pragma unroll
              for (int q = 0; q < kPieces; ++q)
                ldg_cols16(rbuf[h ^ 1][q], rrow + col0 + kCW + q * 16, ncols_at(col0 + kCW + q * 16), wide_res);
            }
          }
          uint32_t ob[kPieces][8];
#pragma unroll
          for (int g = 0; g < kCW / 8; ++g) {  // groups of 8 columns
            Epi8 x8;                                 // packed pairs: see f32x2 above
            x8.load(&raw[h][g * 8]);
            const float4 b0 = lds128f(bias_u + (c * kCW + g * 8) * 4);
            const float4 b1 = lds128f(bias_u + (c * kCW + g * 8 + 4) * 4);
            bool did_ln = false;
            if constexpr (kLnCapable) {
              if (ln) {
                const float4 c0 = lds128f(csum_u + (c * kCW + g * 8) * 4);
                const float4 c1 = lds128f(csum_u + (c * kCW + g * 8 + 4) * 4);
                x8.ln_fold(nmean, rstd, c0, c1, b0, b1);
                did_ln = true;
              }
            }
            if (!did_ln) x8.add_bias(b0, b1);
            if constexpr (EPI == kEpiAct) {
              if (e.act == kActGeluErf) x8.gelu_erf_();
              else x8.gelu_tanh_();
            }
            if constexpr (EPI == kEpiEmbed) {
              const int col = col0 + g * 8;
              if (pos_row && row_ok && col < p.N)
                x8.add8(__ldg(reinterpret_cast<const float4*>(pos_row + col)), __ldg(reinterpret_cast<const float4*>(pos_row + col + 4)));
              if (time_row && row_ok && col < p.N)
                x8.add8(__ldg(reinterpret_cast<const float4*>(time_row + col)), __ldg(reinterpret_cast<const float4*>(time_row + col + 4)));
            }
            if constexpr (EPI == kEpiResidual) {
              x8.template residual<T>(gscale, &rbuf[h][g >> 1][(g & 1) * 4]);
            } else if constexpr (EPI == kEpiBias) {
              if (e.gate) x8.scale(gscale);
            }
            uint32_t* oq = &ob[g >> 1][(g & 1) * 4];
            x8.template pack<T>(oq);
            if constexpr (kStatsCapable) {
              if (want_stats) {   // statistics of the values as stored (rounded), what the next GEMM multiplies
                const float2 q0 = Pack2<T>::unpack(oq[0]), q1 = Pack2<T>::unpack(oq[1]);
                const float2 q2 = Pack2<T>::unpack(oq[2]), q3 = Pack2<T>::unpack(oq[3]);
                st1 += ((q0.x + q0.y) + (q1.x + q1.y)) + ((q2.x + q2.y) + (q3.x + q3.y));
                st2 = fmaf(q0.x, q0.x, fmaf(q0.y, q0.y, fmaf(q1.x, q1.x, fmaf(q1.y, q1.y, st2))));
                st2 = fmaf(q2.x, q2.x, fmaf(q2.y, q2.y, fmaf(q3.x, q3.x, fmaf(q3.y, q3.y, st2))));
              }
            }
          }
#pragma unroll
          for (int q = 0; q < kPieces; ++q) stg_cols16(orow + col0 + q * 16, ob[q], ncols_at(col0 + q * 16), wide_out);
        }
      }
      if constexpr (kStatsCapable) {
        if (want_stats && row_ok)
          e.stats_out[static_cast<long>(wcol0 / kColsPerWarp) * p.M + r] = make_float2(st1, st2);
      }
      // all TMEM reads of this accumulator and all reads of the aux stage are complete -> hand both back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_remote(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
        mbar_arrive(&aux_empty[acc]);
      }
      SF_GTL(if (it == 0 && blockIdx.x == 0 && warp == 4 && lane == 0 && p.M < 1024) printf("gemm N=%d K=%d epi=%d: acc ready +%llu epilogue done +%llu ns\n", p.N, p.K, EPI, tl_acc - tl_entry, gtimer() - tl_entry);)
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
// 2D K-major operand map: dims {K, rows}, box {64, box_rows}, 128B swizzle, zero OOB fill.
int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

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

// Slot path: how many consecutive N tiles one work item covers (G) and how many A K blocks stay
// resident for the item (R).  The main loop of a 256 x 256 CTA-pair tile is bound by operand ingest
// (~46 B per SM clock from L2, profiles/r1_ncu_layer.md) before it is bound by the tensor pipe
// (512 clk per 64-wide K block), so re-using A across the N tiles of an item shortens every tile
// after the first; larger groups leave fewer items to balance over the workers.  Pick the G that
// minimises rounds x (first tile + (G-1) later tiles).
void pick_group(int m_tiles, int n_tiles, int k_blocks, int workers, int max_res, int* G_out, int* R_out) {
  // measured on B200: grouping + resident A is neutral (the K=768 GEMMs were epilogue-bound, not
  // ingest-bound), so it is opt-in: SF_GEMM_GROUP=g forces groups of g N tiles, 0 = cost model
  static const int forced_g = env_int("SF_GEMM_GROUP", 1);
  static const int forced_r = env_int("SF_GEMM_RES", -1);
  int rmax = forced_r >= 0 ? forced_r : 8;
  if (rmax > max_res) rmax = max_res;
  if (rmax > k_blocks) rmax = k_blocks;
  const double mma = 512.0 * k_blocks, rate = 46.0;
  const double first = 2.0 * k_blocks * 16384.0 / rate;
  const double later = (2.0 * k_blocks - rmax) * 16384.0 / rate;
  double best = 1e30;
  int bg = 1;
  for (int g = 1; g <= 6 && g <= n_tiles; ++g) {
    if (forced_g > 0 && g != (forced_g < n_tiles ? forced_g : n_tiles)) continue;
    if (g > 1 && rmax == 0) break;
    const long items = static_cast<long>(m_tiles) * ((n_tiles + g - 1) / g);
    const long rounds = (items + workers - 1) / workers;
    const double cost = rounds * ((first > mma ? first : mma) + (g - 1) * (later > mma ? later : mma));
    if (cost < best * 0.999) { best = cost; bg = g; }
  }
  *G_out = bg;
  *R_out = bg > 1 ? rmax : 0;
}

// 2D map of an output / residual matrix for the TMA-store epilogue: dims {cols, rows}, box
// {box_cols, 32}, swizzle matching the staged row size (64 or 32 bytes).
int make_io_map(CUtensorMap* map, int dtype, const void* base, int rows, int cols, int ld, int box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -3;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), 32};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt =
      dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  box_cols * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(io map) failed (%d): rows=%d cols=%d ld=%d base=%p", (int)r, rows, cols, ld, base);
    return -3;
  }
  return 0;
}

template <typename T, int BN, int CG, int EPI, int EW, bool TS>
int launch_gemm(cudaStream_t stream, int dtype, const void* A, int lda, const void* W, int ldw,
                const GemmParams& p_in) {
  using L = SmemLayout<BN, CG, EW, TS>;
  GemmParams p = p_in;
  CUtensorMap tmA, tmB, tmOut, tmRes;
  int rc = make_operand_map(&tmA, dtype, A, p.M, p.K, lda, kBM);
  if (rc) return rc;
  rc = make_operand_map(&tmB, dtype, W, p.N, p.K, ldw, BN / CG);
  if (rc) return rc;
  if constexpr (TS) {
    rc = make_io_map(&tmOut, dtype, p.out, p.M, p.N, p.ldo, L::kChunkCols);
    if (rc) return rc;
    if (p.epi.residual) {
      rc = make_io_map(&tmRes, dtype, p.epi.residual, p.M, p.N, p.epi.ldr, L::kChunkCols);
      if (rc) return rc;
    } else {
      tmRes = tmOut;
    }
  } else {
    tmOut = tmA;   // unused by the kernel
    tmRes = tmA;
  }
  auto kernel = gemm_tcgen05_kernel<T, BN, CG, EPI, EW, TS>;
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
  const int max_workers = num_sms() / CG;
  p.group = 1;
  p.res_kb = 0;
  if constexpr (L::kSlotPath) pick_group(m_tiles, n_tiles, (p.K + kBK - 1) / kBK, max_workers, L::kMaxRes, &p.group, &p.res_kb);
  const int items = m_tiles * ((n_tiles + p.group - 1) / p.group);
  const int workers = items < max_workers ? items : max_workers;
  LaunchCfg lc(dim3(static_cast<unsigned>(workers * CG)), dim3(128 + EW * 32), L::kTotal, stream, CG);
  cudaError_t e;
  {
    ProfScope ps(stream, kProfGemm, 2.0 * p.M * p.N * p.K,
                 2.0 * (static_cast<double>(p.M) * p.K + static_cast<double>(p.N) * p.K +
                        static_cast<double>(p.M) * p.N * (p.epi.residual ? 2 : 1)));
    e = cudaLaunchKernelEx(&lc.cfg, kernel, tmA, tmB, tmOut, tmRes, p);
  }
  count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("gemm launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

// Tile shape for an M x N output: 0 -> 256 x 256 on CTA pairs (cta_group::2: half the B traffic per
// CTA) when that fills the machine, 1 -> 128 x 256 single-CTA tiles, 2 -> 128 x 128 for small
// problems (more CTAs in flight), 3 -> 128 x 64 for problems that leave even those tiles on under half the SMs.  SF_GEMM_MODE=1 forces single-CTA tiles (debug / A-B comparison).
int pick_shape(int M, int N) {
  static const int mode = env_int("SF_GEMM_MODE", 0);
  const int m_tiles = (M + kBM - 1) / kBM;
  const int m_tiles2 = (M + 2 * kBM - 1) / (2 * kBM);
  const int n_tiles = (N + 255) / 256;
  if (mode != 1 && N >= 256 && m_tiles2 * n_tiles >= num_sms() / 2) return 0;
  if (N >= 256 && m_tiles * n_tiles >= num_sms()) return 1;
  // 3 -> 128 x 64 tiles (four epilogue warps): streaming steps (M = 784) leave an N = 768 GEMM with 42 tiles of
  // 128 x 128 on 148 SMs, each a load -> MMA -> epilogue chain of ~6 us (12 us at K = 3072); half-width tiles double the
  // CTAs and halve the weight bytes, the MMA time and the epilogue of each.  SF_GEMM_BN64=0 disables.
  static const bool bn64 = env_int("SF_GEMM_BN64", 1) != 0;
  if (bn64 && m_tiles * ((N + 127) / 128) < num_sms() / 2 && m_tiles * ((N + 63) / 64) <= num_sms()) return 3;
  return 2;
}
// Epilogue warps: 16 (four per scheduler, 16-column chunks) for the GELU epilogue, whose MUFU/FMA
// latency two warps per scheduler cannot hide (fc1: 114 -> 102 us); 8 otherwise (measured equal or better).
// SF_GEMM_EW=8|16 forces one choice (A-B comparison).
int pick_epi_warps(int epi_mode, const GemmEpilogue& e) {
  static const int forced = env_int("SF_GEMM_EW", 0);
  if (forced == 8 || forced == 16) return forced;
  (void)e;
  return epi_mode == kEpiAct ? 16 : 8;
}

template <typename T, int EPI>
int dispatch_shape(cudaStream_t stream, int dtype, const void* A, int lda, const void* W, int ldw,
                   const GemmParams& p) {
  const int shape = pick_shape(p.M, p.N);
  // TMA-store epilogue: CTA-pair tiles writing rows in GEMM order (SF_GEMM_TS=0: per-thread stores)
  static const bool ts_on = env_int("SF_GEMM_TS", 1) != 0;
  if constexpr (EPI != kEpiEmbed) {
    if (shape == 0 && ts_on && p.epi.row_map == kRowIdentity) {
      if (pick_epi_warps(EPI, p.epi) == 16) return launch_gemm<T, 256, 2, EPI, 16, true>(stream, dtype, A, lda, W, ldw, p);
      return launch_gemm<T, 256, 2, EPI, 8, true>(stream, dtype, A, lda, W, ldw, p);
    }
  }
  if constexpr (EPI != kEpiAct) {
    if (shape == 3 && pick_epi_warps(EPI, p.epi) == 8) return launch_gemm<T, 64, 1, EPI, 4, false>(stream, dtype, A, lda, W, ldw, p);
  }
  if (pick_epi_warps(EPI, p.epi) == 16) {
    switch (shape) {
      case 0: return launch_gemm<T, 256, 2, EPI, 16, false>(stream, dtype, A, lda, W, ldw, p);
      case 1: return launch_gemm<T, 256, 1, EPI, 16, false>(stream, dtype, A, lda, W, ldw, p);
      default: return launch_gemm<T, 128, 1, EPI, 16, false>(stream, dtype, A, lda, W, ldw, p);
    }
  }
  switch (shape) {
    case 0: return launch_gemm<T, 256, 2, EPI, 8, false>(stream, dtype, A, lda, W, ldw, p);
    case 1: return launch_gemm<T, 256, 1, EPI, 8, false>(stream, dtype, A, lda, W, ldw, p);
    default: return launch_gemm<T, 128, 1, EPI, 8, false>(stream, dtype, A, lda, W, ldw, p);
  }
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


// =====================================================================================  GEMM chains
// Several dependent GEMMs over the SAME rows (attention out-proj -> fc1 -> fc2 -> the next layer's
// QKV, ...) as ONE persistent launch.  Each stand-alone GEMM pays ~12 us of pipeline fill (first
// operand fetch + a whole main loop before any epilogue work exists) and drain (the last tile's
// epilogue with idle tensor cores) — 18 % of a K = 768 GEMM, six times per layer.  In a chain the
// tiles of all phases form one sequence dealt round-robin to the CTA pairs: the drain of phase p
// overlaps the fill of phase p+1.  Dependencies are row-local — tile (p, m, n) needs every tile
// (p-1, m, *) — and are tracked with one counter per (phase, M tile) in global memory: epilogue
// warps release-increment it once their TMA stores have completed, the TMA producer / aux warp /
// residual prefetch of a dependent tile acquire-spin on it (normally already satisfied: the
// producing tiles are ~num_workers positions earlier in the sequence).  All CTAs are co-resident
// (persistent grid <= SM count) and every worker walks the sequence in order, so the earliest
// unfinished tile can always run: no deadlock.  The last CTA to leave zeroes the counters again, so
// every launch starts from a clean set without host-side state (safe under CUDA-graph replay).

struct ChainMaps {
  CUtensorMap a[kChainMaxPhases], b[kChainMaxPhases], out[kChainMaxPhases], res[kChainMaxPhases];
};
struct ChainPhase {
  int N, K, n_tiles, k_blocks, tile_begin;
  int mode;            // kEpiBias (bias / folded LN), kEpiAct, kEpiResidual
  int signal;          // a later phase consumes this one's rows
  uint32_t expected;   // arrivals per (phase, M tile) counter and launch: n_tiles * 2 CTAs * epilogue warps
  void* out;
  int ldo;
  GemmEpilogue epi;
};
struct ChainParams {
  int M, m_tiles, num_phases, total_tiles;
  uint32_t* done;      // [kChainMaxPhases][m_tiles] arrival counters + one CTA-exit counter, all zero between launches
  ChainPhase ph[kChainMaxPhases];
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// position of one worker in the tile sequence of a chain
struct ChainWalk {
  int g, step, p;
  __device__ ChainWalk(int worker, int workers) : g(worker), step(workers), p(0) {}
  __device__ bool valid(const ChainParams& c) const { return g < c.total_tiles; }
  __device__ void locate(const ChainParams& c) {
    while (p + 1 < c.num_phases && g >= c.ph[p + 1].tile_begin) ++p;
  }
  __device__ int m_tile(const ChainParams& c) const { return (g - c.ph[p].tile_begin) / c.ph[p].n_tiles; }
  __device__ int n_tile(const ChainParams& c) const { return (g - c.ph[p].tile_begin) % c.ph[p].n_tiles; }
  __device__ void next(const ChainParams& c) { g += step; locate(c); }
};

// rows of M tile `m_tile` written by phase p-1 are complete and visible (to generic and async proxy)
__device__ __forceinline__ void chain_wait(const ChainParams& c, int p, int m_tile) {
  if (p == 0) return;
  const uint32_t* ctr = c.done + (p - 1) * c.m_tiles + m_tile;
  const uint32_t target = c.ph[p - 1].expected;
  while (ld_acquire_gpu(ctr) < target) __nanosleep(40);
  // no proxy fence: the rows were complete in L2 (bulk-group completion) before the producer's
  // release, and TMA / ld.global.cg reads are served from L2
}

template <typename T, int EW>
__global__ void __launch_bounds__(128 + EW * 32, 1)
gemm_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams cp) {
  constexpr int BN = 256, CG = 2;
  using L = SmemLayout<BN, CG, EW, true>;
  constexpr int kTileM = kBM * CG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + L::kSlots;
  uint64_t* tmem_full = empty_bar + L::kSlots;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* aux_full = tmem_empty + 2;
  uint64_t* aux_empty = aux_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_empty + 2);
  uint64_t* res_bar = aux_empty + 3;
  static_assert((2 * L::kSlots + 8 + 1 + 3 * EW) * 8 <= L::kBarBytes, "chain barrier area overflow");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int worker = blockIdx.x / CG;
  const int num_workers = gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    for (int q = 0; q < cp.num_phases; ++q) {
      tma_prefetch_desc(&maps.a[q]);
      tma_prefetch_desc(&maps.b[q]);
      tma_prefetch_desc(&maps.out[q]);
      if (cp.ph[q].mode == kEpiResidual) tma_prefetch_desc(&maps.res[q]);
    }
    for (int s = 0; s < L::kSlots; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 3 * EW; ++i) mbar_init(&res_bar[i], 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], EW * CG);
      mbar_init(&aux_full[a], 1);
      mbar_init(&aux_empty[a], EW);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_slot, 2 * BN);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (every CTA)
    if (elect_one_sync()) {
      int slot = 0;
      uint32_t phase = 0;
      auto ring_load = [&](const CUtensorMap* tm, int c0, int c1) {
        mbar_wait(&empty_bar[slot], phase ^ 1);
        if (leader) mbar_arrive_expect_tx(&full_bar[slot], 2 * L::kSlotBytes);
        tma_load_2d_cg2(smem + slot * L::kSlotBytes, tm, &full_bar[slot], c0, c1);
        if (++slot == L::kSlots) { slot = 0; phase ^= 1; }
      };
      // the dependency counter of the NEXT tile is requested while this tile's operands stream in, so
      // the (normally satisfied) check costs no round trip on the critical path
      uint32_t pre_val = 0;
      bool pre_have = false;
      for (ChainWalk w(worker, num_workers); w.valid(cp); w.next(cp)) {
        const int q = w.p;
        const int mt = w.m_tile(cp);
        const int m0 = mt * kTileM + static_cast<int>(cta_rank) * kBM;
        const int n0 = w.n_tile(cp) * BN + static_cast<int>(cta_rank) * (BN / CG);
        // the A rows come from the previous phase
        if (q > 0 && !(pre_have && pre_val >= cp.ph[q - 1].expected)) chain_wait(cp, q, mt);
        {
          ChainWalk nx = w;
          nx.next(cp);
          pre_have = false;
          if (nx.valid(cp) && nx.p > 0) {
            pre_val = ld_acquire_gpu(cp.done + (nx.p - 1) * cp.m_tiles + nx.m_tile(cp));
            pre_have = true;
          }
        }
        const int kbs = cp.ph[q].k_blocks;
        for (int kb = 0; kb < kbs; ++kb) {
          ring_load(&maps.a[q], kb * kBK, m0);
          ring_load(&maps.b[q], kb * kBK, n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    constexpr uint32_t idesc = umma_idesc_f16(kTileM, BN, UmmaFmt<T>::value);
    if (leader && elect_one_sync()) {
      int slot = 0;
      uint32_t phase = 0;
      int it = 0;
      for (ChainWalk w(worker, num_workers); w.valid(cp); w.next(cp), ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int kbs = cp.ph[w.p].k_blocks;
        for (int kb = 0; kb < kbs; ++kb) {
          mbar_wait(&full_bar[slot], phase);
          const int a_slot = slot;
          if (++slot == L::kSlots) { slot = 0; phase ^= 1; }
          mbar_wait(&full_bar[slot], phase);
          const int b_slot = slot;
          if (++slot == L::kSlots) { slot = 0; phase ^= 1; }
          tc_fence_after();
          const uint64_t da = umma_desc_sw128_kmajor(smem_u32(smem + a_slot * L::kSlotBytes));
          const uint64_t db = umma_desc_sw128_kmajor(smem_u32(smem + b_slot * L::kSlotBytes));
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            umma_f16_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_cg2_mc(&empty_bar[a_slot], 0x3);
          umma_commit_cg2_mc(&empty_bar[b_slot], 0x3);
        }
        umma_commit_cg2_mc(&tmem_full[acc], 0x3);
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ aux warp (one tile ahead)
    int it = 0;
    for (ChainWalk w(worker, num_workers); w.valid(cp); w.next(cp), ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const ChainPhase& P = cp.ph[w.p];
      const GemmEpilogue& e = P.epi;
      const bool ln = P.mode != kEpiResidual && e.ln_stats != nullptr;
      const int mt = w.m_tile(cp);
      const int m0 = mt * kTileM + static_cast<int>(cta_rank) * kBM;
      const int n0 = w.n_tile(cp) * BN;
      const float inv_k = 1.0f / static_cast<float>(P.K);
      constexpr int kVec = BN / 4 / 32;
      float4 b4[kVec], c4[kVec];
#pragma unroll
      for (int v = 0; v < kVec; ++v) {
        const int col = n0 + (v * 32 + lane) * 4;
        b4[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        c4[v] = b4[v];
        if (col < P.N) {
          if (e.bias) b4[v] = __ldg(reinterpret_cast<const float4*>(e.bias + col));
          if (ln) c4[v] = __ldg(reinterpret_cast<const float4*>(e.ln_colsum + col));
        }
      }
      float s1x[kBM / 32], s2x[kBM / 32];
      if (ln) {
        chain_wait(cp, w.p, mt);                     // the row statistics come from the previous phase
        // two row groups per round trip (the statistics were written inside this launch: coherent loads)
        constexpr int kParts = 12;                   // partials per batch (N = 768 producers: 12 x 64 columns)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float a1[2] = {0.f, 0.f}, a2[2] = {0.f, 0.f};
          for (int q = 0; q < e.ln_parts; q += kParts) {
            float2 t[2][kParts];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const int m = m0 + (half * 2 + rr) * 32 + lane;
#pragma unroll
              for (int u = 0; u < kParts; ++u) {
                t[rr][u] = make_float2(0.f, 0.f);
                if (q + u < e.ln_parts && m < cp.M) t[rr][u] = __ldcg(e.ln_stats + static_cast<long>(q + u) * cp.M + m);
              }
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
              for (int u = 0; u < kParts; ++u) {
                a1[rr] += t[rr][u].x;
                a2[rr] += t[rr][u].y;
              }
            }
          }
          s1x[half * 2] = a1[0]; s2x[half * 2] = a2[0];
          s1x[half * 2 + 1] = a1[1]; s2x[half * 2 + 1] = a2[1];
        }
      }
      mbar_wait(&aux_empty[st], ph ^ 1);
      const uint32_t aux_u = smem_u32(smem + L::kAuxOffset + st * L::kAuxBytesPerStage);
#pragma unroll
      for (int v = 0; v < kVec; ++v) {
        sts128f(aux_u + L::kAuxBias + (v * 32 + lane) * 16, b4[v]);
        sts128f(aux_u + L::kAuxColsum + (v * 32 + lane) * 16, c4[v]);
      }
      if (ln) {
#pragma unroll
        for (int rr = 0; rr < kBM / 32; ++rr) {
          const float mean = s1x[rr] * inv_k;
          const float var = fmaxf(fmaf(s2x[rr], inv_k, -mean * mean), 0.f);
          const float rstd = rsqrtf(var + e.ln_eps);
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(aux_u + L::kAuxRows + (rr * 32 + lane) * 8), "f"(-mean), "f"(rstd)
                       : "memory");
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&aux_full[st]);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue warps (TMA-store path)
    const int ew = warp - 4;
    const int quarter = warp & 3;
    const int colgrp = ew >> 2;
    constexpr int kColsPerWarp = BN / (EW / 4);     // 64 (16 warps) / 128 (8 warps)
    constexpr int kCW = L::kChunkCols;              // 16 / 32
    constexpr int kChunks = kColsPerWarp / kCW;     // 4
    constexpr int kRowBytes = kCW * 2;
    constexpr int kBufBytes = L::kStageBufBytes;
    constexpr int kC16 = kRowBytes / 16;
    uint8_t* stg = smem + L::kStagingOffset + ew * (3 * kBufBytes);
    const uint32_t stg_u = smem_u32(stg);
    uint64_t* rbar = res_bar + ew * 3;
    const uint32_t swz = (static_cast<uint32_t>(lane) / (128 / kRowBytes)) & (kC16 - 1);
    const uint32_t my_row_u = static_cast<uint32_t>(lane) * kRowBytes;
    uint32_t f = 0;          // chunks processed so far: staging buffer f % 3
    uint32_t rpar = 0;       // bit b: parity of the next residual arrival in buffer b
    // residual look-ahead (lane 0): the chunk two ahead of the one being processed
    ChainWalk wl(worker, num_workers);
    int lc = 0;
    uint32_t lf = 0;
    // Advances the look-ahead by one chunk (requesting its residual box if it belongs to a residual
    // phase).  With block == false it gives up (returns false) when the chunk's tile still waits for
    // an earlier phase: that phase may need THIS warp's current tile (short chains: the producing
    // tile can be exactly one round earlier on the same worker), so spinning here would deadlock.
    auto advance_res = [&](bool block) -> bool {
      if (wl.valid(cp)) {
        const ChainPhase& LP = cp.ph[wl.p];
        if (LP.mode == kEpiResidual) {
          const int mt = wl.m_tile(cp);
          if (lc == 0 && wl.p > 0) {
            const uint32_t* ctr = cp.done + (wl.p - 1) * cp.m_tiles + mt;
            if (!block && ld_acquire_gpu(ctr) < cp.ph[wl.p - 1].expected) return false;
            chain_wait(cp, wl.p, mt);
          }
          const int lrow = mt * kTileM + static_cast<int>(cta_rank) * kBM + quarter * 32;
          const int lcol = wl.n_tile(cp) * BN + colgrp * kColsPerWarp + lc * kCW;
          const uint32_t b = lf % 3;
          mbar_arrive_expect_tx(&rbar[b], kBufBytes);
          tma_load_2d(stg + b * kBufBytes, &maps.res[wl.p], &rbar[b], lcol, lrow);
        }
        if (++lc == kChunks) { lc = 0; wl.next(cp); }
      }
      ++lf;
      return true;
    };
    if (lane == 0) { while (lf < 2 && advance_res(false)) {} }
    // Completion of a tile is signalled lazily, while the NEXT tile's second chunk is processed (its
    // TMA stores have landed by then, so nothing stalls) — unless the next accumulator is not ready:
    // then the wait is free, and it might even be this very signal the next tile depends on.
    uint32_t* pend_ctr = nullptr;
    auto flush_signal = [&]() {
      if (pend_ctr != nullptr) {
        if (lane == 0) {
          tma_store_wait<0>();
          __threadfence();
          red_release_gpu_add(pend_ctr, 1u);
        }
        pend_ctr = nullptr;
      }
    };
    int it = 0;
    for (ChainWalk w(worker, num_workers); w.valid(cp); w.next(cp), ++it) {
      const ChainPhase& P = cp.ph[w.p];
      const GemmEpilogue& e = P.epi;
      const int mode = P.mode;
      const bool ln = mode != kEpiResidual && e.ln_stats != nullptr;
      const bool want_stats = mode == kEpiResidual && e.stats_out != nullptr;
      const float gscale = e.gate ? tanhf(__ldg(e.gate)) : 1.0f;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int mt = w.m_tile(cp);
      const int m0 = mt * kTileM + static_cast<int>(cta_rank) * kBM;
      const int n0 = w.n_tile(cp) * BN;
      const int row0 = m0 + quarter * 32;
      const int m = row0 + lane;
      const int wcol0 = n0 + colgrp * kColsPerWarp;
      const uint32_t aux_u = smem_u32(smem + L::kAuxOffset + acc * L::kAuxBytesPerStage);
      const uint32_t bias_u = aux_u + L::kAuxBias + colgrp * kColsPerWarp * 4;
      const uint32_t csum_u = aux_u + L::kAuxColsum + colgrp * kColsPerWarp * 4;
      if (pend_ctr != nullptr) {
        // about to block on this tile's aux stage / accumulator: if either is not there yet, publish
        // the previous tile first (the wait is free, and this tile may be waiting for that signal)
        const bool r = mbar_try_wait(&aux_full[acc], acc_phase) && mbar_try_wait(&tmem_full[acc], acc_phase);
        if (!__shfl_sync(0xffffffffu, r ? 1u : 0u, 0)) flush_signal();
      }
      mbar_wait(&aux_full[acc], acc_phase);
      float nmean = 0.f, rstd = 1.f;
      if (ln) {
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(nmean), "=f"(rstd)
                     : "r"(aux_u + L::kAuxRows + (quarter * 32 + lane) * 8));
      }
      float st1 = 0.f, st2 = 0.f;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + colgrp * kColsPerWarp;
      uint32_t raw[2][kCW];
      auto tmem_fetch = [&](int c, uint32_t (&dst)[kCW]) {
        if constexpr (kCW == 32) tmem_ld_32x32b_x32(t_base + c * kCW, dst);
        else tmem_ld_32x32b_x16(t_base + c * kCW, dst);
      };
      tmem_fetch(0, raw[0]);
#pragma unroll 1
      for (int cpair = 0; cpair < kChunks / 2; ++cpair) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = cpair * 2 + h;
          const int col0 = wcol0 + c * kCW;
          tmem_ld_wait();
          if (c + 1 < kChunks) {
            tmem_fetch(c + 1, raw[h ^ 1]);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&tmem_empty[acc], 0);
          }
          if (c == 1 && pend_ctr != nullptr) {
            // only this tile's first store may still be in flight: the previous tile's have landed
            if (lane == 0) {
              tma_store_wait<1>();
              __threadfence();
              red_release_gpu_add(pend_ctr, 1u);
            }
            pend_ctr = nullptr;
          }
          const uint32_t buf = f % 3;
          const uint32_t buf_u = stg_u + buf * kBufBytes;
          uint32_t rb[kCW / 2];
          if (mode == kEpiResidual) {
            // this tile's dependencies are satisfied (its accumulator exists): catch up if the
            // look-ahead had to hold back
            if (lane == 0) { while (lf <= f) advance_res(true); }
            mbar_wait(&rbar[buf], (rpar >> buf) & 1);
            rpar ^= 1u << buf;
#pragma unroll
            for (int q = 0; q < kC16; ++q) lds128u(buf_u + my_row_u + ((static_cast<uint32_t>(q) ^ swz) << 4), &rb[q * 4]);
          }
          uint32_t ob[kCW / 2];
#pragma unroll
          for (int g = 0; g < kCW / 8; ++g) {
            float vv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) vv[j] = __uint_as_float(raw[h][g * 8 + j]);
            const float4 b0 = lds128f(bias_u + (c * kCW + g * 8) * 4);
            const float4 b1 = lds128f(bias_u + (c * kCW + g * 8 + 4) * 4);
            if (ln) {
              const float4 c0 = lds128f(csum_u + (c * kCW + g * 8) * 4);
              const float4 c1 = lds128f(csum_u + (c * kCW + g * 8 + 4) * 4);
              vv[0] = fmaf(rstd, fmaf(nmean, c0.x, vv[0]), b0.x); vv[1] = fmaf(rstd, fmaf(nmean, c0.y, vv[1]), b0.y);
              vv[2] = fmaf(rstd, fmaf(nmean, c0.z, vv[2]), b0.z); vv[3] = fmaf(rstd, fmaf(nmean, c0.w, vv[3]), b0.w);
              vv[4] = fmaf(rstd, fmaf(nmean, c1.x, vv[4]), b1.x); vv[5] = fmaf(rstd, fmaf(nmean, c1.y, vv[5]), b1.y);
              vv[6] = fmaf(rstd, fmaf(nmean, c1.z, vv[6]), b1.z); vv[7] = fmaf(rstd, fmaf(nmean, c1.w, vv[7]), b1.w);
            } else {
              vv[0] += b0.x; vv[1] += b0.y; vv[2] += b0.z; vv[3] += b0.w;
              vv[4] += b1.x; vv[5] += b1.y; vv[6] += b1.z; vv[7] += b1.w;
            }
            if (mode == kEpiAct) {
              if (e.act == kActGeluErf) {
#pragma unroll
                for (int j = 0; j < 8; ++j) vv[j] = gelu_erf(vv[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) vv[j] = gelu_tanh(vv[j]);
              }
            } else if (mode == kEpiResidual) {
              const uint32_t* rq = &rb[g * 4];
              const float2 r0 = Pack2<T>::unpack(rq[0]), r1 = Pack2<T>::unpack(rq[1]);
              const float2 r2 = Pack2<T>::unpack(rq[2]), r3 = Pack2<T>::unpack(rq[3]);
              vv[0] = fmaf(gscale, vv[0], r0.x); vv[1] = fmaf(gscale, vv[1], r0.y);
              vv[2] = fmaf(gscale, vv[2], r1.x); vv[3] = fmaf(gscale, vv[3], r1.y);
              vv[4] = fmaf(gscale, vv[4], r2.x); vv[5] = fmaf(gscale, vv[5], r2.y);
              vv[6] = fmaf(gscale, vv[6], r3.x); vv[7] = fmaf(gscale, vv[7], r3.y);
            }
            uint32_t* oq = &ob[g * 4];
            oq[0] = Pack2<T>::pack(vv[0], vv[1]); oq[1] = Pack2<T>::pack(vv[2], vv[3]);
            oq[2] = Pack2<T>::pack(vv[4], vv[5]); oq[3] = Pack2<T>::pack(vv[6], vv[7]);
            if (want_stats) {
              const float2 q0 = Pack2<T>::unpack(oq[0]), q1 = Pack2<T>::unpack(oq[1]);
              const float2 q2 = Pack2<T>::unpack(oq[2]), q3 = Pack2<T>::unpack(oq[3]);
              st1 += ((q0.x + q0.y) + (q1.x + q1.y)) + ((q2.x + q2.y) + (q3.x + q3.y));
              st2 = fmaf(q0.x, q0.x, fmaf(q0.y, q0.y, fmaf(q1.x, q1.x, fmaf(q1.y, q1.y, st2))));
              st2 = fmaf(q2.x, q2.x, fmaf(q2.y, q2.y, fmaf(q3.x, q3.x, fmaf(q3.y, q3.y, st2))));
            }
          }
          // staging buffer hand-over: (residual in ->) output out
          if (lane == 0) tma_store_wait_read<2>();
          __syncwarp();
#pragma unroll
          for (int q = 0; q < kC16; ++q) sts128u(buf_u + my_row_u + ((static_cast<uint32_t>(q) ^ swz) << 4), &ob[q * 4]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (row0 < cp.M) tma_store_2d(&maps.out[w.p], stg + buf * kBufBytes, col0, row0);
            tma_store_commit();
            tma_store_wait_read<1>();
            while (lf < f + 3 && advance_res(false)) {}     // residual boxes of chunks f + 1, f + 2
          }
          ++f;
        }
      }
      if (want_stats && m < cp.M)
        e.stats_out[static_cast<long>(wcol0 / kColsPerWarp) * cp.M + m] = make_float2(st1, st2);
      __syncwarp();
      if (lane == 0) mbar_arrive(&aux_empty[acc]);
      // this warp's part of the tile is complete once its TMA stores have landed; the statistics
      // stores of all lanes are ordered before lane 0's release by the __syncwarp above
      if (P.signal) pend_ctr = cp.done + w.p * cp.m_tiles + mt;
    }
    flush_signal();
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 2 * BN);
  }
  // every wait of this CTA is behind it: count it out; the last CTA of the grid re-arms the counters
  if (threadIdx.x == 0) {
    uint32_t* exit_ctr = cp.done + kChainMaxPhases * cp.m_tiles;
    __threadfence();
    if (atomicAdd(exit_ctr, 1u) == gridDim.x - 1) {
      for (int i = 0; i < kChainMaxPhases * cp.m_tiles; ++i) cp.done[i] = 0u;
      *exit_ctr = 0u;
      __threadfence();
    }
  }
}

}  // namespace

int gemm_stats_parts(int M, int N) {
  // stats_out comes from the residual / embed epilogues, which run with 8 epilogue warps unless forced
  static const int forced = env_int("SF_GEMM_EW", 0);
  const int ew = forced == 16 ? 16 : 8;
  const int shape = pick_shape(M, N);
  // 128-wide tiles / 8 warps and 64-wide tiles / 4 warps both leave one partial per 64 columns
  const int cols_per_part = (shape == 3 && ew == 8) ? 64 : (shape >= 2 ? 128 : 256) / (ew / 4);
  return (N + cols_per_part - 1) / cols_per_part;
}

// Chains are opt-in (sf_set_option("gemm_chain", 1) or SF_GEMM_CHAIN=1): correct and tested, but on
// B200 the chained schedule currently measures SLOWER than one launch per GEMM (7.43 vs 6.52 ms per
// cfg2 step), so the default stays one launch per GEMM.
int g_chain_override = -1;
bool gemm_chain_supported(int dtype, const GemmCall* calls, int n) {
  static const bool env_on = env_int("SF_GEMM_CHAIN", 0) != 0;
  const bool on = g_chain_override >= 0 ? g_chain_override != 0 : env_on;
  if (!on || n < 2 || n > kChainMaxPhases || (dtype != kBF16 && dtype != kF16)) return false;
  for (int i = 0; i < n; ++i) {
    const GemmCall& c = calls[i];
    const GemmEpilogue& e = c.epi;
    if (c.M != calls[0].M || c.M <= 0 || c.N <= 0 || c.K <= 0) return false;
    if ((c.K % 8) || (c.N % 8) || (c.lda % 8) || (c.ldw % 8) || (c.ldo % 8) || (e.residual && (e.ldr % 8))) return false;
    if (pick_shape(c.M, c.N) != 0) return false;                       // CTA-pair 256 x 256 tiles only
    if (e.row_map != kRowIdentity || e.pos || e.time_emb) return false;
    if (e.residual && (e.act != kActNone || e.ln_stats)) return false;
    if (e.stats_out && !e.residual) return false;
    if (e.gate && !e.residual) return false;
  }
  return true;
}

// 16 epilogue warps (64 columns per row-statistics partial) when a phase has an activation (the
// GELU epilogue needs four warps per scheduler), else 8 (128 columns per partial)
int chain_epi_warps(const GemmCall* calls, int n) {
  for (int i = 0; i < n; ++i)
    if (calls[i].epi.act != kActNone) return 16;
  return 8;
}
void set_gemm_chain(int on) { g_chain_override = on; }
int gemm_chain_stats_parts(const GemmCall* calls, int n, int N) {
  const int cols = 256 / (chain_epi_warps(calls, n) / 4);
  return (N + cols - 1) / cols;
}

size_t gemm_chain_counter_bytes(int M) {
  return (static_cast<size_t>(kChainMaxPhases) * ((M + 2 * kBM - 1) / (2 * kBM)) + 1) * sizeof(uint32_t);
}

template <int EW>
int launch_chain(cudaStream_t stream, int dtype, const GemmCall* calls, int n, void* counters);

int gemm_chain(cudaStream_t stream, int dtype, const GemmCall* calls, int n, void* counters) {
  if (!gemm_chain_supported(dtype, calls, n)) { set_error("gemm_chain: unsupported chain"); return -1; }
  if (chain_epi_warps(calls, n) == 16) return launch_chain<16>(stream, dtype, calls, n, counters);
  return launch_chain<8>(stream, dtype, calls, n, counters);
}

template <int EW>
int launch_chain(cudaStream_t stream, int dtype, const GemmCall* calls, int n, void* counters) {
  using L = SmemLayout<256, 2, EW, true>;
  ChainMaps maps;
  ChainParams cp;
  memset(&cp, 0, sizeof(cp));
  cp.M = calls[0].M;
  cp.m_tiles = (cp.M + 2 * kBM - 1) / (2 * kBM);
  cp.num_phases = n;
  cp.done = static_cast<uint32_t*>(counters);
  int tiles = 0;
  double flops = 0.0, bytes = 0.0;
  for (int i = 0; i < n; ++i) {
    const GemmCall& c = calls[i];
    ChainPhase& P = cp.ph[i];
    P.N = c.N; P.K = c.K;
    P.n_tiles = (c.N + 255) / 256;
    P.k_blocks = (c.K + kBK - 1) / kBK;
    P.tile_begin = tiles;
    tiles += cp.m_tiles * P.n_tiles;
    P.mode = c.epi.residual ? kEpiResidual : (c.epi.act != kActNone ? kEpiAct : kEpiBias);
    P.expected = static_cast<uint32_t>(P.n_tiles) * 2u * EW;
    P.signal = i + 1 < n ? 1 : 0;
    P.out = c.out; P.ldo = c.ldo;
    P.epi = c.epi;
    int rc = make_operand_map(&maps.a[i], dtype, c.A, c.M, c.K, c.lda, kBM);
    if (rc) return rc;
    rc = make_operand_map(&maps.b[i], dtype, c.W, c.N, c.K, c.ldw, 128);
    if (rc) return rc;
    rc = make_io_map(&maps.out[i], dtype, c.out, c.M, c.N, c.ldo, L::kChunkCols);
    if (rc) return rc;
    if (c.epi.residual) {
      rc = make_io_map(&maps.res[i], dtype, c.epi.residual, c.M, c.N, c.epi.ldr, L::kChunkCols);
      if (rc) return rc;
    } else {
      maps.res[i] = maps.out[i];
    }
    flops += 2.0 * c.M * c.N * c.K;
    bytes += 2.0 * (static_cast<double>(c.M) * c.K + static_cast<double>(c.N) * c.K +
                    static_cast<double>(c.M) * c.N * (c.epi.residual ? 2 : 1));
  }
  for (int i = n; i < kChainMaxPhases; ++i) {
    maps.a[i] = maps.a[0]; maps.b[i] = maps.b[0]; maps.out[i] = maps.out[0]; maps.res[i] = maps.res[0];
    cp.ph[i].tile_begin = tiles;
  }
  cp.total_tiles = tiles;
  const int max_workers = num_sms() / 2;
  const int workers = tiles < max_workers ? tiles : max_workers;
  LaunchCfg lc(dim3(static_cast<unsigned>(workers * 2)), dim3(128 + EW * 32), L::kTotal, stream, 2);
  cudaError_t e;
  {
    ProfScope ps(stream, kProfGemm, flops, bytes);
    if (dtype == kBF16) {
      static bool attr = false;
      if (!attr) { cudaFuncSetAttribute(gemm_chain_kernel<__nv_bfloat16, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal); attr = true; }
      e = cudaLaunchKernelEx(&lc.cfg, gemm_chain_kernel<__nv_bfloat16, EW>, maps, cp);
    } else {
      static bool attr = false;
      if (!attr) { cudaFuncSetAttribute(gemm_chain_kernel<__half, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal); attr = true; }
      e = cudaLaunchKernelEx(&lc.cfg, gemm_chain_kernel<__half, EW>, maps, cp);
    }
  }
  count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("gemm_chain launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

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
  if (epi.ln_stats && (epi.residual || epi.pos || epi.time_emb || !epi.ln_colsum || epi.ln_parts <= 0)) {
    set_error("gemm: a folded LayerNorm needs ln_colsum/ln_parts and cannot be combined with residual/embed epilogues");
    return -1;
  }
  if (epi.stats_out && !(epi.residual || epi.pos || epi.time_emb)) {
    set_error("gemm: stats_out is produced by the residual / embed epilogues only");
    return -1;
  }
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.out = out; p.ldo = ldo; p.epi = epi;
  if (dtype == kBF16) return dispatch_epilogue<__nv_bfloat16>(stream, dtype, A, lda, W, ldw, p);
  return dispatch_epilogue<__half>(stream, dtype, A, lda, W, ldw, p);
}

}  // namespace sf
