// attention.cu — the attention cores of the encoder (head_dim = 64, fp32 softmax, flash-style
// online normalisation, bf16/fp16 operands on the warp-level tensor-core path):
//   * temporal attention over frames, causal (reference TimesformerCausalSelfAttention.forward,
//     models/modeling_timesformer_siglip.py:575-615; KV-cache twin
//     downstream/VideoQA/llava/model/multimodal_encoder/timesformer_encoder.py:491-560) or
//     bidirectional (TimesformerSelfAttention, :688-717, when enable_causal_temporal=False);
//     K/V come either from the fresh QKV projection or from the pre-allocated streaming cache.
//   * spatial attention inside each frame (reference :688-717, lora_forward :649-683).
//   * the probe-query pooling attention of the SigLIP head (reference :1141-1148).
// These contractions are 3 % of the encoder FLOPs and HBM/L2-bound stand-alone, so they read
// Q/K/V in place (128-byte head rows, no permute copies) and never materialise the score tensor.
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include <cuda.h>

#include "sf_kernels.h"
#include "sf_ptx.cuh"
#include "sf_tma.h"

namespace sf {
namespace {

constexpr int kHd = 64;                       // head dim
constexpr float kLog2e = 1.4426950408889634f;

// swizzled byte offset of 16-byte chunk `chunk` (0..7) of row `row` in a [rows][64] 2-byte tile
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// Q fragments (A operand, 16 rows x 64) straight from global memory. Row g -> a[ks][0], a[ks][2];
// row g+8 -> a[ks][1], a[ks][3].
template <typename T>
__device__ __forceinline__ void load_q_frags(uint32_t (&a)[4][4], const T* q0, const T* q1, bool ok0,
                                             bool ok1, int c) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = ok0 ? *reinterpret_cast<const uint32_t*>(q0 + ks * 16 + 2 * c) : 0u;
    a[ks][1] = ok1 ? *reinterpret_cast<const uint32_t*>(q1 + ks * 16 + 2 * c) : 0u;
    a[ks][2] = ok0 ? *reinterpret_cast<const uint32_t*>(q0 + ks * 16 + 8 + 2 * c) : 0u;
    a[ks][3] = ok1 ? *reinterpret_cast<const uint32_t*>(q1 + ks * 16 + 8 + 2 * c) : 0u;
  }
}

// S(16 x 16 keys) = Q . K^T for the 16 keys starting at smem row `krow0` of the K tile.
template <bool kBf16>
__device__ __forceinline__ void qk_16keys(float (&s0)[4], float (&s1)[4], const uint32_t (&a)[4][4],
                                          uint32_t k_smem, int krow0, int lane) {
  const int mi = lane >> 3, rin = lane & 7;
  const int row = krow0 + (mi >> 1) * 8 + rin;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t b[4];
    ldmatrix_x4(b, k_smem + tile_off(row, 2 * ks + (mi & 1)));
    mma_16816<kBf16>(s0, a[ks], b[0], b[1]);
    mma_16816<kBf16>(s1, a[ks], b[2], b[3]);
  }
}

// O(16 x 64) += P(16 x 16 keys) . V for the 16 keys starting at smem row `vrow0` of the V tile.
template <bool kBf16>
__device__ __forceinline__ void pv_16keys(float (&o)[8][4], const uint32_t (&pa)[4], uint32_t v_smem,
                                          int vrow0, int lane) {
  const int mi = lane >> 3, rin = lane & 7;
  const int row = vrow0 + (mi & 1) * 8 + rin;
#pragma unroll
  for (int dp = 0; dp < 4; ++dp) {
    uint32_t b[4];
    ldmatrix_x4_trans(b, v_smem + tile_off(row, 2 * dp + (mi >> 1)));
    mma_16816<kBf16>(o[2 * dp], pa, b[0], b[1]);
    mma_16816<kBf16>(o[2 * dp + 1], pa, b[2], b[3]);
  }
}

// One online-softmax step over a 16-key slab (two n-tiles s0, s1 already scaled+masked, log2 domain).
// Produces the packed P fragments and rescales the running state.
template <typename T>
__device__ __forceinline__ void softmax_step(float (&s0)[4], float (&s1)[4], float (&m_run)[2],
                                             float (&l_run)[2], float (&o)[8][4], uint32_t (&pa)[4]) {
  float mx0 = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1]));
  float mx1 = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
  const float mu0 = (mn0 == -INFINITY) ? 0.f : mn0;  // fully masked so far: avoid (-inf)-(-inf)
  const float mu1 = (mn1 == -INFINITY) ? 0.f : mn1;
  const float al0 = exp2f(m_run[0] - mu0), al1 = exp2f(m_run[1] - mu1);
  m_run[0] = mn0;
  m_run[1] = mn1;
  const float p00 = exp2f(s0[0] - mu0), p01 = exp2f(s0[1] - mu0);
  const float p02 = exp2f(s0[2] - mu1), p03 = exp2f(s0[3] - mu1);
  const float p10 = exp2f(s1[0] - mu0), p11 = exp2f(s1[1] - mu0);
  const float p12 = exp2f(s1[2] - mu1), p13 = exp2f(s1[3] - mu1);
  l_run[0] = l_run[0] * al0 + (p00 + p01 + p10 + p11);
  l_run[1] = l_run[1] * al1 + (p02 + p03 + p12 + p13);
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    o[d][0] *= al0; o[d][1] *= al0; o[d][2] *= al1; o[d][3] *= al1;
  }
  pa[0] = Pack2<T>::pack(p00, p01);  // row g,   keys 2c..2c+1
  pa[1] = Pack2<T>::pack(p02, p03);  // row g+8, keys 2c..
  pa[2] = Pack2<T>::pack(p10, p11);  // row g,   keys 8+2c..
  pa[3] = Pack2<T>::pack(p12, p13);  // row g+8, keys 8+2c..
}

template <typename T>
__device__ __forceinline__ void finalize_store(float (&o)[8][4], float (&l_run)[2], T* out0, T* out1,
                                               bool ok0, bool ok1, int c) {
  float l0 = l_run[0], l1 = l_run[1];
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    if (ok0) *reinterpret_cast<uint32_t*>(out0 + d * 8 + 2 * c) = Pack2<T>::pack(o[d][0] * i0, o[d][1] * i0);
    if (ok1) *reinterpret_cast<uint32_t*>(out1 + d * 8 + 2 * c) = Pack2<T>::pack(o[d][2] * i1, o[d][3] * i1);
  }
}

// ------------------------------------------------------------------------------- temporal
struct TemporalArgs {
  const void* q;  long q_ld;                         // q row (site*Tq+i) at q + row*q_ld + h*64
  const void* k;  const void* v;
  long kv_site_stride, kv_head_stride, kv_row_stride;  // elements
  void* out; long out_ld;
  int sites, heads, Tq, Tk, q_off, causal, qtiles;
  long tasks;
  float scale_log2;
  const int* seen_dev;   // streaming under a CUDA graph: q_off = *seen_dev, Tk = q_off + Tq
};

constexpr int kTWarps = 4;

template <typename T>
__global__ void __launch_bounds__(kTWarps * 32) temporal_attn_kernel(const TemporalArgs a_in) {
  constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
  __shared__ __align__(128) uint8_t smem[kTWarps][2][2][16 * 128];  // per warp: two stages of (K tile, V tile)
  griddep_wait();
  griddep_launch_dependents();
  TemporalArgs a = a_in;
  if (a.seen_dev) {
    a.q_off = *a.seen_dev;
    a.Tk = a.q_off + a.Tq;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;
  const long task = static_cast<long>(blockIdx.x) * kTWarps + warp;
  if (task >= a.tasks) return;
  // task order: head fastest, then q-tile, then site
  const int h = static_cast<int>(task % a.heads);
  const long rest = task / a.heads;
  const int qt = static_cast<int>(rest % a.qtiles);
  const long site = rest / a.qtiles;
  const int i0 = qt * 16;

  const T* qbase = reinterpret_cast<const T*>(a.q) + (site * a.Tq) * a.q_ld + h * kHd;
  const bool ok0 = (i0 + g) < a.Tq, ok1 = (i0 + g + 8) < a.Tq;
  uint32_t qa[4][4];
  load_q_frags<T>(qa, qbase + static_cast<long>(i0 + g) * a.q_ld, qbase + static_cast<long>(i0 + g + 8) * a.q_ld,
                  ok0, ok1, c);

  const T* kbase = reinterpret_cast<const T*>(a.k) + site * a.kv_site_stride + h * a.kv_head_stride;
  const T* vbase = reinterpret_cast<const T*>(a.v) + site * a.kv_site_stride + h * a.kv_head_stride;
  // 16-key slabs are double-buffered: slab i+1 is in flight (cp.async) while slab i is consumed
  auto issue_slab = [&](int kb0, int stage) {
    uint8_t* ks = smem[warp][stage][0];
    uint8_t* vs = smem[warp][stage][1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = lane + 32 * i;
      const int row = idx >> 3, ch = idx & 7;
      const bool ok = (kb0 + row) < a.Tk;
      const long roff = static_cast<long>(ok ? kb0 + row : 0) * a.kv_row_stride + ch * 8;
      cp_async_16(ks + tile_off(row, ch), kbase + roff, ok);
      cp_async_16(vs + tile_off(row, ch), vbase + roff, ok);
    }
    cp_async_commit();
  };

  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float o[8][4];
#pragma unroll
  for (int d = 0; d < 8; ++d) { o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f; }

  int last_q = i0 + 15;
  if (last_q > a.Tq - 1) last_q = a.Tq - 1;
  int kmax = a.causal ? (a.q_off + last_q + 1) : a.Tk;
  if (kmax > a.Tk) kmax = a.Tk;

  if (kmax > 0) issue_slab(0, 0);
  int stage = 0;
  for (int kb0 = 0; kb0 < kmax; kb0 += 16, stage ^= 1) {
    // K/V rows kb0..kb0+15 (zero-filled past Tk) are landing in `stage`; request the next slab first
    if (kb0 + 16 < kmax) {
      issue_slab(kb0 + 16, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    const uint32_t ks_u = smem_u32(smem[warp][stage][0]), vs_u = smem_u32(smem[warp][stage][1]);

    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    qk_16keys<kBf16>(s0, s1, qa, ks_u, 0, lane);
    // scale (log2 domain) + mask
    const int lim0 = a.causal ? (a.q_off + i0 + g) : (a.Tk - 1);
    const int lim1 = a.causal ? (a.q_off + i0 + g + 8) : (a.Tk - 1);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j0 = kb0 + 2 * c + e, j1 = kb0 + 8 + 2 * c + e;
      s0[e] = (j0 < a.Tk && j0 <= lim0) ? s0[e] * a.scale_log2 : -INFINITY;
      s0[2 + e] = (j0 < a.Tk && j0 <= lim1) ? s0[2 + e] * a.scale_log2 : -INFINITY;
      s1[e] = (j1 < a.Tk && j1 <= lim0) ? s1[e] * a.scale_log2 : -INFINITY;
      s1[2 + e] = (j1 < a.Tk && j1 <= lim1) ? s1[2 + e] * a.scale_log2 : -INFINITY;
    }
    uint32_t pa[4];
    softmax_step<T>(s0, s1, m_run, l_run, o, pa);
    pv_16keys<kBf16>(o, pa, vs_u, 0, lane);
    __syncwarp();  // tile is reused by the next slab
  }

  T* obase = reinterpret_cast<T*>(a.out) + (site * a.Tq) * a.out_ld + h * kHd;
  finalize_store<T>(o, l_run, obase + static_cast<long>(i0 + g) * a.out_ld,
                    obase + static_cast<long>(i0 + g + 8) * a.out_ld, ok0, ok1, c);
}

// Fast path for the batch forward (no cache, all T <= 16 frames of a site in one 16-row tile): a
// persistent warp walks (site, head) tasks with the NEXT task's Q, K and V tiles already in flight
// (cp.async, two stages per warp), takes its Q fragments from shared memory with ldmatrix and sends
// the output through the consumed Q tile so that it leaves as 16-byte stores.  The generic kernel
// above spends 44 % of its samples waiting for the one load round trip each short-lived CTA makes
// and keeps the LSU 59 % busy with 4-byte Q loads / output stores (profiles/r1_ncu_layer.md).
template <typename T>
__global__ void __launch_bounds__(kTWarps * 32, 4) temporal_attn_fast_kernel(const TemporalArgs a) {
  constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
  extern __shared__ __align__(128) uint8_t tsm[];   // [warp][stage][Q | K | V][16 x 128 B]
  griddep_wait();
  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;
  uint8_t* wbase = tsm + warp * (2 * 3 * 2048);
  const long stride = static_cast<long>(gridDim.x) * kTWarps;
  long task = static_cast<long>(blockIdx.x) * kTWarps + warp;
  const T* qg = reinterpret_cast<const T*>(a.q);
  const T* kg = reinterpret_cast<const T*>(a.k);
  const T* vg = reinterpret_cast<const T*>(a.v);
  auto issue = [&](long t, int stage) {
    const int h = static_cast<int>(t % a.heads);
    const long site = t / a.heads;
    const T* qb = qg + (site * a.Tq) * a.q_ld + h * kHd;
    const T* kb = kg + site * a.kv_site_stride + h * a.kv_head_stride;
    const T* vb = vg + site * a.kv_site_stride + h * a.kv_head_stride;
    uint8_t* st = wbase + stage * (3 * 2048);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = lane + 32 * i;
      const int row = idx >> 3, ch = idx & 7;
      const bool ok = row < a.Tq;
      const long r = ok ? row : 0;
      cp_async_16(st + tile_off(row, ch), qb + r * a.q_ld + ch * 8, ok);
      cp_async_16(st + 2048 + tile_off(row, ch), kb + r * a.kv_row_stride + ch * 8, ok);
      cp_async_16(st + 4096 + tile_off(row, ch), vb + r * a.kv_row_stride + ch * 8, ok);
    }
    cp_async_commit();
  };
  int stage = 0;
  if (task < a.tasks) issue(task, 0);
  for (; task < a.tasks; task += stride, stage ^= 1) {
    const long next = task + stride;
    if (next < a.tasks) {
      issue(next, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    uint8_t* st = wbase + stage * (3 * 2048);
    const uint32_t qs_u = smem_u32(st), ks_u = qs_u + 2048, vs_u = qs_u + 4096;
    uint32_t qa[4][4];
    {
      const int mi = lane >> 3;
      const int row = (lane & 7) + (mi & 1) * 8;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(qa[ks], qs_u + tile_off(row, 2 * ks + (mi >> 1)));
    }
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    qk_16keys<kBf16>(s0, s1, qa, ks_u, 0, lane);
    const int lim0 = a.causal ? g : (a.Tk - 1);
    const int lim1 = a.causal ? (g + 8) : (a.Tk - 1);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j0 = 2 * c + e, j1 = 8 + 2 * c + e;
      s0[e] = (j0 < a.Tk && j0 <= lim0) ? s0[e] * a.scale_log2 : -INFINITY;
      s0[2 + e] = (j0 < a.Tk && j0 <= lim1) ? s0[2 + e] * a.scale_log2 : -INFINITY;
      s1[e] = (j1 < a.Tk && j1 <= lim0) ? s1[e] * a.scale_log2 : -INFINITY;
      s1[2 + e] = (j1 < a.Tk && j1 <= lim1) ? s1[2 + e] * a.scale_log2 : -INFINITY;
    }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float o[8][4];
#pragma unroll
    for (int d = 0; d < 8; ++d) { o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f; }
    uint32_t pa[4];
    softmax_step<T>(s0, s1, m_run, l_run, o, pa);
    pv_16keys<kBf16>(o, pa, vs_u, 0, lane);
    float l0 = l_run[0], l1 = l_run[1];
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    __syncwarp();   // every lane has its Q fragments: the Q tile becomes the output staging tile
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      *reinterpret_cast<uint32_t*>(st + tile_off(g, d) + 4 * c) = Pack2<T>::pack(o[d][0] * i0, o[d][1] * i0);
      *reinterpret_cast<uint32_t*>(st + tile_off(g + 8, d) + 4 * c) = Pack2<T>::pack(o[d][2] * i1, o[d][3] * i1);
    }
    __syncwarp();
    const int h = static_cast<int>(task % a.heads);
    const long site = task / a.heads;
    T* obase = reinterpret_cast<T*>(a.out) + (site * a.Tq) * a.out_ld + h * kHd;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = lane + 32 * i;
      const int row = idx >> 3, ch = idx & 7;
      if (row < a.Tq) *reinterpret_cast<uint4*>(obase + static_cast<long>(row) * a.out_ld + ch * 8) = *reinterpret_cast<const uint4*>(st + tile_off(row, ch));
    }
    __syncwarp();   // the tile is refilled two iterations from now
  }
}

// K/V slices of the fresh QKV projection -> cache[site][head][pos0 + i][64]
// Streaming decode: ONE new frame per site attending to its cached history (BASELINE configs[2]: B=4, one
// frame per step; reference: the KV twin's cached attention, …timesformer_encoder.py:491-560).  A (site, head)
// history is one contiguous block of rows in the cache, staged into shared memory as 16-row slabs by TMA
// (cp.async.bulk.tensor, SWIZZLE_128B — the layout ldmatrix wants) completing on an mbarrier; every warp is
// persistent and keeps a ring of up to 8 tasks in flight (the ring depth adapts to the history length, 48 KB of
// staging per warp), and because the history does not depend on the QKV GEMM that precedes this kernel in the
// stream, the first ring fill is issued BEFORE griddepcontrol.wait and overlaps that GEMM's tail.  The new
// frame's key/value row comes straight from the QKV buffer, is dropped into the staged tile and appended to the
// cache by the same warp (kv_append fused); the arithmetic is the generic kernel's (mma.sync m16n8k16 over
// 16-key slabs with online softmax; one valid query row per tile — the work is ~0.03 GFLOP per launch).
// The per-task dependent chain (ldmatrix -> mma -> softmax -> mma) is latency-bound, so the CTA carries as
// many warps as the staging budget allows: 192 KB of rings split over 12 warps (caches of up to 64 frames: a
// whole history fits one 16 KB stage) or 8 warps (up to 96 frames); the next task's query / new-row words are
// fetched while the current task is computed.
constexpr int kDecStaging = 192 * 1024;
constexpr int kDecMaxStages = 8;
constexpr int kDecSlabBytes = 16 * 128;             // 16 rows x 64 two-byte elements
constexpr int kDecMaxRows = 96;                     // frames incl. the new one
constexpr int kDecMaxWarps = 12;
constexpr int kDecSmem = kDecStaging + kDecMaxWarps * kDecMaxStages * 8 + 1024;
inline int dec_warps(int Tcap) {
  static const int forced = [] { const char* e = getenv("SF_DEC_WARPS"); return e ? atoi(e) : 0; }();   // tuning knob
  if (forced >= 1 && forced <= kDecMaxWarps) return forced;
  return Tcap <= 64 ? 12 : 8;
}

template <typename T>
__global__ void __launch_bounds__(kDecMaxWarps * 32, 1)
temporal_decode_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                       const TemporalArgs a, T* __restrict__ kc, T* __restrict__ vc, int Tcap) {
  const int kDecWarps = blockDim.x >> 5;
  const int kDecRing = kDecStaging / kDecWarps;      // staging bytes per warp (16 KB or 24 KB)
  constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
  extern __shared__ uint8_t dsm_raw[];
  uint8_t* dsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsm_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;
  uint8_t* ring = dsm + warp * kDecRing;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dsm + kDecStaging) + warp * kDecMaxStages;
  if (lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < kDecMaxStages; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  // rows of a slab that no load covers must hold finite values (their probabilities are exactly 0)
  for (int i = lane; i < kDecRing / 16; i += 32) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  __syncwarp();
  // frames cached before this step (device counter under a captured graph: written by the PREVIOUS step's graph)
  const int seen = a.seen_dev ? *a.seen_dev : a.q_off;
  const long nwarps = static_cast<long>(gridDim.x) * kDecWarps;
  const long task0 = static_cast<long>(blockIdx.x) * kDecWarps + warp;
  const int slabs = (seen + 1 + 15) >> 4;            // slabs of the staged tile incl. the new row
  const int loaded = (seen + 15) >> 4;               // slabs that come from the cache
  const uint32_t region = static_cast<uint32_t>(slabs) * kDecSlabBytes;   // bytes of the K (or V) part of a stage
  int nst = kDecRing / static_cast<int>(2 * region);   // >= 1: the host picks the warp count from the cache capacity
  if (nst > kDecMaxStages) nst = kDecMaxStages;
  auto issue = [&](long task, int s) {
    if (loaded == 0) return;
    uint8_t* base = ring + static_cast<size_t>(s) * 2 * region;
    mbar_arrive_expect_tx(&bars[s], 2u * loaded * kDecSlabBytes);
    const int row0 = static_cast<int>(task * Tcap);
    for (int i = 0; i < loaded; ++i) {
      tma_load_2d(base + i * kDecSlabBytes, &tmK, &bars[s], 0, row0 + 16 * i);
      tma_load_2d(base + region + i * kDecSlabBytes, &tmV, &bars[s], 0, row0 + 16 * i);
    }
  };
  if (lane == 0) {
    for (int s = 0; s < nst; ++s) {
      const long t = task0 + s * nwarps;
      if (t < a.tasks) issue(t, s);
    }
  }
  griddep_wait();               // from here on: q / k_new / v_new written by the QKV GEMM
  griddep_launch_dependents();

  const int D = a.heads * kHd;
  const long task_stride = static_cast<long>(Tcap) * kHd;           // elements between (site, head) blocks
  int s = 0;
  uint32_t phase = 0;
  // the (site, head) row words of a task: query fragments (one valid query row: tile row 0), new key / value pair
  auto row_of = [&](long task) {
    const long site = task / a.heads;
    const int h = static_cast<int>(task - site * a.heads);
    return reinterpret_cast<const T*>(a.q) + site * a.q_ld + h * kHd;
  };
  uint32_t qa_n[4][4], k2_n = 0u, v2_n = 0u;
  if (task0 < a.tasks) {
    const T* row = row_of(task0);
    k2_n = *reinterpret_cast<const uint32_t*>(row + D + 2 * lane);
    v2_n = *reinterpret_cast<const uint32_t*>(row + 2 * D + 2 * lane);
    load_q_frags<T>(qa_n, row, row, g == 0, false, c);
  }
  for (long task = task0; task < a.tasks; task += nwarps) {
    const long site = task / a.heads;
    const int h = static_cast<int>(task - site * a.heads);
    const uint32_t k2 = k2_n, v2 = v2_n;
    uint32_t qa[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { qa[i][0] = qa_n[i][0]; qa[i][1] = qa_n[i][1]; qa[i][2] = qa_n[i][2]; qa[i][3] = qa_n[i][3]; }
    if (task + nwarps < a.tasks) {          // next task's words travel while this one is computed
      const T* row = row_of(task + nwarps);
      k2_n = *reinterpret_cast<const uint32_t*>(row + D + 2 * lane);
      v2_n = *reinterpret_cast<const uint32_t*>(row + 2 * D + 2 * lane);
      load_q_frags<T>(qa_n, row, row, g == 0, false, c);
    }
    // fused kv_append: cache[site][head][seen][:] = new row
    *reinterpret_cast<uint32_t*>(kc + task * task_stride + static_cast<long>(seen) * kHd + 2 * lane) = k2;
    *reinterpret_cast<uint32_t*>(vc + task * task_stride + static_cast<long>(seen) * kHd + 2 * lane) = v2;
    uint8_t* Ks = ring + static_cast<size_t>(s) * 2 * region;
    uint8_t* Vs = Ks + region;
    if (loaded) mbar_wait(&bars[s], phase);
    // the new row joins the staged tile at row `seen` (lane covers 4 bytes of chunk lane/4)
    *reinterpret_cast<uint32_t*>(Ks + tile_off(seen, lane >> 2) + (lane & 3) * 4) = k2;
    *reinterpret_cast<uint32_t*>(Vs + tile_off(seen, lane >> 2) + (lane & 3) * 4) = v2;
    __syncwarp();
    const uint32_t ks_u = smem_u32(Ks), vs_u = smem_u32(Vs);
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float o[8][4];
#pragma unroll
    for (int d = 0; d < 8; ++d) { o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f; }
    for (int kb0 = 0; kb0 <= seen; kb0 += 16) {
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      qk_16keys<kBf16>(s0, s1, qa, ks_u, kb0, lane);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j0 = kb0 + 2 * c + e, j1 = kb0 + 8 + 2 * c + e;
        s0[e] = j0 <= seen ? s0[e] * a.scale_log2 : -INFINITY;
        s0[2 + e] = j0 <= seen ? s0[2 + e] * a.scale_log2 : -INFINITY;
        s1[e] = j1 <= seen ? s1[e] * a.scale_log2 : -INFINITY;
        s1[2 + e] = j1 <= seen ? s1[2 + e] * a.scale_log2 : -INFINITY;
      }
      uint32_t pa[4];
      softmax_step<T>(s0, s1, m_run, l_run, o, pa);
      pv_16keys<kBf16>(o, pa, vs_u, kb0, lane);
    }
    T* orow = reinterpret_cast<T*>(a.out) + site * a.out_ld + h * kHd;
    finalize_store<T>(o, l_run, orow, orow, g == 0, false, c);
    __syncwarp();                 // every lane is done with this stage before it is refilled
    const long nxt = task + static_cast<long>(nst) * nwarps;
    if (lane == 0 && nxt < a.tasks) {
      fence_proxy_async_smem();   // the generic-proxy stores of the new row precede the async-proxy refill
      issue(nxt, s);
    }
    if (++s == nst) { s = 0; phase ^= 1; }
  }
}

// Streaming decode, register-direct form (round 2, the default): the same contract as temporal_decode_kernel
// above, but the history goes global -> registers with 16-byte loads and the arithmetic runs on the FMA pipe.
//
// Why: at 32+ cached frames a (site, head) task is 8.7 KB of K/V; the TMA-ring kernel above has room for ONE
// stage per warp there (16 KB of staging x 12 warps), so every warp alternates between waiting for its history
// and a latency-bound ldmatrix -> mma.sync -> softmax -> mma.sync chain: 30 us per layer at 34 frames = 2.7 TB/s
// of an 82 MB read (ncu launch list, profiles/r2_stream_step.md).  The arithmetic is 2 x 64 x T MACs per task —
// nothing a tensor core is needed for — so this kernel drops shared memory, barriers and mma altogether:
//   lane = (r = lane / 8, c = lane % 8): rows j = j0 + 4 i + r of a 32-row batch, 16-byte chunk c of the row;
//   one load instruction fetches 4 consecutive 128-byte rows, the 16 independent loads (K and V) of a batch are
//   in flight together, 16 warps per SM (a 16-row double-buffered variant measured slower: 26.3 vs 25.0 us);
//   scores: 8-lane dot products (3 shuffles per 4 rows), online softmax over batches, P stays fp32;
//   O: every lane accumulates its rows' p x v for its 8 dims, two shuffle steps fold the four row groups.
// The cache rows of a warp's first task are requested BEFORE griddepcontrol.wait (they do not depend on the QKV
// GEMM); the new frame's key / value come from the QKV buffer and are appended to the cache by the lanes that
// own row `seen`.
template <typename T>
__global__ void __launch_bounds__(256, 2)
temporal_decode_direct_kernel(const TemporalArgs a, T* __restrict__ kc, T* __restrict__ vc, int Tcap) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = lane >> 3, c = lane & 7;
  const int seen = a.seen_dev ? *a.seen_dev : a.q_off;     // frames cached before this step
  const long nwarps = static_cast<long>(gridDim.x) * (blockDim.x >> 5);
  const long task0 = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + warp;
  const long task_stride = static_cast<long>(Tcap) * kHd;
  const int D = a.heads * kHd;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

  uint4 kk[8], vv[8];
  // cache rows j0 + 4 i + r (< seen) of `task`
  auto load_batch = [&](long task, int j0) {
    const T* kb = kc + task * task_stride + c * 8;
    const T* vb = vc + task * task_stride + c * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = j0 + 4 * i + r;
      kk[i] = zero4;
      vv[i] = zero4;
      if (j < seen) {
        kk[i] = __ldcs(reinterpret_cast<const uint4*>(kb + static_cast<long>(j) * kHd));
        vv[i] = __ldcs(reinterpret_cast<const uint4*>(vb + static_cast<long>(j) * kHd));
      }
    }
  };
  if (task0 < a.tasks) load_batch(task0, 0);   // the history does not depend on the QKV GEMM
  griddep_wait();               // from here on: q / k_new / v_new written by the QKV GEMM
  griddep_launch_dependents();

  for (long task = task0; task < a.tasks; task += nwarps) {
    const long site = task / a.heads;
    const int h = static_cast<int>(task - site * a.heads);
    const T* row = reinterpret_cast<const T*>(a.q) + site * a.q_ld + h * kHd + c * 8;
    const uint4 q4 = *reinterpret_cast<const uint4*>(row);
    const uint4 knew = *reinterpret_cast<const uint4*>(row + D);
    const uint4 vnew = *reinterpret_cast<const uint4*>(row + 2 * D);
    if (task != task0) load_batch(task, 0);   // requested before anything waits on the row words above
    float qf[8];
    {
      const float2 q0 = Pack2<T>::unpack(q4.x), q1 = Pack2<T>::unpack(q4.y), q2 = Pack2<T>::unpack(q4.z), q3 = Pack2<T>::unpack(q4.w);
      qf[0] = q0.x * a.scale_log2; qf[1] = q0.y * a.scale_log2; qf[2] = q1.x * a.scale_log2; qf[3] = q1.y * a.scale_log2;
      qf[4] = q2.x * a.scale_log2; qf[5] = q2.y * a.scale_log2; qf[6] = q3.x * a.scale_log2; qf[7] = q3.y * a.scale_log2;
    }
    if (r == (seen & 3)) {        // fused kv_append: cache[site][head][seen][:] = new row
      *reinterpret_cast<uint4*>(kc + task * task_stride + static_cast<long>(seen) * kHd + c * 8) = knew;
      *reinterpret_cast<uint4*>(vc + task * task_stride + static_cast<long>(seen) * kHd + c * 8) = vnew;
    }
    float m_run = -INFINITY, l_part = 0.f;
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j0 = 0; j0 <= seen; j0 += 32) {
      if (j0 > 0) load_batch(task, j0);
      float sc[8];
      float bm = -INFINITY;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sc[i] = -INFINITY;
        if (j0 + 4 * i <= seen) {     // warp-uniform: the row group holds at least one valid key
          const int j = j0 + 4 * i + r;
          if (j == seen) { kk[i] = knew; vv[i] = vnew; }
          const float2 k0 = Pack2<T>::unpack(kk[i].x), k1 = Pack2<T>::unpack(kk[i].y), k2 = Pack2<T>::unpack(kk[i].z), k3 = Pack2<T>::unpack(kk[i].w);
          float d = qf[0] * k0.x;
          d = fmaf(qf[1], k0.y, d); d = fmaf(qf[2], k1.x, d); d = fmaf(qf[3], k1.y, d);
          d = fmaf(qf[4], k2.x, d); d = fmaf(qf[5], k2.y, d); d = fmaf(qf[6], k3.x, d); d = fmaf(qf[7], k3.y, d);
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          if (j <= seen) sc[i] = d;
          bm = fmaxf(bm, sc[i]);
        }
      }
      bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 8));
      bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
      const float m_new = fmaxf(m_run, bm);       // finite: every batch holds at least one valid key
      const float resc = exp2f(m_run - m_new);    // 0 on the first batch
      l_part *= resc;
#pragma unroll
      for (int d = 0; d < 8; ++d) o[d] *= resc;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (j0 + 4 * i <= seen) {
          const float pj = exp2f(sc[i] - m_new);  // masked keys: exp2(-inf) = 0
          l_part += pj;
          const float2 v0 = Pack2<T>::unpack(vv[i].x), v1 = Pack2<T>::unpack(vv[i].y), v2 = Pack2<T>::unpack(vv[i].z), v3 = Pack2<T>::unpack(vv[i].w);
          o[0] = fmaf(pj, v0.x, o[0]); o[1] = fmaf(pj, v0.y, o[1]); o[2] = fmaf(pj, v1.x, o[2]); o[3] = fmaf(pj, v1.y, o[3]);
          o[4] = fmaf(pj, v2.x, o[4]); o[5] = fmaf(pj, v2.y, o[5]); o[6] = fmaf(pj, v3.x, o[6]); o[7] = fmaf(pj, v3.y, o[7]);
        }
      }
      m_run = m_new;
    }
    // fold the four row groups
    l_part += __shfl_xor_sync(0xffffffffu, l_part, 8);
    l_part += __shfl_xor_sync(0xffffffffu, l_part, 16);
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      o[d] += __shfl_xor_sync(0xffffffffu, o[d], 8);
      o[d] += __shfl_xor_sync(0xffffffffu, o[d], 16);
    }
    if (r == 0) {
      const float inv = 1.0f / l_part;
      uint4 w;
      w.x = Pack2<T>::pack(o[0] * inv, o[1] * inv); w.y = Pack2<T>::pack(o[2] * inv, o[3] * inv);
      w.z = Pack2<T>::pack(o[4] * inv, o[5] * inv); w.w = Pack2<T>::pack(o[6] * inv, o[7] * inv);
      *reinterpret_cast<uint4*>(reinterpret_cast<T*>(a.out) + site * a.out_ld + h * kHd + c * 8) = w;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
kv_append_kernel(const T* __restrict__ qkv, long ld, T* __restrict__ kc, T* __restrict__ vc, int Tcap,
                 int sites, int heads, int Tq, int pos0, const int* __restrict__ seen_dev) {
  griddep_wait();
  griddep_launch_dependents();
  if (seen_dev) pos0 = *seen_dev;
  const int D = heads * kHd;
  const long total = static_cast<long>(sites) * Tq * heads * 8;  // 16-byte chunks per K (and per V)
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i & 7);
    long r = i >> 3;
    const int h = static_cast<int>(r % heads);
    r /= heads;
    const int t = static_cast<int>(r % Tq);
    const long site = r / Tq;
    const T* src = qkv + (site * Tq + t) * ld + D + h * kHd + ch * 8;
    const long dst = ((site * heads + h) * Tcap + pos0 + t) * kHd + ch * 8;
    *reinterpret_cast<uint4*>(kc + dst) = *reinterpret_cast<const uint4*>(src);
    *reinterpret_cast<uint4*>(vc + dst) = *reinterpret_cast<const uint4*>(src + D);
  }
}

// ------------------------------------------------------------------------------- spatial
constexpr int kSWarps = 4;        // 64 query rows per CTA
constexpr int kSKeys = 64;        // keys per pipeline stage

struct SpatialArgs {
  const void* qkv; long ld;
  void* out; long out_ld;
  int frames, heads, S, qblocks;
  int T_inner;   // <= 1: frame rows contiguous; > 1: frame (b,t) lives at rows (b*S + n)*T_inner + t
  float scale_log2;
};

template <typename T>
__global__ void __launch_bounds__(kSWarps * 32) spatial_attn_kernel(const SpatialArgs a) {
  griddep_wait();
  griddep_launch_dependents();
  constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
  __shared__ __align__(128) uint8_t smem[2][2][kSKeys * 128];  // [stage][K|V]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;
  const int D = a.heads * kHd;
  int bid = blockIdx.x;
  const int qb = bid % a.qblocks;
  bid /= a.qblocks;
  const int h = bid % a.heads;
  const long frame = bid / a.heads;
  // first row of the frame and distance (in rows) between its consecutive tokens
  long row0 = frame * a.S, rstep = 1;
  if (a.T_inner > 1) {
    row0 = (frame / a.T_inner) * a.S * a.T_inner + frame % a.T_inner;
    rstep = a.T_inner;
  }
  const long ld = a.ld * rstep, old = a.out_ld * rstep;

  const T* base = reinterpret_cast<const T*>(a.qkv) + row0 * a.ld + h * kHd;
  const T* kbase = base + D;
  const T* vbase = base + 2 * D;
  const int i0 = qb * (kSWarps * 16) + warp * 16;
  const bool warp_active = i0 < a.S;
  const bool ok0 = (i0 + g) < a.S, ok1 = (i0 + g + 8) < a.S;

  uint32_t qa[4][4];
  load_q_frags<T>(qa, base + static_cast<long>(i0 + g) * ld, base + static_cast<long>(i0 + g + 8) * ld,
                  ok0, ok1, c);

  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float o[8][4];
#pragma unroll
  for (int d = 0; d < 8; ++d) { o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f; }

  const int nblocks = (a.S + kSKeys - 1) / kSKeys;
  auto prefetch = [&](int kb, int stage) {
    const int kb0 = kb * kSKeys;
#pragma unroll
    for (int i = 0; i < (kSKeys * 8) / (kSWarps * 32); ++i) {
      const int idx = threadIdx.x + i * (kSWarps * 32);
      const int row = idx >> 3, ch = idx & 7;
      const bool ok = (kb0 + row) < a.S;
      const long roff = static_cast<long>(ok ? kb0 + row : 0) * ld + ch * 8;
      cp_async_16(smem[stage][0] + tile_off(row, ch), kbase + roff, ok);
      cp_async_16(smem[stage][1] + tile_off(row, ch), vbase + roff, ok);
    }
    cp_async_commit();
  };

  prefetch(0, 0);
  for (int kb = 0; kb < nblocks; ++kb) {
    const int stage = kb & 1;
    if (kb + 1 < nblocks) {
      prefetch(kb + 1, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (warp_active) {
      const int kb0 = kb * kSKeys;
      int nsub = (a.S - kb0 + 15) >> 4;   // valid 16-key slabs in this block
      if (nsub > kSKeys / 16) nsub = kSKeys / 16;
      const uint32_t ks_u = smem_u32(smem[stage][0]), vs_u = smem_u32(smem[stage][1]);
#pragma unroll
      for (int sub = 0; sub < kSKeys / 16; ++sub) {
        if (sub < nsub) {
          float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
          qk_16keys<kBf16>(s0, s1, qa, ks_u, sub * 16, lane);
          const int kk = kb0 + sub * 16;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const bool v0 = (kk + 2 * c + e) < a.S, v1 = (kk + 8 + 2 * c + e) < a.S;
            s0[e] = v0 ? s0[e] * a.scale_log2 : -INFINITY;
            s0[2 + e] = v0 ? s0[2 + e] * a.scale_log2 : -INFINITY;
            s1[e] = v1 ? s1[e] * a.scale_log2 : -INFINITY;
            s1[2 + e] = v1 ? s1[2 + e] * a.scale_log2 : -INFINITY;
          }
          uint32_t pa[4];
          softmax_step<T>(s0, s1, m_run, l_run, o, pa);
          pv_16keys<kBf16>(o, pa, vs_u, sub * 16, lane);
        }
      }
    }
    __syncthreads();
  }
  if (warp_active) {
    T* obase = reinterpret_cast<T*>(a.out) + row0 * a.out_ld + h * kHd;
    finalize_store<T>(o, l_run, obase + static_cast<long>(i0 + g) * old,
                      obase + static_cast<long>(i0 + g + 8) * old, ok0, ok1, c);
  }
}

// Debug path (output_attentions=True): normalised probabilities, one warp per query row.
template <typename T>
__global__ void __launch_bounds__(128)
spatial_probs_kernel(const T* __restrict__ qkv, long ld_in, float* __restrict__ probs, int frames, int heads,
                     int S, int T_inner, float scale) {
  griddep_wait();
  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row_id = static_cast<long>(blockIdx.x) * 4 + warp;  // (frame, head, query)
  const long total = static_cast<long>(frames) * heads * S;
  if (row_id >= total) return;
  const int qi = static_cast<int>(row_id % S);
  const int h = static_cast<int>((row_id / S) % heads);
  const long frame = row_id / (static_cast<long>(S) * heads);
  const int D = heads * kHd;
  long row0 = frame * S, rstep = 1;
  if (T_inner > 1) {
    row0 = (frame / T_inner) * S * T_inner + frame % T_inner;
    rstep = T_inner;
  }
  const long ld = ld_in * rstep;
  const T* base = qkv + row0 * ld_in + h * kHd;
  const T* q = base + static_cast<long>(qi) * ld;
  float* prow = probs + row_id * S;
  float qf[2];
  {
    const float2 t = Pack2<T>::unpack(*reinterpret_cast<const uint32_t*>(q + 2 * lane));
    qf[0] = t.x; qf[1] = t.y;
  }
  float mx = -INFINITY;
  for (int j = 0; j < S; ++j) {
    const float2 kf = Pack2<T>::unpack(*reinterpret_cast<const uint32_t*>(base + D + static_cast<long>(j) * ld + 2 * lane));
    const float s = warp_sum(qf[0] * kf.x + qf[1] * kf.y) * scale;
    if (lane == 0) prow[j] = s;
    mx = fmaxf(mx, s);
  }
  __syncwarp();
  float sum = 0.f;
  for (int j = lane; j < S; j += 32) {
    const float e = __expf(prow[j] - mx);
    prow[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int j = lane; j < S; j += 32) prow[j] *= inv;
}

// ------------------------------------------------------------------------------- pooling probe
// One CTA per (frame, head): scores of the S keys against a fixed fp32 query, softmax, weighted V sum.  The four
// warps take a quarter of the keys each; lane = (r = lane / 8, c = lane % 8) reads 16-byte chunk c of rows
// n0 + 4 i + r, so one load instruction fetches four whole 128-byte head slices and eight of them are in flight per
// lane (round 1 walked the S value rows one dependent 4-byte load at a time: 28 us for the 48 tasks of a streaming
// step).  Scores go through shared memory, row maxima / sums / the four partial outputs meet there too.
constexpr int kPoolMaxS = 1024;
template <typename T>
__global__ void __launch_bounds__(128)
pool_attn_kernel(const T* __restrict__ kv, long ld, const float* __restrict__ q, T* __restrict__ out,
                 long out_ld, int frames, int heads, int S) {
  __shared__ float sc[kPoolMaxS];
  __shared__ float red[4][2];
  __shared__ float part[4][kHd];
  griddep_wait();
  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = lane >> 3, c = lane & 7;
  const long task = blockIdx.x;
  const int h = static_cast<int>(task % heads);
  const long frame = task / heads;
  const int D = heads * kHd;
  const T* kb = kv + frame * S * ld + h * kHd + c * 8;
  const T* vb = kb + D;
  float qf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) qf[j] = __ldg(q + h * kHd + c * 8 + j);
  const int per = (S + 3) >> 2;                 // keys per warp
  const int n_lo = warp * per, n_hi = (n_lo + per < S) ? n_lo + per : S;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

  // pass 1: scores of this warp's keys, their maximum
  float mx = -INFINITY;
  for (int n0 = n_lo; n0 < n_hi; n0 += 32) {
    uint4 kk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + 4 * i + r;
      kk[i] = n < n_hi ? *reinterpret_cast<const uint4*>(kb + static_cast<long>(n) * ld) : zero4;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + 4 * i + r;
      const float2 k0 = Pack2<T>::unpack(kk[i].x), k1 = Pack2<T>::unpack(kk[i].y), k2 = Pack2<T>::unpack(kk[i].z), k3 = Pack2<T>::unpack(kk[i].w);
      float d = qf[0] * k0.x;
      d = fmaf(qf[1], k0.y, d); d = fmaf(qf[2], k1.x, d); d = fmaf(qf[3], k1.y, d);
      d = fmaf(qf[4], k2.x, d); d = fmaf(qf[5], k2.y, d); d = fmaf(qf[6], k3.x, d); d = fmaf(qf[7], k3.y, d);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      if (n < n_hi) {
        if (c == 0) sc[n] = d;
        mx = fmaxf(mx, d);
      }
    }
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp][0] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0][0], red[1][0]), fmaxf(red[2][0], red[3][0]));

  // pass 2: probabilities and the weighted value sum of this warp's keys
  float sum = 0.f;
  float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int n0 = n_lo; n0 < n_hi; n0 += 32) {
    uint4 vv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + 4 * i + r;
      vv[i] = n < n_hi ? *reinterpret_cast<const uint4*>(vb + static_cast<long>(n) * ld) : zero4;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + 4 * i + r;
      const float pj = n < n_hi ? __expf(sc[n] - mx) : 0.f;
      sum += pj;
      const float2 v0 = Pack2<T>::unpack(vv[i].x), v1 = Pack2<T>::unpack(vv[i].y), v2 = Pack2<T>::unpack(vv[i].z), v3 = Pack2<T>::unpack(vv[i].w);
      o[0] = fmaf(pj, v0.x, o[0]); o[1] = fmaf(pj, v0.y, o[1]); o[2] = fmaf(pj, v1.x, o[2]); o[3] = fmaf(pj, v1.y, o[3]);
      o[4] = fmaf(pj, v2.x, o[4]); o[5] = fmaf(pj, v2.y, o[5]); o[6] = fmaf(pj, v3.x, o[6]); o[7] = fmaf(pj, v3.y, o[7]);
    }
  }
  // fold the four row groups of the warp (lanes of one group hold the same probabilities), then the four warps
  sum += __shfl_xor_sync(0xffffffffu, sum, 8);
  sum += __shfl_xor_sync(0xffffffffu, sum, 16);
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    o[d] += __shfl_xor_sync(0xffffffffu, o[d], 8);
    o[d] += __shfl_xor_sync(0xffffffffu, o[d], 16);
  }
  if (r == 0) {
#pragma unroll
    for (int d = 0; d < 8; ++d) part[warp][c * 8 + d] = o[d];
    if (c == 0) red[warp][1] = sum;
  }
  __syncthreads();
  if (warp == 0) {
    const float inv = 1.f / ((red[0][1] + red[1][1]) + (red[2][1] + red[3][1]));
    const float a0 = ((part[0][2 * lane] + part[1][2 * lane]) + (part[2][2 * lane] + part[3][2 * lane])) * inv;
    const float a1 = ((part[0][2 * lane + 1] + part[1][2 * lane + 1]) + (part[2][2 * lane + 1] + part[3][2 * lane + 1])) * inv;
    *reinterpret_cast<uint32_t*>(out + frame * out_ld + h * kHd + 2 * lane) = Pack2<T>::pack(a0, a1);
  }
}

// ------------------------------------------------------------------------------- pooling probe, collapsed
// The SigLIP pooling head attends with ONE learned probe (reference :1141-1148), so its key and value
// projections never have to be materialised: with q_h the (scaled) probe query of head h,
//   score_h(n) = q_h . (W_k,h x_n + b_k,h) = x_n . u_h + const,   u_h = W_k,h^T q_h   (const cancels in softmax)
//   out_h      = sum_n p_h(n) (W_v,h x_n + b_v,h) = W_v,h (sum_n p_h(n) x_n) + b_v,h
// i.e. 12 dot products per token, a probability-weighted token sum per head and one 64 x D mat-vec
// per head and frame — instead of a [tokens, D] x [D, 2D] GEMM (59 GFLOP and 77 MB at cfg2) plus an
// attention pass over its output.  One CTA per frame, one warp per head; tokens stream through shared
// memory in chunks twice (scores, then the weighted sum; the second pass hits L2).
constexpr int kProbeChunk = 16;     // tokens per shared-memory chunk (staged as fp32)
constexpr int kProbeMaxD = 1024;
template <typename T>
__global__ void __launch_bounds__(512)
pool_probe_kernel(const T* __restrict__ x, long ld, const float* __restrict__ u, const T* __restrict__ wv,
                  const float* __restrict__ bv, T* __restrict__ out, long out_ld, int heads, int S) {
  extern __shared__ __align__(16) uint8_t psm[];
  const int D = heads * kHd;
  const int Sp = (S + 3) & ~3;
  float* sx = reinterpret_cast<float*>(psm);                                       // [2][kProbeChunk][D] fp32
  float* sc = sx + 2 * kProbeChunk * D;                                            // [heads][Sp]
  float* sv = sc + heads * Sp;                                                     // [heads][D] pooled tokens
  griddep_wait();
  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // warp == head
  const long frame = blockIdx.x;
  const T* xf = x + frame * S * ld;
  const int groups = D / 128;                                   // float4 per lane and row (6 for D = 768)
  constexpr int kMaxGroups = kProbeMaxD / 128;
  // this lane's slice of u_h: elements g*128 + lane*4 .. +3
  float4 ur[kMaxGroups];
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g)
    ur[g] = g < groups ? __ldg(reinterpret_cast<const float4*>(u + warp * D + g * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);

  // staging: every thread moves `per_thread` 8-element pieces of a chunk (global 16-bit -> shared fp32);
  // the next chunk's pieces are fetched into registers before the current one is consumed
  const int chunks = (S + kProbeChunk - 1) / kProbeChunk;
  const int pieces_per_chunk = kProbeChunk * D / 8;
  constexpr int kMaxPer = 8;
  const int per_thread = (pieces_per_chunk + static_cast<int>(blockDim.x) - 1) / static_cast<int>(blockDim.x);   // 4 at D=768, 384 threads
  uint4 stage[kMaxPer];
  auto fetch = [&](int c) {
    const int n0 = c * kProbeChunk;
#pragma unroll
    for (int q = 0; q < kMaxPer; ++q) {
      stage[q] = make_uint4(0u, 0u, 0u, 0u);
      if (q < per_thread) {
        const int i = q * blockDim.x + threadIdx.x;
        const int r = i / (D / 8), piece = i % (D / 8);
        if (i < pieces_per_chunk && n0 + r < S) stage[q] = *reinterpret_cast<const uint4*>(xf + static_cast<long>(n0 + r) * ld + piece * 8);
      }
    }
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int q = 0; q < kMaxPer; ++q) {
      if (q < per_thread) {
        const int i = q * blockDim.x + threadIdx.x;
        if (i < pieces_per_chunk) {
          const float2 a = Pack2<T>::unpack(stage[q].x), b = Pack2<T>::unpack(stage[q].y);
          const float2 c2 = Pack2<T>::unpack(stage[q].z), d = Pack2<T>::unpack(stage[q].w);
          float4* dst = reinterpret_cast<float4*>(sx + buf * kProbeChunk * D + i * 8);
          dst[0] = make_float4(a.x, a.y, b.x, b.y);
          dst[1] = make_float4(c2.x, c2.y, d.x, d.y);
        }
      }
    }
  };

  // ---- pass 1: scores of this head for every token
  fetch(0);
  for (int c = 0; c < chunks; ++c) {
    const int buf = c & 1;
    commit(buf);
    if (c + 1 < chunks) fetch(c + 1);
    __syncthreads();                       // chunk c staged (buffer buf was last read two chunks ago)
    const float* xb = sx + buf * kProbeChunk * D;
    float part[kProbeChunk];
#pragma unroll
    for (int r = 0; r < kProbeChunk; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int g = 0; g < kMaxGroups; ++g) {
        if (g < groups) {
          const float4 xv = *reinterpret_cast<const float4*>(xb + r * D + g * 128 + lane * 4);
          acc = fmaf(xv.x, ur[g].x, fmaf(xv.y, ur[g].y, fmaf(xv.z, ur[g].z, fmaf(xv.w, ur[g].w, acc))));
        }
      }
      part[r] = acc;
    }
    // transpose-reduce 16 partial sums over the warp: lane l ends with the total of token (l & 15)
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) {      // 16 -> 8 -> 4 -> 2 -> 1 values per lane
      const int bit = w;                    // lanes with (lane & bit) keep the upper half
#pragma unroll
      for (int i = 0; i < w; ++i) {
        const float keep = (lane & bit) ? part[i + w] : part[i];
        const float send = (lane & bit) ? part[i] : part[i + w];
        part[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
      }
    }
    part[0] += __shfl_xor_sync(0xffffffffu, part[0], 16);
    // after the butterfly lane l holds token index with bits (l&8 -> +8, l&4 -> +4, l&2 -> +2, l&1 -> +1)
    if (lane < 16) {
      const int n = c * kProbeChunk + lane;
      if (n < S) sc[warp * Sp + n] = part[0];
    }
    // (no barrier here: the next iteration writes the OTHER buffer, and its __syncthreads orders
    // this chunk's reads before the write two iterations ahead)
  }
  __syncwarp();
  // ---- softmax over the frame's tokens (this warp's head)
  float mx = -INFINITY;
  for (int n = lane; n < S; n += 32) mx = fmaxf(mx, sc[warp * Sp + n]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int n = lane; n < S; n += 32) {
    const float e = __expf(sc[warp * Sp + n] - mx);
    sc[warp * Sp + n] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  __syncthreads();                         // every warp is done with pass 1's buffers
  // ---- pass 2: s = sum_n p(n) x_n; lane owns elements g*128 + lane*4 .. +3
  float4 acc[kMaxGroups];
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
  fetch(0);
  for (int c = 0; c < chunks; ++c) {
    const int buf = c & 1;
    commit(buf);
    if (c + 1 < chunks) fetch(c + 1);
    __syncthreads();
    const float* xb = sx + buf * kProbeChunk * D;
    const int rows = S - c * kProbeChunk < kProbeChunk ? S - c * kProbeChunk : kProbeChunk;
    for (int r = 0; r < rows; ++r) {
      const float p = sc[warp * Sp + c * kProbeChunk + r];
#pragma unroll
      for (int g = 0; g < kMaxGroups; ++g) {
        if (g < groups) {
          const float4 xv = *reinterpret_cast<const float4*>(xb + r * D + g * 128 + lane * 4);
          acc[g].x = fmaf(p, xv.x, acc[g].x); acc[g].y = fmaf(p, xv.y, acc[g].y);
          acc[g].z = fmaf(p, xv.z, acc[g].z); acc[g].w = fmaf(p, xv.w, acc[g].w);
        }
      }
    }
  }
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) {
    if (g < groups)
      *reinterpret_cast<float4*>(sv + warp * D + g * 128 + lane * 4) = make_float4(acc[g].x * inv, acc[g].y * inv, acc[g].z * inv, acc[g].w * inv);
  }
  __syncwarp();
  // ---- value projection of the pooled token: lane computes outputs d = lane and d = lane + 32 of this head
  const float* sh = sv + warp * D;
  const T* w0 = wv + static_cast<long>(warp * kHd + lane) * D;
  const T* w1 = w0 + static_cast<long>(32) * D;
  float r0 = 0.f, r1 = 0.f;
#pragma unroll 4
  for (int k = 0; k < D; k += 8) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(w0 + k));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(w1 + k));
    const float4 s0 = *reinterpret_cast<const float4*>(sh + k);
    const float4 s1 = *reinterpret_cast<const float4*>(sh + k + 4);
    const float2 a0 = Pack2<T>::unpack(a.x), a1 = Pack2<T>::unpack(a.y), a2 = Pack2<T>::unpack(a.z), a3 = Pack2<T>::unpack(a.w);
    const float2 b0 = Pack2<T>::unpack(b.x), b1 = Pack2<T>::unpack(b.y), b2 = Pack2<T>::unpack(b.z), b3 = Pack2<T>::unpack(b.w);
    r0 += a0.x * s0.x + a0.y * s0.y + a1.x * s0.z + a1.y * s0.w + a2.x * s1.x + a2.y * s1.y + a3.x * s1.z + a3.y * s1.w;
    r1 += b0.x * s0.x + b0.y * s0.y + b1.x * s0.z + b1.y * s0.w + b2.x * s1.x + b2.y * s1.y + b3.x * s1.z + b3.y * s1.w;
  }
  T* orow = out + frame * out_ld + warp * kHd;
  orow[lane] = static_cast<T>(r0 + __ldg(bv + warp * kHd + lane));
  orow[lane + 32] = static_cast<T>(r1 + __ldg(bv + warp * kHd + lane + 32));
}

// ------------------------------------------------------------------------------- SigLIP task head
// The step right after the encoder: the zero-shot classification head and the SigLIP sigmoid loss
// on the (all-gathered) last-frame pooler_output (reference models/modeling_timesformer_siglip.py:1704-1726 and
// 244-295, 2324-2351), as ONE launch: L2 normalisation of the image rows (and optionally of the text rows) folded
// into the epilogue as fp32 row / column scales, logits = exp(logit_scale) * <x^, t^> + logit_bias on the tensor
// cores (mma.sync m16n8k16, fp32 accumulate), -logsigmoid(label * logit) summed into a device scalar, and — for
// training — d loss / d logits in the same pass.  Labels: +1 at column targets[i] (classification) or at column
// i + diag_offset (contrastive; diag_offset < 0: negatives only, the reference's ring-exchange terms), -1 elsewhere.
struct HeadArgs {
  const void* image; long ld_i;        // [B, D]
  const void* text; long ld_t;         // [L, D]
  int B, L, D;
  const float* logit_scale;            // device scalar, pre-exp
  const float* logit_bias;             // device scalar or nullptr
  int norm_image, norm_text;
  const long long* targets;            // [B] or nullptr
  int diag_offset;
  float loss_scale;                    // 1 / batch divisor
  float* logits; long ld_l;            // [B, L] fp32 or nullptr
  float* loss;                         // += sum of -logsigmoid(label * logit) * loss_scale
  void* dlogits; long ld_d;            // [B, L] activation dtype or nullptr: d loss / d logit
  float* dparams;                      // nullptr or [2]: += d loss / d logit_scale, d loss / d logit_bias
};

template <typename T>
__global__ void __launch_bounds__(128) siglip_head_kernel(const HeadArgs a) {
  constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
  __shared__ __align__(1024) uint8_t a_tile[16 * 128];
  __shared__ __align__(1024) uint8_t b_tile[128 * 128];
  __shared__ float inv_a[16], inv_b[128], red[3][4];
  griddep_wait();
  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;
  const int m0 = blockIdx.x * 16, n0 = blockIdx.y * 128;
  const T* img = reinterpret_cast<const T*>(a.image);
  const T* txt = reinterpret_cast<const T*>(a.text);
  // fp32 row norms (one warp per row)
  for (int r = warp; r < 16 + 128; r += 4) {
    const bool is_a = r < 16;
    const int row = is_a ? m0 + r : n0 + (r - 16);
    const bool valid = is_a ? row < a.B : row < a.L;
    const bool want = is_a ? a.norm_image != 0 : a.norm_text != 0;
    float ss = 0.f;
    if (valid && want) {
      const T* p = is_a ? img + static_cast<long>(row) * a.ld_i : txt + static_cast<long>(row) * a.ld_t;
      for (int d = lane; d < a.D; d += 32) { const float v = static_cast<float>(p[d]); ss = fmaf(v, v, ss); }
      ss = warp_sum(ss);
    }
    if (lane == 0) {
      const float inv = (valid && want) ? rsqrtf(ss) : 1.0f;
      if (is_a) inv_a[r] = inv; else inv_b[r - 16] = inv;
    }
  }
  float acc[2][2][4];     // [16-column group][n-tile][fragment]
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
  for (int k0 = 0; k0 < a.D; k0 += 64) {
    __syncthreads();
    {   // A: 16 rows x 8 chunks = one 16-byte chunk per thread; B: 128 rows x 8 chunks = 8 per thread
      const int row = threadIdx.x >> 3, ch = threadIdx.x & 7;
      const bool ok = (m0 + row) < a.B && (k0 + ch * 8) < a.D;
      cp_async_16(a_tile + tile_off(row, ch), img + static_cast<long>(ok ? m0 + row : 0) * a.ld_i + (ok ? k0 + ch * 8 : 0), ok);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = threadIdx.x + 128 * i;
        const int brow = idx >> 3, bch = idx & 7;
        const bool okb = (n0 + brow) < a.L && (k0 + bch * 8) < a.D;
        cp_async_16(b_tile + tile_off(brow, bch), txt + static_cast<long>(okb ? n0 + brow : 0) * a.ld_t + (okb ? k0 + bch * 8 : 0), okb);
      }
      cp_async_commit();
      cp_async_wait<0>();
    }
    __syncthreads();
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(qa[ks], smem_u32(a_tile) + tile_off(lane & 15, 2 * ks + (lane >> 4)));
    qk_16keys<kBf16>(acc[0][0], acc[0][1], qa, smem_u32(b_tile), warp * 32, lane);
    qk_16keys<kBf16>(acc[1][0], acc[1][1], qa, smem_u32(b_tile), warp * 32 + 16, lane);
  }
  const float scale = expf(*a.logit_scale);
  const float bias = a.logit_bias ? *a.logit_bias : 0.f;
  float loss = 0.f, dscale = 0.f, dbias = 0.f;
#pragma unroll
  for (int grp = 0; grp < 2; ++grp)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = (e < 2) ? g : g + 8;
        const int col_in = warp * 32 + grp * 16 + nt * 8 + 2 * c + (e & 1);
        const int row = m0 + r, col = n0 + col_in;
        if (row < a.B && col < a.L) {
          const float dot = acc[grp][nt][e] * inv_a[r] * inv_b[col_in];
          const float z = fmaf(scale, dot, bias);
          if (a.logits) a.logits[static_cast<long>(row) * a.ld_l + col] = z;
          float label = -1.f;
          if (a.targets) { if (a.targets[row] == col) label = 1.f; }
          else if (a.diag_offset >= 0 && col == row + a.diag_offset) label = 1.f;
          const float y = label * z;
          // -logsigmoid(y) = max(-y, 0) + log1p(exp(-|y|))
          loss += fmaxf(-y, 0.f) + log1pf(expf(-fabsf(y)));
          if (a.dlogits) {
            const float dz = -label / (1.0f + expf(y)) * a.loss_scale;      // d(-logsigmoid(y))/dz = -label * sigmoid(-y)
            reinterpret_cast<T*>(a.dlogits)[static_cast<long>(row) * a.ld_d + col] = static_cast<T>(dz);
            dscale = fmaf(dz, z - bias, dscale);     // d z / d logit_scale = exp(logit_scale) * dot = z - bias
            dbias += dz;
          }
        }
      }
  loss = warp_sum(loss); dscale = warp_sum(dscale); dbias = warp_sum(dbias);
  if (lane == 0) { red[0][warp] = loss; red[1][warp] = dscale; red[2][warp] = dbias; }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(a.loss, (red[0][0] + red[0][1] + red[0][2] + red[0][3]) * a.loss_scale);
    if (a.dparams) {
      atomicAdd(a.dparams, red[1][0] + red[1][1] + red[1][2] + red[1][3]);
      atomicAdd(a.dparams + 1, red[2][0] + red[2][1] + red[2][2] + red[2][3]);
    }
  }
}

// Backward of the row-wise L2 normalisation x^ = x / |x| feeding the head above:
//   dx = (g - x^ (x^ . g)) / |x|,  g = gscale * dxhat  (gscale = exp(logit_scale), a device scalar or nullptr)
template <typename T>
__global__ void __launch_bounds__(128) l2norm_bwd_kernel(const T* __restrict__ x, long ldx, const T* __restrict__ dxhat, long ldg,
                                                         const float* __restrict__ gscale, T* __restrict__ dx, long ldo, int B, int D) {
  griddep_wait();
  griddep_launch_dependents();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float gs = gscale ? expf(*gscale) : 1.0f;
  const T* xr = x + row * ldx;
  const T* gr = dxhat + row * ldg;
  float ss = 0.f, dot = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = static_cast<float>(xr[d]), g = static_cast<float>(gr[d]) * gs;
    ss = fmaf(v, v, ss);
    dot = fmaf(v, g, dot);
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float r = rsqrtf(ss);
  const float k = dot * r * r;          // (x^ . g) / |x|  expressed on the un-normalised x
  for (int d = lane; d < D; d += 32) {
    const float v = static_cast<float>(xr[d]), g = static_cast<float>(gr[d]) * gs;
    dx[row * ldo + d] = static_cast<T>((g - v * k) * r);
  }
}

// ------------------------------------------------------------------------------- attention backward
// Backward of softmax(scale Q K^T [+ causal mask]) V for both attention cores of a layer:
//   temporal: one task = (site, head), L = T frames, rows consecutive               (…siglip.py:575-615)
//   spatial : one task = (frame, head), L = S tokens, rows T_inner apart in the (b,n,t) stream  (…siglip.py:688-717)
// All of Q, K, V, dO of a task sit in shared memory (128B-swizzled 16-row tiles); probabilities are recomputed
// (never stored by the forward), so the kernel makes two passes without any atomics:
//   pass 1, warp = 16-query tile:  row log-sum-exp, then dS = P o (dO V^T - D), dQ = scale dS K
//   pass 2, warp = 16-key tile  :  P^T, dS^T recomputed against every query tile, dV = P^T dO, dK = scale dS^T Q
// with D_i = <dO_i, O_i>.  Tensor-core work on mma.sync m16n8k16 (fp32 accumulate), P / dS rounded to the 16-bit
// operand type between the two contractions exactly as the forward rounds P.
struct AttnBwdArgs {
  const void* qkv; long ld;          // q | k | v column blocks of width heads*64
  const void* out; long ld_o;        // forward output (ctx)
  const void* dout; long ld_do;
  void* dqkv; long ld_dq;            // dq | dk | dv, same layout as qkv
  int L, heads, T_inner, S, mode, causal;
  long tasks;
  float scale, scale_log2;
};

template <typename T, int kTiles, int kTasksPerCta>
__global__ void __launch_bounds__(kTiles * kTasksPerCta * 32) attn_bwd_kernel(const AttnBwdArgs a) {
  constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
  constexpr int Lp = kTiles * 16;
  constexpr int kTileBytes = Lp * 128;
  constexpr int kTaskBytes = 4 * kTileBytes + 2 * Lp * 4;
  extern __shared__ uint8_t bsm_raw[];
  uint8_t* bsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bsm_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  griddep_wait();
  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;
  const int slot = warp / kTiles, tw = warp % kTiles;           // task slot inside the CTA, tile owned by this warp
  const long task = static_cast<long>(blockIdx.x) * kTasksPerCta + slot;
  const bool live = task < a.tasks;
  uint8_t* base = bsm + static_cast<size_t>(slot) * ((kTaskBytes + 1023) / 1024 * 1024);
  uint8_t *Qs = base, *Ks = base + kTileBytes, *Vs = base + 2 * kTileBytes, *Gs = base + 3 * kTileBytes;
  float* Lse = reinterpret_cast<float*>(base + 4 * kTileBytes);
  float* Dv = Lse + Lp;
  const int D = a.heads * kHd;
  const long grp = live ? task / a.heads : 0;
  const int h = live ? static_cast<int>(task % a.heads) : 0;
  long row0, rstride;
  if (a.mode == 0) { row0 = grp * a.L; rstride = 1; }
  else if (a.T_inner > 1) { row0 = (grp / a.T_inner) * a.S * a.T_inner + grp % a.T_inner; rstride = a.T_inner; }
  else { row0 = grp * a.S; rstride = 1; }
  const T* qkv = reinterpret_cast<const T*>(a.qkv);
  const T* dO = reinterpret_cast<const T*>(a.dout);
  const T* Og = reinterpret_cast<const T*>(a.out);
  // ---- stage Q, K, V, dO (rows >= L zero-filled)
  for (int i = tw * 32 + lane; i < Lp * 8; i += kTiles * 32) {
    const int r = i >> 3, ch = i & 7;
    const bool ok = live && r < a.L;
    const long grow = row0 + static_cast<long>(ok ? r : 0) * rstride;
    const T* src = qkv + grow * a.ld + h * kHd + ch * 8;
    cp_async_16(Qs + tile_off(r, ch), src, ok);
    cp_async_16(Ks + tile_off(r, ch), src + D, ok);
    cp_async_16(Vs + tile_off(r, ch), src + 2 * D, ok);
    cp_async_16(Gs + tile_off(r, ch), dO + grow * a.ld_do + h * kHd + ch * 8, ok);
  }
  cp_async_commit();
  // D_i = <dO_i, O_i> for the rows of this warp's tile (straight from global memory)
  for (int rr = 0; rr < 16; ++rr) {
    const int r = tw * 16 + rr;
    float d = 0.f;
    if (live && r < a.L) {
      const long grow = row0 + static_cast<long>(r) * rstride;
      const float2 x = Pack2<T>::unpack(*reinterpret_cast<const uint32_t*>(dO + grow * a.ld_do + h * kHd + 2 * lane));
      const float2 y = Pack2<T>::unpack(*reinterpret_cast<const uint32_t*>(Og + grow * a.ld_o + h * kHd + 2 * lane));
      d = warp_sum(x.x * y.x + x.y * y.y);
    }
    if (lane == 0) Dv[r] = d;
  }
  cp_async_wait<0>();
  if constexpr (kTiles * kTasksPerCta > 1) __syncthreads(); else __syncwarp();
  const uint32_t q_u = smem_u32(Qs), k_u = smem_u32(Ks), v_u = smem_u32(Vs), g_u = smem_u32(Gs);
  const int ntiles = (a.L + 15) >> 4;
  T* dq_out = reinterpret_cast<T*>(a.dqkv);

  auto load_a = [&](uint32_t (&fr)[4][4], uint32_t tile_u, int r0) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(fr[ks], tile_u + tile_off(r0 + (lane & 15), 2 * ks + (lane >> 4)));
  };
  // scaled + masked scores of a 16 x 16 block: rows i0+g / i0+g+8 (fragment halves), columns j0 + ...
  auto mask_scale = [&](float (&s0)[4], float (&s1)[4], int i0, int j0, bool rows_are_queries) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ri = i0 + g + ((e >> 1) << 3);
      const int c0 = j0 + 2 * c + (e & 1), c1 = c0 + 8;
      const int q0 = rows_are_queries ? ri : c0, k0 = rows_are_queries ? c0 : ri;
      const int q1 = rows_are_queries ? ri : c1, k1 = rows_are_queries ? c1 : ri;
      const bool ok0 = k0 < a.L && (!a.causal || k0 <= q0);
      const bool ok1 = k1 < a.L && (!a.causal || k1 <= q1);
      s0[e] = ok0 ? s0[e] * a.scale_log2 : -INFINITY;
      s1[e] = ok1 ? s1[e] * a.scale_log2 : -INFINITY;
    }
  };

  if (live && tw < ntiles) {
    // ================================================================ pass 1: this warp's 16 queries
    const int i0 = tw * 16;
    uint32_t qa[4][4], ga[4][4];
    load_a(qa, q_u, i0);
    load_a(ga, g_u, i0);
    const int kend = a.causal ? tw + 1 : ntiles;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    for (int kb = 0; kb < kend; ++kb) {
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      qk_16keys<kBf16>(s0, s1, qa, k_u, kb * 16, lane);
      mask_scale(s0, s1, i0, kb * 16, true);
      float mx0 = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1]));
      float mx1 = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
      const float mu0 = (mn0 == -INFINITY) ? 0.f : mn0, mu1 = (mn1 == -INFINITY) ? 0.f : mn1;
      float p0 = exp2f(s0[0] - mu0) + exp2f(s0[1] - mu0) + exp2f(s1[0] - mu0) + exp2f(s1[1] - mu0);
      float p1 = exp2f(s0[2] - mu1) + exp2f(s0[3] - mu1) + exp2f(s1[2] - mu1) + exp2f(s1[3] - mu1);
      p0 += __shfl_xor_sync(0xffffffffu, p0, 1); p0 += __shfl_xor_sync(0xffffffffu, p0, 2);
      p1 += __shfl_xor_sync(0xffffffffu, p1, 1); p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
      l_run[0] = l_run[0] * exp2f(m_run[0] - mu0) + p0;
      l_run[1] = l_run[1] * exp2f(m_run[1] - mu1) + p1;
      m_run[0] = mn0; m_run[1] = mn1;
    }
    const float lse0 = m_run[0] + log2f(l_run[0]), lse1 = m_run[1] + log2f(l_run[1]);
    if (c == 0) { Lse[i0 + g] = lse0; Lse[i0 + g + 8] = lse1; }
    const float d0 = Dv[i0 + g], d1 = Dv[i0 + g + 8];
    float dq[8][4];
#pragma unroll
    for (int d = 0; d < 8; ++d) { dq[d][0] = dq[d][1] = dq[d][2] = dq[d][3] = 0.f; }
    for (int kb = 0; kb < kend; ++kb) {
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f};
      qk_16keys<kBf16>(s0, s1, qa, k_u, kb * 16, lane);
      mask_scale(s0, s1, i0, kb * 16, true);
      qk_16keys<kBf16>(t0, t1, ga, v_u, kb * 16, lane);                  // dP = dO V^T
      float e0[4], e1[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float ls = (e < 2) ? lse0 : lse1, dd = (e < 2) ? d0 : d1;
        e0[e] = exp2f(s0[e] - ls) * (t0[e] - dd);                        // dS = P o (dP - D)
        e1[e] = exp2f(s1[e] - ls) * (t1[e] - dd);
      }
      uint32_t pa[4];
      pa[0] = Pack2<T>::pack(e0[0], e0[1]); pa[1] = Pack2<T>::pack(e0[2], e0[3]);
      pa[2] = Pack2<T>::pack(e1[0], e1[1]); pa[3] = Pack2<T>::pack(e1[2], e1[3]);
      pv_16keys<kBf16>(dq, pa, k_u, kb * 16, lane);                      // dQ += dS K
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = i0 + g + half * 8;
      if (r < a.L) {
        T* dst = dq_out + (row0 + static_cast<long>(r) * rstride) * a.ld_dq + h * kHd;
#pragma unroll
        for (int d = 0; d < 8; ++d)
          *reinterpret_cast<uint32_t*>(dst + d * 8 + 2 * c) = Pack2<T>::pack(dq[d][2 * half] * a.scale, dq[d][2 * half + 1] * a.scale);
      }
    }
  }
  if constexpr (kTiles * kTasksPerCta > 1) __syncthreads(); else __syncwarp();
  if (live && tw < ntiles) {
    // ================================================================ pass 2: this warp's 16 keys
    const int j0 = tw * 16;
    uint32_t ka[4][4], va[4][4];
    load_a(ka, k_u, j0);
    load_a(va, v_u, j0);
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int d = 0; d < 8; ++d) { dk[d][0] = dk[d][1] = dk[d][2] = dk[d][3] = 0.f; dv[d][0] = dv[d][1] = dv[d][2] = dv[d][3] = 0.f; }
    const int qbeg = a.causal ? tw : 0;
    for (int qb = qbeg; qb < ntiles; ++qb) {
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f};
      qk_16keys<kBf16>(s0, s1, ka, q_u, qb * 16, lane);                  // S^T block: rows = keys, columns = queries
      mask_scale(s0, s1, j0, qb * 16, false);
      qk_16keys<kBf16>(t0, t1, va, g_u, qb * 16, lane);                  // dP^T = V dO^T
      float p0[4], p1[4], e0[4], e1[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int q0 = qb * 16 + 2 * c + (e & 1), q1 = q0 + 8;
        p0[e] = exp2f(s0[e] - Lse[q0]);
        p1[e] = exp2f(s1[e] - Lse[q1]);
        e0[e] = p0[e] * (t0[e] - Dv[q0]);
        e1[e] = p1[e] * (t1[e] - Dv[q1]);
      }
      uint32_t pa[4];
      pa[0] = Pack2<T>::pack(p0[0], p0[1]); pa[1] = Pack2<T>::pack(p0[2], p0[3]);
      pa[2] = Pack2<T>::pack(p1[0], p1[1]); pa[3] = Pack2<T>::pack(p1[2], p1[3]);
      pv_16keys<kBf16>(dv, pa, g_u, qb * 16, lane);                      // dV += P^T dO
      pa[0] = Pack2<T>::pack(e0[0], e0[1]); pa[1] = Pack2<T>::pack(e0[2], e0[3]);
      pa[2] = Pack2<T>::pack(e1[0], e1[1]); pa[3] = Pack2<T>::pack(e1[2], e1[3]);
      pv_16keys<kBf16>(dk, pa, q_u, qb * 16, lane);                      // dK += dS^T Q
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = j0 + g + half * 8;
      if (r < a.L) {
        T* dst = dq_out + (row0 + static_cast<long>(r) * rstride) * a.ld_dq + h * kHd;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          *reinterpret_cast<uint32_t*>(dst + D + d * 8 + 2 * c) = Pack2<T>::pack(dk[d][2 * half] * a.scale, dk[d][2 * half + 1] * a.scale);
          *reinterpret_cast<uint32_t*>(dst + 2 * D + d * 8 + 2 * c) = Pack2<T>::pack(dv[d][2 * half], dv[d][2 * half + 1]);
        }
      }
    }
  }
}

// Backward of the pooling attention of the SigLIP head (one learned probe as the only query, …siglip.py:1141-1148):
// per (frame, head): p = softmax_n(q . K_n), out = sum_n p_n V_n.  One warp per task; lane owns keys lane, lane+32, ...
//   dV_n = p_n dout, dK_n = ds_n q, dq += sum_n ds_n K_n with ds = p o (dout . V_n - sum_m p_m dout . V_m)
template <typename T>
__global__ void __launch_bounds__(128) pool_attn_bwd_kernel(const T* __restrict__ kv, long ld_kv, const float* __restrict__ q,
                                                            const T* __restrict__ dout, long ld_do, T* __restrict__ dkv, long ld_dkv,
                                                            float* __restrict__ dq, int frames, int heads, int S) {
  griddep_wait();
  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long task = static_cast<long>(blockIdx.x) * 4 + warp;
  if (task >= static_cast<long>(frames) * heads) return;
  const int f = static_cast<int>(task / heads), h = static_cast<int>(task % heads);
  const int D = heads * kHd;
  constexpr int kMaxPer = 32;                                    // S <= 1024
  float sc[kMaxPer], dp[kMaxPer];
  const float* qh = q + h * kHd;
  const T* go = dout + static_cast<long>(f) * ld_do + h * kHd;
  float mx = -INFINITY;
  int cnt = 0;
  for (int n = lane; n < S; n += 32, ++cnt) {
    const T* kr = kv + (static_cast<long>(f) * S + n) * ld_kv + h * kHd;
    float s = 0.f, d = 0.f;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      float kf[8], vf[8], gf[8];
      const uint4 ku = *reinterpret_cast<const uint4*>(kr + ch * 8), vu = *reinterpret_cast<const uint4*>(kr + D + ch * 8);
      const uint4 gu = *reinterpret_cast<const uint4*>(go + ch * 8);
      const uint32_t kw[4] = {ku.x, ku.y, ku.z, ku.w}, vw[4] = {vu.x, vu.y, vu.z, vu.w}, gw[4] = {gu.x, gu.y, gu.z, gu.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = Pack2<T>::unpack(kw[e]), b = Pack2<T>::unpack(vw[e]), cc = Pack2<T>::unpack(gw[e]);
        kf[2 * e] = a.x; kf[2 * e + 1] = a.y; vf[2 * e] = b.x; vf[2 * e + 1] = b.y; gf[2 * e] = cc.x; gf[2 * e + 1] = cc.y;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { s = fmaf(kf[e], qh[ch * 8 + e], s); d = fmaf(vf[e], gf[e], d); }
    }
    sc[cnt] = s; dp[cnt] = d;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float l = 0.f;
  for (int i = 0; i < cnt; ++i) { sc[i] = __expf(sc[i] - mx); l += sc[i]; }
  l = warp_sum(l);
  const float inv = 1.0f / l;
  float pd = 0.f;
  for (int i = 0; i < cnt; ++i) { sc[i] *= inv; pd = fmaf(sc[i], dp[i], pd); }
  pd = warp_sum(pd);
  float dqa[kHd];
#pragma unroll
  for (int e = 0; e < kHd; ++e) dqa[e] = 0.f;
  cnt = 0;
  for (int n = lane; n < S; n += 32, ++cnt) {
    const float p = sc[cnt], ds = p * (dp[cnt] - pd);
    const T* kr = kv + (static_cast<long>(f) * S + n) * ld_kv + h * kHd;
    T* dr = dkv + (static_cast<long>(f) * S + n) * ld_dkv + h * kHd;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      const uint4 ku = *reinterpret_cast<const uint4*>(kr + ch * 8), gu = *reinterpret_cast<const uint4*>(go + ch * 8);
      const uint32_t kw[4] = {ku.x, ku.y, ku.z, ku.w}, gw[4] = {gu.x, gu.y, gu.z, gu.w};
      uint32_t ok[4], ov[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = Pack2<T>::unpack(kw[e]), cc = Pack2<T>::unpack(gw[e]);
        dqa[ch * 8 + 2 * e] = fmaf(ds, a.x, dqa[ch * 8 + 2 * e]);
        dqa[ch * 8 + 2 * e + 1] = fmaf(ds, a.y, dqa[ch * 8 + 2 * e + 1]);
        ok[e] = Pack2<T>::pack(ds * qh[ch * 8 + 2 * e], ds * qh[ch * 8 + 2 * e + 1]);
        ov[e] = Pack2<T>::pack(p * cc.x, p * cc.y);
      }
      *reinterpret_cast<uint4*>(dr + ch * 8) = make_uint4(ok[0], ok[1], ok[2], ok[3]);
      *reinterpret_cast<uint4*>(dr + D + ch * 8) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
    }
  }
  if (dq) {
#pragma unroll
    for (int e = 0; e < kHd; ++e) {
      const float v = warp_sum(dqa[e]);
      if (lane == 0) atomicAdd(dq + h * kHd + e, v);
    }
  }
}

int check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

}  // namespace

int temporal_attention(cudaStream_t stream, int dtype, const void* qkv, int ld_qkv, const void* kcache,
                       const void* vcache, int Tcap, void* out, int ld_out, int sites, int heads, int Tq,
                       int Tk, int q_off, int causal, float scale, const int* seen_dev) {
  if (sites <= 0 || Tq <= 0) return 0;
  if (seen_dev && !kcache) { set_error("temporal_attention: seen_dev needs a cache"); return -1; }
  if (dtype != kBF16 && dtype != kF16) { set_error("temporal_attention: dtype must be bf16/f16"); return -1; }
  if ((ld_qkv % 8) || (ld_out % 2)) { set_error("temporal_attention: bad leading dims"); return -1; }
  const int D = heads * kHd;
  TemporalArgs a;
  a.q = qkv; a.q_ld = ld_qkv;
  const size_t es = 2;
  if (kcache) {
    if (Tk > Tcap) { set_error("temporal_attention: Tk=%d exceeds cache capacity %d", Tk, Tcap); return -1; }
    a.k = kcache; a.v = vcache;
    a.kv_site_stride = static_cast<long>(heads) * Tcap * kHd;
    a.kv_head_stride = static_cast<long>(Tcap) * kHd;
    a.kv_row_stride = kHd;
  } else {
    if (Tk != Tq) { set_error("temporal_attention: Tk must equal Tq without a cache"); return -1; }
    a.k = reinterpret_cast<const uint8_t*>(qkv) + static_cast<size_t>(D) * es;
    a.v = reinterpret_cast<const uint8_t*>(qkv) + static_cast<size_t>(2 * D) * es;
    a.kv_site_stride = static_cast<long>(Tq) * ld_qkv;
    a.kv_head_stride = kHd;
    a.kv_row_stride = ld_qkv;
  }
  a.out = out; a.out_ld = ld_out;
  a.sites = sites; a.heads = heads; a.Tq = Tq; a.Tk = Tk; a.q_off = q_off; a.causal = causal;
  a.qtiles = (Tq + 15) / 16;
  a.tasks = static_cast<long>(sites) * heads * a.qtiles;
  a.scale_log2 = scale * kLog2e;
  a.seen_dev = seen_dev;
  const long blocks = (a.tasks + kTWarps - 1) / kTWarps;
  ProfScope ps(stream, kProfTemporalAttn, 4.0 * sites * heads * static_cast<double>(Tq) * Tk * kHd,
               2.0 * sites * heads * kHd * (2.0 * Tq + 2.0 * Tk));
  static const bool fast_on = [] { const char* e = getenv("SF_TEMPORAL_FAST"); return !(e && e[0] == '0'); }();
  if (fast_on && !kcache && Tq == Tk && Tq <= 16 && q_off == 0 && (ld_out % 8) == 0) {
    const long max_blocks = static_cast<long>(num_sms()) * 4;
    const size_t smem = static_cast<size_t>(kTWarps) * 2 * 3 * 2048;
    LaunchCfg lf(dim3(static_cast<unsigned>(blocks < max_blocks ? blocks : max_blocks)), dim3(kTWarps * 32), smem, stream);
    if (dtype == kBF16) cudaLaunchKernelEx(&lf.cfg, temporal_attn_fast_kernel<__nv_bfloat16>, a);
    else cudaLaunchKernelEx(&lf.cfg, temporal_attn_fast_kernel<__half>, a);
    return check_launch("temporal_attention");
  }
  LaunchCfg lc(dim3(static_cast<unsigned>(blocks)), dim3(kTWarps * 32), 0, stream);
  if (dtype == kBF16) cudaLaunchKernelEx(&lc.cfg, temporal_attn_kernel<__nv_bfloat16>, a);
  else cudaLaunchKernelEx(&lc.cfg, temporal_attn_kernel<__half>, a);
  return check_launch("temporal_attention");
}

int g_decode_tma_opt = -1;
void set_decode_tma(int on) { g_decode_tma_opt = on; }

bool temporal_decode_supported(int Tcap, int Tq) {
  static const bool on = [] { const char* e = getenv("SF_TEMPORAL_DECODE"); return !(e && e[0] == '0'); }();
  return on && Tq == 1 && Tcap <= kDecMaxRows && Tcap >= 1;
}

// 2-D map over a cache tensor viewed as [sites*heads*Tcap rows][64]: 16-row boxes, SWIZZLE_128B
int make_cache_map(CUtensorMap* map, int dtype, const void* base, long rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -3;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(kHd), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(kHd) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kHd), 16u};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(cache map) failed (%d): rows=%ld base=%p", (int)r, rows, base); return -3; }
  return 0;
}

int temporal_decode(cudaStream_t stream, int dtype, const void* qkv, int ld_qkv, void* kcache, void* vcache, int Tcap,
                    void* out, int ld_out, int sites, int heads, int seen, float scale, const int* seen_dev) {
  if (sites <= 0) return 0;
  if (dtype != kBF16 && dtype != kF16) { set_error("temporal_decode: dtype must be bf16/f16"); return -1; }
  if (!temporal_decode_supported(Tcap, 1)) { set_error("temporal_decode: cache capacity %d exceeds %d frames", Tcap, kDecMaxRows); return -1; }
  if (static_cast<long>(sites) * heads * Tcap > 0x7fffffffL) { set_error("temporal_decode: cache too large for 32-bit row coordinates"); return -1; }
  if (seen + 1 > Tcap) { set_error("temporal_decode: %d + 1 frames exceed cache capacity %d", seen, Tcap); return -1; }
  if ((ld_qkv % 2) || (ld_out % 2)) { set_error("temporal_decode: bad leading dims"); return -1; }
  TemporalArgs a;
  a.q = qkv; a.q_ld = ld_qkv;
  a.k = kcache; a.v = vcache;
  a.kv_site_stride = static_cast<long>(heads) * Tcap * kHd;
  a.kv_head_stride = static_cast<long>(Tcap) * kHd;
  a.kv_row_stride = kHd;
  a.out = out; a.out_ld = ld_out;
  a.sites = sites; a.heads = heads; a.Tq = 1; a.Tk = seen + 1; a.q_off = seen; a.causal = 1; a.qtiles = 1;
  a.tasks = static_cast<long>(sites) * heads;
  a.scale_log2 = scale * kLog2e;
  a.seen_dev = seen_dev;
  ProfScope ps(stream, kProfTemporalAttn, 4.0 * sites * heads * static_cast<double>(seen + 1) * kHd,
               2.0 * sites * heads * kHd * (5.0 + 2.0 * seen));
  // SF_DECODE_TMA=1 selects the TMA-ring / mma.sync kernel; the default is the register-direct kernel
  static const bool tma_env = [] { const char* e = getenv("SF_DECODE_TMA"); return e && e[0] == '1'; }();
  const bool use_tma = g_decode_tma_opt < 0 ? tma_env : g_decode_tma_opt != 0;
  if (!use_tma && (ld_qkv % 8 == 0) && (ld_out % 8 == 0) && (reinterpret_cast<uintptr_t>(qkv) % 16 == 0) &&
      (reinterpret_cast<uintptr_t>(out) % 16 == 0)) {
    long blocks = (a.tasks + 7) / 8;
    if (blocks > 2L * num_sms()) blocks = 2L * num_sms();
    LaunchCfg lc(dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream);
    if (dtype == kBF16) cudaLaunchKernelEx(&lc.cfg, temporal_decode_direct_kernel<__nv_bfloat16>, a, reinterpret_cast<__nv_bfloat16*>(kcache),
                                           reinterpret_cast<__nv_bfloat16*>(vcache), Tcap);
    else cudaLaunchKernelEx(&lc.cfg, temporal_decode_direct_kernel<__half>, a, reinterpret_cast<__half*>(kcache),
                            reinterpret_cast<__half*>(vcache), Tcap);
    return check_launch("temporal_decode");
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(temporal_decode_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecSmem);
    cudaFuncSetAttribute(temporal_decode_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecSmem);
    attr_set = true;
  }
  const int kDecWarps = dec_warps(Tcap);
  long blocks = (a.tasks + kDecWarps - 1) / kDecWarps;
  if (blocks > num_sms()) blocks = num_sms();
  CUtensorMap tmK, tmV;
  int rc = make_cache_map(&tmK, dtype, kcache, static_cast<long>(sites) * heads * Tcap);
  if (rc) return rc;
  rc = make_cache_map(&tmV, dtype, vcache, static_cast<long>(sites) * heads * Tcap);
  if (rc) return rc;
  LaunchCfg lc(dim3(static_cast<unsigned>(blocks)), dim3(kDecWarps * 32), kDecSmem, stream);
  if (dtype == kBF16) cudaLaunchKernelEx(&lc.cfg, temporal_decode_kernel<__nv_bfloat16>, tmK, tmV, a, reinterpret_cast<__nv_bfloat16*>(kcache),
                                         reinterpret_cast<__nv_bfloat16*>(vcache), Tcap);
  else cudaLaunchKernelEx(&lc.cfg, temporal_decode_kernel<__half>, tmK, tmV, a, reinterpret_cast<__half*>(kcache),
                          reinterpret_cast<__half*>(vcache), Tcap);
  return check_launch("temporal_decode");
}

template <typename T, int kTiles, int kTasksPerCta>
static int launch_attn_bwd(cudaStream_t stream, const AttnBwdArgs& a) {
  constexpr int Lp = kTiles * 16;
  constexpr int kTaskBytes = ((4 * Lp * 128 + 2 * Lp * 4) + 1023) / 1024 * 1024;
  constexpr int smem = kTaskBytes * kTasksPerCta + 1024;
  auto kernel = attn_bwd_kernel<T, kTiles, kTasksPerCta>;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr_set = true; }
  const long blocks = (a.tasks + kTasksPerCta - 1) / kTasksPerCta;
  LaunchCfg lc(dim3(static_cast<unsigned>(blocks)), dim3(kTiles * kTasksPerCta * 32), smem, stream);
  cudaLaunchKernelEx(&lc.cfg, kernel, a);
  return check_launch("attention_backward");
}

// mode 0: temporal (groups = sites, L = T frames); mode 1: spatial (groups = frames, L = S tokens, rows as in spatial_attention)
int attention_backward(cudaStream_t stream, int dtype, int mode, const void* qkv, int ld_qkv, const void* out, int ld_o, const void* dout,
                       int ld_do, void* dqkv, int ld_dq, int groups, int heads, int L, int T_inner, int causal, float scale) {
  if (groups <= 0 || L <= 0) return 0;
  if (dtype != kBF16 && dtype != kF16) { set_error("attention_backward: dtype must be bf16/f16"); return -1; }
  if ((ld_qkv % 8) || (ld_do % 8) || (ld_o % 2) || (ld_dq % 2)) { set_error("attention_backward: bad leading dims"); return -1; }
  if (L > 208) { set_error("attention_backward: %d tokens per attention group exceed the 208 this round's backward kernel stages in shared memory", L); return -1; }
  AttnBwdArgs a;
  a.qkv = qkv; a.ld = ld_qkv; a.out = out; a.ld_o = ld_o; a.dout = dout; a.ld_do = ld_do; a.dqkv = dqkv; a.ld_dq = ld_dq;
  a.L = L; a.heads = heads; a.T_inner = T_inner; a.S = L; a.mode = mode; a.causal = causal;
  a.tasks = static_cast<long>(groups) * heads;
  a.scale = scale; a.scale_log2 = scale * kLog2e;
  ProfScope ps(stream, mode == 0 ? kProfTemporalAttn : kProfSpatialAttn, 14.0 * groups * heads * static_cast<double>(L) * L * kHd,
               2.0 * groups * heads * kHd * 8.0 * L);
#define SF_ABWD(TT, TP) (dtype == kBF16 ? launch_attn_bwd<__nv_bfloat16, TT, TP>(stream, a) : launch_attn_bwd<__half, TT, TP>(stream, a))
  if (L <= 16) return SF_ABWD(1, 4);
  if (L <= 32) return SF_ABWD(2, 2);
  if (L <= 64) return SF_ABWD(4, 1);
  if (L <= 128) return SF_ABWD(8, 1);
  return SF_ABWD(13, 1);
#undef SF_ABWD
}

int pool_attention_backward(cudaStream_t stream, int dtype, const void* kv, int ld_kv, const float* q, const void* dout, int ld_do,
                            void* dkv, int ld_dkv, float* dq, int frames, int heads, int S) {
  if (frames <= 0) return 0;
  if (dtype != kBF16 && dtype != kF16) { set_error("pool_attention_backward: dtype must be bf16/f16"); return -1; }
  if (S > 1024 || (ld_kv % 8) || (ld_do % 8) || (ld_dkv % 8)) { set_error("pool_attention_backward: S <= 1024 and 16-byte aligned rows required"); return -1; }
  const long tasks = static_cast<long>(frames) * heads;
  ProfScope ps(stream, kProfPoolAttn, 8.0 * tasks * S * kHd, 8.0 * tasks * S * kHd);
  LaunchCfg lc(dim3(static_cast<unsigned>((tasks + 3) / 4)), dim3(128), 0, stream);
  if (dtype == kBF16)
    cudaLaunchKernelEx(&lc.cfg, pool_attn_bwd_kernel<__nv_bfloat16>, reinterpret_cast<const __nv_bfloat16*>(kv), static_cast<long>(ld_kv), q,
                       reinterpret_cast<const __nv_bfloat16*>(dout), static_cast<long>(ld_do), reinterpret_cast<__nv_bfloat16*>(dkv),
                       static_cast<long>(ld_dkv), dq, frames, heads, S);
  else
    cudaLaunchKernelEx(&lc.cfg, pool_attn_bwd_kernel<__half>, reinterpret_cast<const __half*>(kv), static_cast<long>(ld_kv), q,
                       reinterpret_cast<const __half*>(dout), static_cast<long>(ld_do), reinterpret_cast<__half*>(dkv), static_cast<long>(ld_dkv),
                       dq, frames, heads, S);
  return check_launch("pool_attention_backward");
}

int l2norm_backward(cudaStream_t stream, int dtype, const void* x, int ldx, const void* dxhat, int ldg, const float* gscale,
                    void* dx, int ldo, int B, int D) {
  if (B <= 0) return 0;
  if (dtype != kBF16 && dtype != kF16) { set_error("l2norm_backward: dtype must be bf16/f16"); return -1; }
  LaunchCfg lc(dim3(static_cast<unsigned>((B + 3) / 4)), dim3(128), 0, stream);
  ProfScope ps(stream, kProfOther, 0.0, 6.0 * B * D);
  if (dtype == kBF16)
    cudaLaunchKernelEx(&lc.cfg, l2norm_bwd_kernel<__nv_bfloat16>, reinterpret_cast<const __nv_bfloat16*>(x), static_cast<long>(ldx),
                       reinterpret_cast<const __nv_bfloat16*>(dxhat), static_cast<long>(ldg), gscale, reinterpret_cast<__nv_bfloat16*>(dx),
                       static_cast<long>(ldo), B, D);
  else
    cudaLaunchKernelEx(&lc.cfg, l2norm_bwd_kernel<__half>, reinterpret_cast<const __half*>(x), static_cast<long>(ldx),
                       reinterpret_cast<const __half*>(dxhat), static_cast<long>(ldg), gscale, reinterpret_cast<__half*>(dx),
                       static_cast<long>(ldo), B, D);
  return check_launch("l2norm_backward");
}

int siglip_head(cudaStream_t stream, int dtype, const void* image, int ld_i, const void* text, int ld_t, int B, int L, int D,
                const float* logit_scale, const float* logit_bias, int norm_image, int norm_text, const long long* targets,
                int diag_offset, float loss_div, float* logits, int ld_l, float* loss, void* dlogits, int ld_d, float* dparams) {
  if (B <= 0 || L <= 0) return 0;
  if (dtype != kBF16 && dtype != kF16) { set_error("siglip_head: dtype must be bf16/f16"); return -1; }
  if (!image || !text || !logit_scale || !loss) { set_error("siglip_head: null argument"); return -1; }
  if ((D % 8) || (ld_i % 8) || (ld_t % 8) || ((reinterpret_cast<uintptr_t>(image) | reinterpret_cast<uintptr_t>(text)) & 15)) {
    set_error("siglip_head: D and the leading dims must be multiples of 8, 16-byte aligned rows (D=%d ld_i=%d ld_t=%d)", D, ld_i, ld_t);
    return -1;
  }
  if (!(loss_div > 0.f)) { set_error("siglip_head: loss divisor must be positive"); return -1; }
  HeadArgs a;
  a.image = image; a.ld_i = ld_i; a.text = text; a.ld_t = ld_t; a.B = B; a.L = L; a.D = D;
  a.logit_scale = logit_scale; a.logit_bias = logit_bias; a.norm_image = norm_image; a.norm_text = norm_text;
  a.targets = targets; a.diag_offset = diag_offset; a.loss_scale = 1.0f / loss_div;
  a.logits = logits; a.ld_l = ld_l; a.loss = loss; a.dlogits = dlogits; a.ld_d = ld_d; a.dparams = dparams;
  ProfScope ps(stream, kProfOther, 2.0 * B * L * static_cast<double>(D), 2.0 * (static_cast<double>(B) + L) * D + 4.0 * B * L);
  LaunchCfg lc(dim3(static_cast<unsigned>((B + 15) / 16), static_cast<unsigned>((L + 127) / 128)), dim3(128), 0, stream);
  if (dtype == kBF16) cudaLaunchKernelEx(&lc.cfg, siglip_head_kernel<__nv_bfloat16>, a);
  else cudaLaunchKernelEx(&lc.cfg, siglip_head_kernel<__half>, a);
  return check_launch("siglip_head");
}

int kv_append(cudaStream_t stream, int dtype, const void* qkv, int ld_qkv, void* kcache, void* vcache,
              int Tcap, int sites, int heads, int Tq, int pos0, const int* seen_dev) {
  if (sites <= 0 || Tq <= 0) return 0;
  if (pos0 + Tq > Tcap) { set_error("kv_append: %d + %d frames exceed cache capacity %d", pos0, Tq, Tcap); return -1; }
  const long total = static_cast<long>(sites) * Tq * heads * 8;
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  // bf16 and fp16 are both 2-byte payloads: one instantiation moves either
  ProfScope ps(stream, kProfKvAppend, 0.0, 8.0 * sites * heads * static_cast<double>(Tq) * kHd);
  LaunchCfg lc(dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream);
  cudaLaunchKernelEx(&lc.cfg, kv_append_kernel<__nv_bfloat16>, reinterpret_cast<const __nv_bfloat16*>(qkv),
                     static_cast<long>(ld_qkv), reinterpret_cast<__nv_bfloat16*>(kcache),
                     reinterpret_cast<__nv_bfloat16*>(vcache), Tcap, sites, heads, Tq, pos0, seen_dev);
  (void)dtype;
  return check_launch("kv_append");
}

int spatial_attention(cudaStream_t stream, int dtype, const void* qkv, int ld_qkv, void* out, int ld_out,
                      int frames, int heads, int S, int T_inner, float scale, float* probs) {
  if (frames <= 0 || S <= 0) return 0;
  if (dtype != kBF16 && dtype != kF16) { set_error("spatial_attention: dtype must be bf16/f16"); return -1; }
  if ((ld_qkv % 8) || (ld_out % 2)) { set_error("spatial_attention: bad leading dims"); return -1; }
  SpatialArgs a;
  a.qkv = qkv; a.ld = ld_qkv; a.out = out; a.out_ld = ld_out;
  a.frames = frames; a.heads = heads; a.S = S; a.T_inner = T_inner;
  if (T_inner > 1 && frames % T_inner) { set_error("spatial_attention: frames=%d is not a multiple of T_inner=%d", frames, T_inner); return -1; }
  a.qblocks = (S + kSWarps * 16 - 1) / (kSWarps * 16);
  a.scale_log2 = scale * kLog2e;
  const long blocks = static_cast<long>(frames) * heads * a.qblocks;
  static const bool use_tc = [] { const char* e = getenv("SF_SPATIAL_TC"); return !(e && e[0] == '0'); }();
  if (use_tc && spatial_attention_tc_supported(ld_qkv, S)) {
    int rc = spatial_attention_tc(stream, dtype, qkv, ld_qkv, out, ld_out, frames, heads, S, T_inner, scale);
    if (rc) return rc;
  } else {
    ProfScope ps(stream, kProfSpatialAttn, 4.0 * frames * heads * static_cast<double>(S) * S * kHd,
                 2.0 * frames * heads * kHd * 4.0 * S);
    LaunchCfg lc(dim3(static_cast<unsigned>(blocks)), dim3(kSWarps * 32), 0, stream);
    if (dtype == kBF16) cudaLaunchKernelEx(&lc.cfg, spatial_attn_kernel<__nv_bfloat16>, a);
    else cudaLaunchKernelEx(&lc.cfg, spatial_attn_kernel<__half>, a);
    int rc = check_launch("spatial_attention");
    if (rc) return rc;
  }
  if (!probs) return 0;
  const long rows = static_cast<long>(frames) * heads * S;
  const long pb = (rows + 3) / 4;
  LaunchCfg lp(dim3(static_cast<unsigned>(pb)), dim3(128), 0, stream);
  if (dtype == kBF16) cudaLaunchKernelEx(&lp.cfg, spatial_probs_kernel<__nv_bfloat16>, reinterpret_cast<const __nv_bfloat16*>(qkv),
                                         static_cast<long>(ld_qkv), probs, frames, heads, S, T_inner, scale);
  else cudaLaunchKernelEx(&lp.cfg, spatial_probs_kernel<__half>, reinterpret_cast<const __half*>(qkv),
                          static_cast<long>(ld_qkv), probs, frames, heads, S, T_inner, scale);
  return check_launch("spatial_probs");
}

int pool_probe(cudaStream_t stream, int dtype, const void* tokens, int ld, const float* u, const void* wv, const float* bv,
               void* out, int ld_out, int frames, int heads, int S) {
  if (frames <= 0) return 0;
  const int D = heads * kHd;
  if (dtype != kBF16 && dtype != kF16) { set_error("pool_probe: dtype must be bf16/f16"); return -1; }
  if (heads < 1 || heads > 16 || (D % 128) || D > kProbeMaxD || S < 1 || (ld % 8) ||
      kProbeChunk * D / 8 > 8 * heads * 32) {
    set_error("pool_probe: unsupported geometry (heads %d, D %d, S %d)", heads, D, S);
    return -1;
  }
  const size_t smem = static_cast<size_t>(2) * kProbeChunk * D * 4 + static_cast<size_t>(heads) * ((S + 3) & ~3) * 4 +
                      static_cast<size_t>(heads) * D * 4;
  if (smem > 227 * 1024) { set_error("pool_probe: S=%d needs %zu bytes of shared memory", S, smem); return -1; }
  ProfScope ps(stream, kProfPoolAttn, 2.0 * frames * (2.0 * S * heads * D + static_cast<double>(D) * D),
               2.0 * frames * S * D * 2.0);
  LaunchCfg lc(dim3(static_cast<unsigned>(frames)), dim3(heads * 32), smem, stream);
  cudaError_t e;
  if (dtype == kBF16) {
    e = cudaFuncSetAttribute(pool_probe_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess)
      e = cudaLaunchKernelEx(&lc.cfg, pool_probe_kernel<__nv_bfloat16>, reinterpret_cast<const __nv_bfloat16*>(tokens), static_cast<long>(ld), u,
                             reinterpret_cast<const __nv_bfloat16*>(wv), bv, reinterpret_cast<__nv_bfloat16*>(out),
                             static_cast<long>(ld_out), heads, S);
  } else {
    e = cudaFuncSetAttribute(pool_probe_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess)
      e = cudaLaunchKernelEx(&lc.cfg, pool_probe_kernel<__half>, reinterpret_cast<const __half*>(tokens), static_cast<long>(ld), u,
                             reinterpret_cast<const __half*>(wv), bv, reinterpret_cast<__half*>(out), static_cast<long>(ld_out),
                             heads, S);
  }
  (void)e;
  return check_launch("pool_probe");
}

int pool_attention(cudaStream_t stream, int dtype, const void* kv, int ld_kv, const float* q, void* out,
                   int ld_out, int frames, int heads, int S) {
  if (frames <= 0) return 0;
  if (S > kPoolMaxS) { set_error("pool_attention: S=%d exceeds %d", S, kPoolMaxS); return -1; }
  if (dtype != kBF16 && dtype != kF16) { set_error("pool_attention: dtype must be bf16/f16"); return -1; }
  if ((ld_kv % 8) || (reinterpret_cast<uintptr_t>(kv) & 15)) { set_error("pool_attention: 16-byte aligned rows required"); return -1; }
  const long tasks = static_cast<long>(frames) * heads;
  const long blocks = tasks;
  ProfScope ps(stream, kProfPoolAttn, 4.0 * frames * heads * static_cast<double>(S) * kHd,
               2.0 * frames * heads * kHd * 2.0 * S);
  LaunchCfg lc(dim3(static_cast<unsigned>(blocks)), dim3(128), 0, stream);
  if (dtype == kBF16) cudaLaunchKernelEx(&lc.cfg, pool_attn_kernel<__nv_bfloat16>, reinterpret_cast<const __nv_bfloat16*>(kv),
                                         static_cast<long>(ld_kv), q, reinterpret_cast<__nv_bfloat16*>(out),
                                         static_cast<long>(ld_out), frames, heads, S);
  else cudaLaunchKernelEx(&lc.cfg, pool_attn_kernel<__half>, reinterpret_cast<const __half*>(kv), static_cast<long>(ld_kv), q,
                          reinterpret_cast<__half*>(out), static_cast<long>(ld_out), frames, heads, S);
  return check_launch("pool_attention");
}

}  // namespace sf
