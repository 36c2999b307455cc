// runtime.cu — host-side schedule of the StreamFormer encoder behind the C ABI
// (include/streamformer_b200.h): weight packing, workspace carving, the per-layer kernel sequence
// of TimesformerLayerSigLIP.forward (models/modeling_timesformer_siglip.py:900-1004), the embedding
// front end (:413-457), the final norm + pooling head (:1330-1346, 1141-1154) and the streaming
// KV cache (downstream/VideoQA/llava/model/multimodal_encoder/timesformer_encoder.py:307-375,
// 491-560, 1340-1349).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/streamformer_b200.h"
#include "sf_kernels.h"
#include "sf_ptx.cuh"

using namespace sf;

#define SF_CHECK(expr)                         \
  do {                                         \
    int _rc = (expr);                          \
    if (_rc != 0) return _rc;                  \
  } while (0)

#define SF_CUDA(expr)                                                              \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SF_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

namespace {

// Makes the context's device current for the duration of a C-ABI call and restores the caller's
// device afterwards (the host may drive several GPUs from one thread).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched && prev >= 0) cudaSetDevice(prev);
  }
};

// ------------------------------------------------------------------ bind-time fp32 helper kernels
__global__ void lora_merge_kernel(float* __restrict__ W, const float* __restrict__ A,
                                  const float* __restrict__ Bm, int O, int I, int R) {
  // W[o,i] += sum_r Bm[o,r] * A[r,i]      (…siglip.py:653-654, 749-751)
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long>(O) * I) return;
  const int o = static_cast<int>(idx / I), i = static_cast<int>(idx % I);
  float acc = 0.f;
  for (int r = 0; r < R; ++r) acc += Bm[o * R + r] * A[r * I + i];
  W[idx] += acc;
}
__global__ void matmul_nn_kernel(float* __restrict__ C, const float* __restrict__ A,
                                 const float* __restrict__ Bm, int M, int N, int K) {
  // C[m,n] = sum_k A[m,k] * Bm[k,n]   (one-off weight folding; correctness over speed)
  __shared__ float as[16][17], bs[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    as[ty][tx] = (m < M && k0 + tx < K) ? A[static_cast<long>(m) * K + k0 + tx] : 0.f;
    bs[ty][tx] = (k0 + ty < K && n < N) ? Bm[static_cast<long>(k0 + ty) * N + n] : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += as[ty][k] * bs[k][tx];
    __syncthreads();
  }
  if (m < M && n < N) C[static_cast<long>(m) * N + n] = acc;
}
// u[h][k] = sum_d Wk[h*64 + d][k] * q[h*64 + d]: the pooling probe pushed through the key projection
__global__ void head_u_kernel(float* __restrict__ u, const float* __restrict__ Wk, const float* __restrict__ q, int H, int D) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * D) return;
  const int h = idx / D, k = idx % D;
  float acc = 0.f;
  for (int d = 0; d < 64; ++d) acc += Wk[static_cast<long>(h * 64 + d) * D + k] * q[h * 64 + d];
  u[idx] = acc;
}
__global__ void matvec_kernel(float* __restrict__ y, const float* __restrict__ W,
                              const float* __restrict__ x, const float* __restrict__ b, int O, int I,
                              float scale) {
  // y[o] = scale * (sum_i W[o,i] x[i] + b[o]); one warp per output
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= O) return;
  float acc = 0.f;
  for (int i = lane; i < I; i += 32) acc += W[static_cast<long>(o) * I + i] * x[i];
  acc = warp_sum(acc);
  if (lane == 0) y[o] = scale * (acc + (b ? b[o] : 0.f));
}

// LayerNorm folded into the next projection (bind time):  LN(x) . W^T + b  ==
//   rstd * (x . W'^T - mean * colsum) + b'   with W'[n,k] = W[n,k] * gamma[k] (rounded to the
//   activation dtype), colsum[n] = sum_k W'[n,k] (of the ROUNDED values, so the mean term cancels
//   exactly against what the tensor cores accumulate) and b'[n] = b[n] + sum_k W[n,k] * beta[k].
// One warp per output row n.
template <typename T>
__global__ void ln_fold_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ bias,
                               T* __restrict__ Wp, float* __restrict__ colsum, float* __restrict__ biasp,
                               int O, int I) {
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= O) return;
  float cs = 0.f, bs = 0.f;
  for (int i = lane; i < I; i += 32) {
    const float w = W[static_cast<long>(o) * I + i];
    const T wr = static_cast<T>(w * gamma[i]);
    Wp[static_cast<long>(o) * I + i] = wr;
    cs += static_cast<float>(wr);
    bs += w * beta[i];
  }
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  if (lane == 0) {
    colsum[o] = cs;
    biasp[o] = (bias ? bias[o] : 0.f) + bs;
  }
}

struct Bump {
  uint8_t* base = nullptr;
  size_t size = 0, off = 0;
  bool overflow = false;
  void* take(size_t bytes) {
    const size_t a = (off + 255) & ~static_cast<size_t>(255);
    if (base == nullptr || a + bytes > size) {
      overflow = true;
      off = a + bytes;
      return nullptr;
    }
    off = a + bytes;
    return base + a;
  }
};

struct LayerW {
  void *t_qkv_w = nullptr, *t_out_w = nullptr, *t_dense_w = nullptr, *t_fold_w = nullptr;
  void *s_qkv_w = nullptr, *s_out_w = nullptr, *fc1_w = nullptr, *fc2_w = nullptr;
  float *t_qkv_b = nullptr, *t_out_b = nullptr, *t_dense_b = nullptr, *t_fold_b = nullptr;
  float *s_qkv_b = nullptr, *s_out_b = nullptr, *fc1_b = nullptr, *fc2_b = nullptr;
  float *ln_t_g = nullptr, *ln_t_b = nullptr, *ln_s_g = nullptr, *ln_s_b = nullptr;
  float *ln_a_g = nullptr, *ln_a_b = nullptr;
  // column sums of the gamma-scaled matrices (folded LayerNorms); t_qkv_b / s_qkv_b / fc1_b hold b + W.beta
  float *t_qkv_cs = nullptr, *s_qkv_cs = nullptr, *fc1_cs = nullptr;
  float* gate = nullptr;
};

}  // namespace

struct sf_ctx {
  sf_config cfg;
  int device = 0;
  int D = 0, H = 0, I = 0, L = 0, S0 = 0, Kp = 0;  // hidden, heads, mlp, layers, default sites, patch K
  bool bound = false;
  // which parameter groups the last sf_bind_weights call found (a stand-alone TimesformerEncoder /
  // embeddings / pooling head binds only its own tensors)
  bool have_embed = false, have_post = false, have_head = false;
  std::vector<char> have_layer;
  // uint8 pixels: (x / 255 - mean[c]) / std[c] inside im2col (Normalize(0.5, 0.5) of the reference's
  // loaders: extract_oad_feature.py:42-48, datasets/kinetics_sparse.py:110-118)
  float pix_mean[4] = {0.5f, 0.5f, 0.5f, 0.5f}, pix_std[4] = {0.5f, 0.5f, 0.5f, 0.5f};
  int pos_epoch = 0;           // bumped whenever pos_alt is re-allocated (captured graphs hold its address)
  // dual-stream forward: the batch runs as two halves on two streams so that one half's kernel fill / drain
  // overlaps the other half's steady state (see forward_dual)
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<LayerW> layers;
  void* patch_w = nullptr; float* patch_b = nullptr;
  float* pos = nullptr;        // [S0, D]
  float* time_emb = nullptr;   // [num_frames, D]
  float* pos_alt = nullptr;    // user-provided interpolated table
  int pos_alt_S = 0;
  float *post_g = nullptr, *post_b = nullptr;
  void *head_kv_w = nullptr, *head_out_w = nullptr, *head_fc1_w = nullptr, *head_fc2_w = nullptr;
  float *head_kv_b = nullptr, *head_out_b = nullptr, *head_fc1_b = nullptr, *head_fc2_b = nullptr;
  float *head_q = nullptr, *head_ln_g = nullptr, *head_ln_b = nullptr;
  float* head_u = nullptr;     // [heads, D]: W_k,h^T q_h (the probe pushed through the key projection)
  // GEMM-chain dependency counters live in the caller's workspace: zeroed when first seen (the
  // chain kernels leave them zero)
  void* chain_ctr_seen = nullptr;
  size_t chain_ctr_bytes = 0;
};

// One captured streaming step (sf_forward_stream under a CUDA graph): fixed shapes and buffers, the
// position in the stream is read by the kernels from sf_kv::d_seen.
struct StreamGraph {
  cudaGraphExec_t exec = nullptr;
  int B = 0, T = 0, H = 0, W = 0, pix_dtype = 0;
  bool pooler = false, hidden = false;
  void* ws = nullptr; size_t ws_bytes = 0;
  uint8_t* stage = nullptr;                 // pixels | last_hidden | pooler | L+1 hidden states (graph-owned addresses)
  size_t pix_bytes = 0, lh_bytes = 0, pool_bytes = 0, hs_bytes = 0;
  std::vector<void*> hs_ptrs;
  int pos_epoch = 0;                        // sf_ctx::pos_epoch at capture time
  int calls = 0;                            // eager calls seen with this key (capture happens on the 2nd)
  int kernels = 0;                          // kernel nodes in the graph (for sf_launch_count)
};

struct sf_kv {
  sf_ctx* ctx = nullptr;
  int B = 0, S = 0, cap = 0, seen = 0, horizon = 0;
  int* d_seen = nullptr;        // device copy of `seen`, advanced by the captured graph itself
  bool seen_dirty = true;       // host `seen` changed outside a graph launch -> re-upload before the next one
  bool graphs_disabled = false; // capture failed once: stay on direct launches
  long graph_launches = 0;      // steps served by a graph replay
  cudaStream_t cap_stream = nullptr;  // capture happens here (the caller's stream may be the legacy default stream, which cannot capture)
  std::vector<StreamGraph> graphs;
  uint8_t* mem = nullptr;
  size_t layer_stride = 0;  // bytes between layers (K then V inside)
  size_t kv_bytes = 0;      // bytes of one K (or V) block
  void* k(int l) const { return mem + l * layer_stride; }
  void* v(int l) const { return mem + l * layer_stride + kv_bytes; }
};

namespace {

struct WsPlan {
  void *qkv, *ctx, *tmp, *mlp;
  void* chain_ctr = nullptr;   // dependency counters of the GEMM chains (zero between launches)
  // partial row statistics (sum, sumsq) feeding the folded LayerNorms: [0] after the temporal
  // branch, [1] after the spatial branch, [2] at the layer boundary (after the MLP / the embedding)
  float2* stats[3];
};

size_t stats_bytes(const sf_ctx* c, long M) {
  return static_cast<size_t>(gemm_stats_parts_max(c->D)) * static_cast<size_t>(M) * sizeof(float2);
}

size_t layer_ws_bytes(const sf_ctx* c, long M) {
  const size_t es = 2;
  const size_t D = c->D, I = c->I;
  // qkv[M,3D] ctx[M,D] tmp[M,D] mlp[M,I] stats[3]  (+256 B alignment each)
  return static_cast<size_t>(M) * (3 * D + D + D + I) * es + 3 * stats_bytes(c, M) + gemm_chain_counter_bytes(static_cast<int>(M)) +
         8 * 256;
}

int carve_layer_ws(const sf_ctx* c, long M, Bump& b, WsPlan& p) {
  const size_t es = 2, D = c->D, I = c->I;
  p.qkv = b.take(M * 3 * D * es);
  p.ctx = b.take(M * D * es);
  p.tmp = b.take(M * D * es);
  p.mlp = b.take(M * I * es);
  for (int i = 0; i < 3; ++i) p.stats[i] = static_cast<float2*>(b.take(stats_bytes(c, M)));
  p.chain_ctr = b.take(gemm_chain_counter_bytes(static_cast<int>(M)));
  if (b.overflow) {
    set_error("workspace too small: need at least %zu bytes, have %zu", b.off, b.size);
    return SF_ERR_WORKSPACE;
  }
  return 0;
}

// Both helpers describe GEMMs against packed weights of this context (sf_bind_weights synchronises the stream before it
// returns, so no kernel in flight writes them): the W operand may be fetched ahead of the PDL dependency wait.
GemmEpilogue epi_bias(const float* bias) {
  GemmEpilogue e;
  e.bias = bias;
  e.w_static = true;
  return e;
}

GemmEpilogue epi_ln(const float* bias, const float* colsum, const float2* stats, int parts, float eps) {
  GemmEpilogue e;
  e.bias = bias;
  e.ln_colsum = colsum; e.ln_stats = stats; e.ln_parts = parts; e.ln_eps = eps;
  e.w_static = true;
  return e;
}

// Runs dependent GEMMs over the same rows: as one persistent chain launch when the geometry allows
// it (CTA-pair tiles for every call), else one launch per call.  Row statistics handed from call
// i-1 to call i (stats_out -> ln_stats) use the partial count of whichever path runs.
int run_gemms(sf_ctx* c, cudaStream_t st, GemmCall* calls, int n, const WsPlan& w, int* last_stats_parts = nullptr) {
  const int dt = c->cfg.dtype;
  const bool chain = n >= 2 && w.chain_ctr && gemm_chain_supported(dt, calls, n);
  if (last_stats_parts)   // partial count of the row statistics the LAST call leaves in its stats_out
    *last_stats_parts = chain ? gemm_chain_stats_parts(calls, n, calls[n - 1].N) : gemm_stats_parts(calls[n - 1].M, calls[n - 1].N);
  for (int i = 1; i < n; ++i) {
    if (calls[i].epi.ln_stats && calls[i].epi.ln_stats == calls[i - 1].epi.stats_out)
      calls[i].epi.ln_parts = chain ? gemm_chain_stats_parts(calls, n, calls[i - 1].N) : gemm_stats_parts(calls[i - 1].M, calls[i - 1].N);
  }
  if (chain) {
    const size_t bytes = gemm_chain_counter_bytes(calls[0].M);
    if (c->chain_ctr_seen != w.chain_ctr || c->chain_ctr_bytes != bytes) {
      SF_CUDA(cudaMemsetAsync(w.chain_ctr, 0, bytes, st));
      c->chain_ctr_seen = w.chain_ctr;
      c->chain_ctr_bytes = bytes;
    }
    return gemm_chain(st, dt, calls, n, w.chain_ctr);
  }
  for (int i = 0; i < n; ++i) {
    GemmCall& g = calls[i];
    g.epi.w_static = true;     // W is a packed weight; sf_bind_weights synchronises the stream before it returns
    SF_CHECK(gemm(st, dt, g.A, g.lda, g.W, g.ldw, g.out, g.ldo, g.M, g.N, g.K, g.epi));
  }
  return 0;
}

GemmCall make_call(const void* A, int lda, const void* W, int ldw, void* out, int ldo, long M, int N, int K,
                   const GemmEpilogue& e) {
  GemmCall g;
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.out = out; g.ldo = ldo;
  g.M = static_cast<int>(M); g.N = N; g.K = K; g.epi = e;
  return g;
}

// One divided space-time block.  Rows stay in the residual stream's (b,n,t) order throughout: the
// temporal branch sees its T frames of a site as consecutive rows, the spatial attention reads the
// S tokens of a frame in place with a row stride of T (no permute copies, reference :962-991).
// The three LayerNorms are folded into the QKV / fc1 GEMMs (GemmEpilogue::ln_stats); the row
// statistics they need are by-products of the epilogues that wrote the rows:
//   st_in (x_in) -> temporal QKV ; w.stats[0] (after temporal) -> spatial QKV ;
//   w.stats[1] (after spatial) -> fc1 ; st_out (after the MLP) -> the next layer.
// GEMMs that follow each other without an attention kernel in between run as chains:
//   [temporal out-proj(.temporal_dense) + gated residual -> spatial QKV]
//   [spatial out-proj + residual -> fc1 + GELU -> fc2 + residual -> the NEXT layer's temporal QKV]
// qkv_ready: this layer's temporal QKV was produced by the previous layer's chain;
// next: the following layer (its temporal QKV is appended to this layer's second chain), or null.
int run_layer(sf_ctx* c, cudaStream_t st, int l, const void* x_in, void* x_out, int B, int T, int S,
              sf_kv* kv, float* probs, const WsPlan& w, const float2* st_in, int parts_in, float2* st_out,
              const int* seen_dev = nullptr, bool qkv_ready = false, const LayerW* next = nullptr,
              bool* next_qkv_done = nullptr, int* st_out_parts = nullptr) {
  const LayerW& lw = c->layers[l];
  const int D = c->D, I = c->I, H = c->H, dt = c->cfg.dtype;
  const long M = static_cast<long>(B) * S * T;
  const float eps = c->cfg.layer_norm_eps;
  const float scale = 0.125f;  // head_dim**-0.5, head_dim == 64 (…siglip.py:512, 628)
  if (next_qkv_done) *next_qkv_done = false;

  // ---- temporal branch (…siglip.py:937-958): rows (b,n,t), T innermost => sites are contiguous
  {
  PhaseScope phase(st, kPhaseAttnBlock);   // the space-time attention block incl. its projections
  if (!qkv_ready)
    SF_CHECK(gemm(st, dt, x_in, D, lw.t_qkv_w, D, w.qkv, 3 * D, M, 3 * D, D,
                  epi_ln(lw.t_qkv_b, lw.t_qkv_cs, st_in, parts_in, eps)));
  if (kv && temporal_decode_supported(kv->cap, T)) {
    SF_CHECK(temporal_decode(st, dt, w.qkv, 3 * D, kv->k(l), kv->v(l), kv->cap, w.ctx, D, B * S, H, kv->seen, scale, seen_dev));
  } else if (kv) {
    SF_CHECK(kv_append(st, dt, w.qkv, 3 * D, kv->k(l), kv->v(l), kv->cap, B * S, H, T, kv->seen, seen_dev));
    SF_CHECK(temporal_attention(st, dt, w.qkv, 3 * D, kv->k(l), kv->v(l), kv->cap, w.ctx, D, B * S, H, T,
                                kv->seen + T, kv->seen, c->cfg.causal_temporal, scale, seen_dev));
  } else {
    SF_CHECK(temporal_attention(st, dt, w.qkv, 3 * D, nullptr, nullptr, 0, w.ctx, D, B * S, H, T, T, 0,
                                c->cfg.causal_temporal, scale));
  }
  {
    // [temporal out-proj (. temporal_dense) + gated residual -> spatial QKV (…siglip.py:954-958, 960-974)]
    GemmCall calls[3];
    int n = 0;
    GemmEpilogue e;
    e.residual = x_in; e.ldr = D; e.gate = lw.gate;
    e.stats_out = w.stats[0];
    if (c->cfg.fold_temporal_proj) {
      e.bias = lw.t_fold_b;
      calls[n++] = make_call(w.ctx, D, lw.t_fold_w, D, x_out, D, M, D, D, e);
    } else {
      calls[n++] = make_call(w.ctx, D, lw.t_out_w, D, w.tmp, D, M, D, D, epi_bias(lw.t_out_b));
      e.bias = lw.t_dense_b;
      calls[n++] = make_call(w.tmp, D, lw.t_dense_w, D, x_out, D, M, D, D, e);
    }
    calls[n++] = make_call(x_out, D, lw.s_qkv_w, D, w.qkv, 3 * D, M, 3 * D, D,
                           epi_ln(lw.s_qkv_b, lw.s_qkv_cs, w.stats[0], 0, eps));
    SF_CHECK(run_gemms(c, st, calls, n, w));
  }
  // ---- spatial branch (…siglip.py:960-996)
  SF_CHECK(spatial_attention(st, dt, w.qkv, 3 * D, w.ctx, D, B * T, H, S, T, scale, probs));
  }
  // ---- [spatial out-proj + residual -> MLP (…siglip.py:993-1000) -> next layer's temporal QKV]
  GemmCall calls[4];
  int n = 0;
  int s_out_parts = 0;
  {
    GemmEpilogue e = epi_bias(lw.s_out_b);
    e.residual = x_out; e.ldr = D;
    e.stats_out = w.stats[1];
    calls[n++] = make_call(w.ctx, D, lw.s_out_w, D, x_out, D, M, D, D, e);
  }
  if (phase_prof_enabled()) {
    // per-phase timing: the out-projection belongs to the attention block, so it runs on its own
    // inside that phase instead of heading the MLP chain (a slightly less fused schedule than production)
    PhaseScope attn_tail(st, kPhaseAttnBlock);
    SF_CHECK(run_gemms(c, st, calls, 1, w, &s_out_parts));
    n = 0;
  }
  PhaseScope phase(st, kPhaseMlp);
  {
    GemmEpilogue e = epi_ln(lw.fc1_b, lw.fc1_cs, w.stats[1], s_out_parts, eps);
    e.act = c->cfg.hidden_act;
    calls[n++] = make_call(x_out, D, lw.fc1_w, D, w.mlp, I, M, I, D, e);
  }
  {
    GemmEpilogue e = epi_bias(lw.fc2_b);
    e.residual = x_out; e.ldr = D;
    e.stats_out = st_out;
    calls[n++] = make_call(w.mlp, I, lw.fc2_w, I, x_out, D, M, D, I, e);
  }
  if (next && st_out) {
    calls[n] = make_call(x_out, D, next->t_qkv_w, D, w.qkv, 3 * D, M, 3 * D, D,
                         epi_ln(next->t_qkv_b, next->t_qkv_cs, st_out, 0, eps));
    if (gemm_chain_supported(dt, calls, n + 1)) {
      ++n;
      if (next_qkv_done) *next_qkv_done = true;
      return run_gemms(c, st, calls, n, w);     // (st_out is consumed inside the chain)
    }
  }
  return run_gemms(c, st, calls, n, w, st_out_parts);
}

int run_embed(sf_ctx* c, cudaStream_t st, const void* pixels, int pix_dtype, int B, int T, int Hh, int Ww,
              int time_off, int time_total, void* x_out, void* patches, float2* stats_out,
              const int* time_off_dev = nullptr, int time_horizon = 0) {
  const int P = c->cfg.patch_size, C = c->cfg.num_channels, D = c->D, dt = c->cfg.dtype;
  const int S = (Hh / P) * (Ww / P);
  const float* pos = nullptr;
  if (S == c->S0 && Hh == Ww) pos = c->pos;                      // …siglip.py:384-385
  else if (c->pos_alt && c->pos_alt_S == S) pos = c->pos_alt;
  else {
    set_error("resolution %dx%d (%d patches) needs an interpolated position table: call sf_set_pos_embed first",
              Hh, Ww, S);
    return SF_ERR_STATE;
  }
  const long M = static_cast<long>(B) * T * S;
  if (!c->have_embed) { set_error("embedding weights are not bound to this context"); return SF_ERR_STATE; }
  SF_CHECK(im2col_patches(st, pix_dtype, pixels, dt, patches, B * T, C, Hh, Ww, P, c->pix_mean, c->pix_std));
  GemmEpilogue e = epi_bias(c->patch_b);
  e.row_map = kRowBTNtoBNT; e.T = T; e.S = S;
  e.pos = pos;
  e.time_emb = c->time_emb; e.time_len = c->cfg.num_frames; e.time_total = time_total; e.time_off = time_off;
  e.time_off_dev = time_off_dev; e.time_horizon = time_horizon;
  e.stats_out = stats_out;
  return gemm(st, dt, patches, c->Kp, c->patch_w, c->Kp, x_out, D, M, D, c->Kp, e);
}

size_t head_ws_bytes(const sf_ctx* c, long frames, long S) {
  const size_t es = 2, D = c->D, I = c->I;
  return static_cast<size_t>(frames) * S * 2 * D * es + static_cast<size_t>(frames) * (3 * D + I) * es + 6 * 256;
}

int run_head(sf_ctx* c, cudaStream_t st, const void* tokens, int frames, int S, void* pooled, Bump& b) {
  const int D = c->D, I = c->I, H = c->H, dt = c->cfg.dtype;
  const size_t es = 2;
  const long M = static_cast<long>(frames) * S;
  if (!c->have_head) { set_error("pooling-head weights are not bound to this context"); return SF_ERR_STATE; }
  void* kvbuf = b.take(M * 2 * D * es);
  void* pc = b.take(static_cast<size_t>(frames) * D * es);
  void* r = b.take(static_cast<size_t>(frames) * D * es);
  void* lnr = b.take(static_cast<size_t>(frames) * D * es);
  void* hbuf = b.take(static_cast<size_t>(frames) * I * es);
  if (b.overflow) {
    set_error("workspace too small for the pooling head: need %zu bytes, have %zu", b.off, b.size);
    return SF_ERR_WORKSPACE;
  }
  // K/V projection of every token (in_proj rows D..3D) + probe attention (…siglip.py:1146-1148), or —
  // opt-in, SF_HEAD_COLLAPSE=1 — the same attention with the K/V projections collapsed into it
  // (pool_probe in attention.cu: exact algebra, 59 GFLOP and 77 MB less per cfg2 step, but its CUDA-core
  // implementation is shared-memory-bandwidth bound and currently SLOWER: 187 vs 87 us; it needs the
  // mma.sync formulation before it can become the default)
  static const bool collapse = [] { const char* e = getenv("SF_HEAD_COLLAPSE"); return e && e[0] == '1'; }();
  if (collapse && H <= 16 && D % 256 == 0 && D <= 1024 && S <= 2048) {
    SF_CHECK(pool_probe(st, dt, tokens, D, c->head_u, static_cast<const uint8_t*>(c->head_kv_w) + static_cast<size_t>(D) * D * es,
                        c->head_kv_b + D, pc, D, frames, H, S));
  } else {
    SF_CHECK(gemm(st, dt, tokens, D, c->head_kv_w, D, kvbuf, 2 * D, M, 2 * D, D, epi_bias(c->head_kv_b)));
    SF_CHECK(pool_attention(st, dt, kvbuf, 2 * D, c->head_q, pc, D, frames, H, S));
  }
  SF_CHECK(gemm(st, dt, pc, D, c->head_out_w, D, r, D, frames, D, D, epi_bias(c->head_out_b)));
  // r + MLP(LN(r)) (…siglip.py:1150-1152)
  SF_CHECK(layernorm(st, dt, r, D, c->head_ln_g, c->head_ln_b, c->cfg.layer_norm_eps, lnr, D, frames, D,
                     kRowIdentity, 1, 1));
  {
    GemmEpilogue e = epi_bias(c->head_fc1_b);
    e.act = c->cfg.hidden_act;
    SF_CHECK(gemm(st, dt, lnr, D, c->head_fc1_w, D, hbuf, I, frames, I, D, e));
  }
  {
    GemmEpilogue e = epi_bias(c->head_fc2_b);
    e.residual = r; e.ldr = D;
    SF_CHECK(gemm(st, dt, hbuf, I, c->head_fc2_w, I, pooled, D, frames, D, I, e));
  }
  return 0;
}

int check_shape(const sf_ctx* c, int B, int T, int Hh, int Ww) {
  if (!c->bound) { set_error("weights not bound: call sf_bind_weights first"); return SF_ERR_STATE; }
  const int P = c->cfg.patch_size;
  if (B <= 0 || T <= 0 || Hh <= 0 || Ww <= 0 || (Hh % P) || (Ww % P)) {
    set_error("bad input shape B=%d T=%d H=%d W=%d (patch %d)", B, T, Hh, Ww, P);
    return SF_ERR_INVALID;
  }
  return 0;
}

int check_groups(const sf_ctx* c, bool embed, int l0, int l1, bool post, bool head) {
  if (embed && !c->have_embed) { set_error("embedding weights are not bound to this context"); return SF_ERR_STATE; }
  for (int l = l0; l < l1; ++l)
    if (l < 0 || l >= c->L || !c->have_layer[l]) { set_error("weights of encoder.layer.%d are not bound to this context", l); return SF_ERR_STATE; }
  if (post && !c->have_post) { set_error("post_layernorm weights are not bound to this context"); return SF_ERR_STATE; }
  if (head && !c->have_head) { set_error("pooling-head weights are not bound to this context"); return SF_ERR_STATE; }
  return 0;
}

int g_dual_stream_opt = -1;   // sf_set_option("dual_stream", v): -1 environment default (SF_DUAL_STREAM, off), 0 off, 1 on
bool dual_stream_enabled() {
  // Measured on B200 (round 2, cfg2): 19 518 frames/s with the two-stream schedule vs 19 605 without — the persistent
  // GEMMs already occupy every SM, so the second stream's kernels only start as the first's drain and each half pays
  // its own fill / drain: no gain, hence opt-in.
  static const bool env_on = [] { const char* e = getenv("SF_DUAL_STREAM"); return e && e[0] == '1'; }();
  return g_dual_stream_opt < 0 ? env_on : g_dual_stream_opt != 0;
}

// One-shot forward of an even batch as TWO half batches on two streams (the caller's and a context-owned one),
// launches issued alternately layer by layer.  Clips are independent, so the halves share nothing but
// the weights.  Every GEMM of a layer depends on the previous kernel of its own half, so a single stream leaves the
// machine partly idle while each persistent kernel fills (first operand tiles in flight) and drains (last tile's
// epilogue, SMs without a tile in the last wave) — ~12 us of a 30..100 us kernel; with two independent chains the SMs one
// half's kernel releases are picked up by the other half's next kernel.
int forward_dual(sf_ctx* c, cudaStream_t st0, const void* pixels, int pix_dtype, int B, int T, int Hh, int Ww, void* last_hidden,
                 void* pooler, void* ws, size_t ws_bytes) {
  const int P = c->cfg.patch_size, D = c->D, C = c->cfg.num_channels;
  const int S = (Hh / P) * (Ww / P);
  const int Bh = B / 2;
  const long Mh = static_cast<long>(Bh) * T * S;
  if (!c->aux_stream) {
    SF_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    SF_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    SF_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  }
  cudaStream_t sts[2] = {st0, c->aux_stream};
  SF_CUDA(cudaEventRecord(c->ev_fork, st0));
  SF_CUDA(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
  Bump b;
  b.base = static_cast<uint8_t*>(ws);
  b.size = ws_bytes;
  void* x[2];
  WsPlan w[2];
  uint8_t* region_end[2];
  for (int h = 0; h < 2; ++h) {
    x[h] = b.take(Mh * D * 2);
    SF_CHECK(carve_layer_ws(c, Mh, b, w[h]));
    w[h].chain_ctr = nullptr;                     // the chained schedule keeps per-context counter state: single-stream only
    region_end[h] = static_cast<uint8_t*>(ws) + (b.off < ws_bytes ? b.off : ws_bytes);
  }
  if (b.overflow) { set_error("workspace too small"); return SF_ERR_WORKSPACE; }
  const size_t pix_half = static_cast<size_t>(Bh) * T * C * Hh * Ww * dtype_size(pix_dtype);
  int parts_d[2];
  bool qkv_ready[2] = {false, false};
  for (int h = 0; h < 2; ++h) {
    SF_CHECK(run_embed(c, sts[h], static_cast<const uint8_t*>(pixels) + h * pix_half, pix_dtype, Bh, T, Hh, Ww, 0, T, x[h], w[h].mlp,
                       w[h].stats[2]));
    parts_d[h] = gemm_stats_parts(static_cast<int>(Mh), D);
  }
  for (int l = 0; l < c->L; ++l) {
    const LayerW* next = (l + 1 < c->L) ? &c->layers[l + 1] : nullptr;
    for (int h = 0; h < 2; ++h) {
      bool next_done = false;
      SF_CHECK(run_layer(c, sts[h], l, x[h], x[h], Bh, T, S, nullptr, nullptr, w[h], w[h].stats[2], parts_d[h], w[h].stats[2], nullptr,
                         qkv_ready[h], next, &next_done, &parts_d[h]));
      qkv_ready[h] = next_done;
    }
  }
  for (int h = 0; h < 2; ++h) {
    uint8_t* lh = static_cast<uint8_t*>(last_hidden) + static_cast<size_t>(h) * Mh * D * 2;
    SF_CHECK(layernorm(sts[h], c->cfg.dtype, x[h], D, c->post_g, c->post_b, c->cfg.layer_norm_eps, lh, D, Mh, D,
                       T > 1 ? kRowBNTtoBTN : kRowIdentity, T, S));
    if (pooler) {
      Bump hb;
      hb.base = static_cast<uint8_t*>(w[h].qkv);
      hb.size = region_end[h] - static_cast<uint8_t*>(w[h].qkv);
      SF_CHECK(run_head(c, sts[h], lh, Bh * T, S, static_cast<uint8_t*>(pooler) + static_cast<size_t>(h) * Bh * T * D * 2, hb));
    }
  }
  SF_CUDA(cudaEventRecord(c->ev_join, c->aux_stream));
  SF_CUDA(cudaStreamWaitEvent(st0, c->ev_join, 0));
  return 0;
}

int forward_impl(sf_ctx* c, cudaStream_t st, sf_kv* kv, const void* pixels, int pix_dtype, int B, int T,
                 int Hh, int Ww, void* last_hidden, void* pooler, void* const* hidden_states,
                 void* const* attentions, void* ws, size_t ws_bytes, bool dev_seen = false) {
  // dev_seen: being captured into a streaming CUDA graph — the kernels take the stream position
  // from kv->d_seen at run time and the host-side counter is advanced by the caller
  SF_CHECK(check_shape(c, B, T, Hh, Ww));
  SF_CHECK(check_groups(c, true, 0, c->L, true, pooler != nullptr));
  const int P = c->cfg.patch_size, D = c->D;
  const int S = (Hh / P) * (Ww / P);
  const long M = static_cast<long>(B) * T * S;
  if (!kv && !hidden_states && !attentions && B >= 2 && (B % 2) == 0 && dual_stream_enabled() && !prof_enabled() &&
      !phase_prof_enabled() && M >= 2 * 4096) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap == cudaStreamCaptureStatusNone)
      return forward_dual(c, st, pixels, pix_dtype, B, T, Hh, Ww, last_hidden, pooler, ws, ws_bytes);
  }
  int time_off = 0, time_total = T;
  if (kv) {
    if (kv->B != B || kv->S != S) {
      set_error("KV cache was created for B=%d S=%d, got B=%d S=%d", kv->B, kv->S, B, S);
      return SF_ERR_INVALID;
    }
    if (kv->seen + T > kv->cap) {
      set_error("KV cache overflow: %d cached + %d new frames > capacity %d", kv->seen, T, kv->cap);
      return SF_ERR_STATE;
    }
    time_off = kv->seen;
    time_total = kv->horizon > 0 ? kv->horizon : kv->seen + T;
    if (time_total < kv->seen + T) time_total = kv->seen + T;
  }
  Bump b;
  b.base = static_cast<uint8_t*>(ws);
  b.size = ws_bytes;
  void* x = nullptr;
  if (!hidden_states) x = b.take(M * D * 2);
  WsPlan w;
  SF_CHECK(carve_layer_ws(c, M, b, w));
  if (b.overflow) { set_error("workspace too small"); return SF_ERR_WORKSPACE; }
  // the im2col operand aliases the (not yet used) MLP buffer: Kp <= I is checked at create time
  void* cur = hidden_states ? hidden_states[0] : x;
  {
    PhaseScope phase(st, kPhaseEmbed);
    SF_CHECK(run_embed(c, st, pixels, pix_dtype, B, T, Hh, Ww, time_off, time_total, cur, w.mlp, w.stats[2],
                       dev_seen ? kv->d_seen : nullptr, dev_seen ? kv->horizon : 0));
  }
  int parts_d = gemm_stats_parts(static_cast<int>(M), D);   // partials left by the embedding GEMM
  bool qkv_ready = false;
  for (int l = 0; l < c->L; ++l) {
    void* nxt = hidden_states ? hidden_states[l + 1] : cur;
    // the phase profiler times the attention block of each layer on its own: keep the next layer's
    // QKV out of this layer's MLP chain then
    const LayerW* next = (l + 1 < c->L && !phase_prof_enabled()) ? &c->layers[l + 1] : nullptr;
    bool next_done = false;
    SF_CHECK(run_layer(c, st, l, cur, nxt, B, T, S, kv, attentions ? static_cast<float*>(attentions[l]) : nullptr, w,
                       w.stats[2], parts_d, w.stats[2], dev_seen ? kv->d_seen : nullptr, qkv_ready, next, &next_done,
                       &parts_d));
    qkv_ready = next_done;
    cur = nxt;
  }
  if (kv && !dev_seen) {
    kv->seen += T;
    kv->seen_dirty = true;
  }
  // post_layernorm, written straight in (b,t,n) order == last_hidden_state (…siglip.py:1330-1346)
  PhaseScope phase(st, kPhaseHead);
  SF_CHECK(layernorm(st, c->cfg.dtype, cur, D, c->post_g, c->post_b, c->cfg.layer_norm_eps, last_hidden, D, M,
                     D, T > 1 ? kRowBNTtoBTN : kRowIdentity, T, S));
  if (pooler) {
    // head scratch aliases the layer scratch (all layers are done)
    Bump hb;
    hb.base = static_cast<uint8_t*>(w.qkv);
    hb.size = static_cast<uint8_t*>(ws) + ws_bytes - static_cast<uint8_t*>(w.qkv);
    SF_CHECK(run_head(c, st, last_hidden, B * T, S, pooler, hb));
  }
  return 0;
}

// ------------------------------------------------------------------ streaming under a CUDA graph
// BASELINE config 3 (64 appends of one frame at B=4) is launch-bound when every step re-issues its
// ~117 kernels from the host (~10 us of CPU per launch incl. two tensor-map encodes).  A streaming
// step is therefore captured ONCE per (shape, buffers) into a CUDA graph whose kernels read the
// stream position from a device counter (sf_kv::d_seen) that the graph's last node advances, so the
// same executable graph replays for every step of every stream; inputs/outputs go through
// graph-owned staging buffers (two small D2D copies per step).
__global__ void set_int_kernel(int* p, int v) { *p = v; }
__global__ void add_int_kernel(int* p, int v) {
  griddep_wait();
  *p += v;
}

int g_stream_graph_opt = -1;   // sf_set_option("stream_graph", v): -1 environment default, 0 off, 1 on
bool stream_graphs_enabled() {
  static const bool env_on = [] { const char* e = getenv("SF_STREAM_GRAPH"); return !(e && e[0] == '0'); }();
  return g_stream_graph_opt < 0 ? env_on : g_stream_graph_opt != 0;
}

void destroy_graph(StreamGraph& g) {
  if (g.exec) cudaGraphExecDestroy(g.exec);
  if (g.stage) cudaFree(g.stage);
  g.exec = nullptr; g.stage = nullptr;
}

// returns 1 when the step was served by a graph launch, 0 when the caller must launch directly,
// < 0 on error
int try_stream_graph(sf_ctx* c, cudaStream_t st, sf_kv* kv, const void* pixels, int pix_dtype, int B, int T, int Hh,
                     int Ww, void* last_hidden, void* pooler, void* const* hidden_states, void* ws, size_t ws_bytes) {
  if (!stream_graphs_enabled() || kv->graphs_disabled || prof_enabled() || phase_prof_enabled()) return 0;
  SF_CHECK(check_shape(c, B, T, Hh, Ww));
  const int P = c->cfg.patch_size, D = c->D;
  const int S = (Hh / P) * (Ww / P);
  if (kv->B != B || kv->S != S || kv->seen + T > kv->cap) return 0;   // the direct path reports the error
  StreamGraph* g = nullptr;
  for (auto& it : kv->graphs)
    if (it.B == B && it.T == T && it.H == Hh && it.W == Ww && it.pix_dtype == pix_dtype && it.pooler == (pooler != nullptr) &&
        it.hidden == (hidden_states != nullptr)) g = &it;
  if (!g) {
    if (kv->graphs.size() >= 8) return 0;
    StreamGraph n;
    n.B = B; n.T = T; n.H = Hh; n.W = Ww; n.pix_dtype = pix_dtype; n.pooler = pooler != nullptr;
    n.hidden = hidden_states != nullptr;
    kv->graphs.push_back(n);
    g = &kv->graphs.back();
  }
  if (g->exec && (g->ws != ws || g->ws_bytes != ws_bytes || g->pos_epoch != c->pos_epoch)) {   // workspace or position table was re-allocated: re-capture
    cudaGraphExecDestroy(g->exec);
    g->exec = nullptr;
    g->calls = 1;
  }
  bool just_captured = false;
  if (!g->exec) {
    // first call with this key runs eagerly (lazy one-time initialisation inside the launchers
    // must not happen under capture); the second one captures
    if (g->calls++ == 0) return 0;
    if (!g->stage) {
      const size_t al = 256;
      g->pix_bytes = (static_cast<size_t>(B) * T * c->cfg.num_channels * Hh * Ww * dtype_size(pix_dtype) + al - 1) / al * al;
      g->lh_bytes = (static_cast<size_t>(B) * T * S * D * 2 + al - 1) / al * al;
      g->pool_bytes = (static_cast<size_t>(B) * T * D * 2 + al - 1) / al * al;
      g->hs_bytes = g->hidden ? g->lh_bytes : 0;     // one [B, S*T, D] tensor per layer boundary (the VideoQA tower
                                                     // always asks for them, …timesformer_encoder.py:1536)
      SF_CUDA(cudaMalloc(&g->stage, g->pix_bytes + g->lh_bytes + g->pool_bytes + g->hs_bytes * (c->L + 1)));
      g->hs_ptrs.clear();
      for (int l = 0; g->hidden && l <= c->L; ++l)
        g->hs_ptrs.push_back(g->stage + g->pix_bytes + g->lh_bytes + g->pool_bytes + g->hs_bytes * l);
    }
    cudaGraph_t graph = nullptr;
    const uint64_t l0 = launch_count();
    static const bool debug = getenv("SF_STREAM_GRAPH_DEBUG") != nullptr;
    auto give_up = [&](const char* what, cudaError_t err, int rc) {
      if (debug) fprintf(stderr, "[streamformer_b200] stream graph disabled: %s: %s (rc %d: %s)\n", what,
                         cudaGetErrorString(err), rc, last_error());
      cudaGetLastError();
      kv->graphs_disabled = true;
      return 0;
    };
    cudaStream_t cs = kv->cap_stream;
    cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return give_up("cudaStreamBeginCapture", e, 0);
    int rc = forward_impl(c, cs, kv, g->stage, pix_dtype, B, T, Hh, Ww, g->stage + g->pix_bytes,
                          pooler ? g->stage + g->pix_bytes + g->lh_bytes : nullptr, g->hidden ? g->hs_ptrs.data() : nullptr,
                          nullptr, ws, ws_bytes, true);
    if (rc == 0) {
      LaunchCfg lc(dim3(1), dim3(1), 0, cs);
      if (cudaLaunchKernelEx(&lc.cfg, add_int_kernel, kv->d_seen, T) != cudaSuccess) rc = SF_ERR_CUDA;
      count_launch();
    }
    e = cudaStreamEndCapture(cs, &graph);
    if (rc != 0 || e != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      return give_up("capture", e, rc);
    }
    e = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      g->exec = nullptr;
      return give_up("cudaGraphInstantiate", e, 0);
    }
    g->ws = ws; g->ws_bytes = ws_bytes; g->pos_epoch = c->pos_epoch;
    g->kernels = static_cast<int>(launch_count() - l0);   // counted while capturing; they run in the launch below
    just_captured = true;
  }
  if (kv->seen_dirty) {
    set_int_kernel<<<1, 1, 0, st>>>(kv->d_seen, kv->seen);
    count_launch();
    kv->seen_dirty = false;
  }
  const size_t pix_n = static_cast<size_t>(B) * T * c->cfg.num_channels * Hh * Ww * dtype_size(pix_dtype);
  SF_CUDA(cudaMemcpyAsync(g->stage, pixels, pix_n, cudaMemcpyDeviceToDevice, st));
  SF_CUDA(cudaGraphLaunch(g->exec, st));
  if (!just_captured) count_launch(g->kernels);   // kernels inside the replayed graph
  SF_CUDA(cudaMemcpyAsync(last_hidden, g->stage + g->pix_bytes, static_cast<size_t>(B) * T * S * D * 2,
                          cudaMemcpyDeviceToDevice, st));
  if (pooler)
    SF_CUDA(cudaMemcpyAsync(pooler, g->stage + g->pix_bytes + g->lh_bytes, static_cast<size_t>(B) * T * D * 2,
                            cudaMemcpyDeviceToDevice, st));
  for (int l = 0; g->hidden && l <= c->L; ++l)
    SF_CUDA(cudaMemcpyAsync(hidden_states[l], g->hs_ptrs[l], static_cast<size_t>(B) * T * S * D * 2, cudaMemcpyDeviceToDevice, st));
  kv->seen += T;
  kv->graph_launches += 1;
  return 1;
}

// ------------------------------------------------------------------ weight binding
struct Binder {
  sf_ctx* c;
  cudaStream_t st;
  std::unordered_map<std::string, const sf_weight_desc*> map;
  Bump arena;
  float* scratch[3] = {nullptr, nullptr, nullptr};
  size_t scratch_elems = 0;

  const sf_weight_desc* find(const std::string& name, bool required = true) {
    auto it = map.find(name);
    if (it == map.end()) {
      if (required) set_error("sf_bind_weights: missing tensor '%s'", name.c_str());
      return nullptr;
    }
    return it->second;
  }
  static long numel(const sf_weight_desc* d) {
    long n = 1;
    for (int i = 0; i < d->ndim; ++i) n *= d->shape[i];
    return n;
  }
  int expect(const sf_weight_desc* d, long n) {
    if (numel(d) != n) {
      set_error("sf_bind_weights: tensor '%s' has %ld elements, expected %ld", d->name, numel(d), n);
      return SF_ERR_INVALID;
    }
    return 0;
  }
  // fp32 vector/table straight into the arena
  int vec(const std::string& name, long n, float** out) {
    const sf_weight_desc* d = find(name);
    if (!d) return SF_ERR_INVALID;
    SF_CHECK(expect(d, n));
    *out = static_cast<float*>(arena.take(n * sizeof(float)));
    if (!*out) { set_error("sf_bind_weights: arena exhausted at '%s'", name.c_str()); return SF_ERR_STATE; }
    return cast(st, d->dtype, d->data, kF32, *out, n);
  }
  // matrix in activation dtype
  int mat(const std::string& name, long rows, long cols, void** out) {
    const sf_weight_desc* d = find(name);
    if (!d) return SF_ERR_INVALID;
    SF_CHECK(expect(d, rows * cols));
    *out = arena.take(rows * cols * 2);
    if (!*out) { set_error("sf_bind_weights: arena exhausted at '%s'", name.c_str()); return SF_ERR_STATE; }
    return cast(st, d->dtype, d->data, c->cfg.dtype, *out, rows * cols);
  }
  int to_scratch(int slot, const sf_weight_desc* d, long n) {
    if (static_cast<size_t>(n) > scratch_elems) { set_error("bind scratch too small"); return SF_ERR_INVALID; }
    return cast(st, d->dtype, d->data, kF32, scratch[slot], n);
  }
  // W (+ lora_b . lora_a, …siglip.py:653-654, 749-751) as fp32 in scratch[0]
  int load_f32(const std::string& wname, const std::string& aname, const std::string& bname, long O, long I) {
    const sf_weight_desc* wd = find(wname);
    if (!wd) return SF_ERR_INVALID;
    SF_CHECK(expect(wd, O * I));
    SF_CHECK(to_scratch(0, wd, O * I));
    const sf_weight_desc* a = aname.empty() ? nullptr : find(aname, false);
    const sf_weight_desc* bm = bname.empty() ? nullptr : find(bname, false);
    if (!a || !bm) return 0;
    const long R = numel(a) / I;
    SF_CHECK(expect(a, R * I));
    SF_CHECK(expect(bm, O * R));
    SF_CHECK(to_scratch(1, a, R * I));
    SF_CHECK(to_scratch(2, bm, O * R));
    const long total = O * I;
    lora_merge_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(scratch[0], scratch[1], scratch[2],
                                                                                  static_cast<int>(O), static_cast<int>(I),
                                                                                  static_cast<int>(R));
    count_launch();
    return 0;
  }
  // W (+ LoRA) in activation dtype
  int mat_lora(const std::string& wname, const std::string& aname, const std::string& bname, long O, long I,
               void** out) {
    SF_CHECK(load_f32(wname, aname, bname, O, I));
    *out = arena.take(O * I * 2);
    if (!*out) { set_error("sf_bind_weights: arena exhausted at '%s'", wname.c_str()); return SF_ERR_STATE; }
    return cast(st, kF32, scratch[0], c->cfg.dtype, *out, O * I);
  }
  // Linear preceded by a LayerNorm (gamma, beta already in the arena): gamma-scaled matrix, folded
  // bias and column sums (see ln_fold_kernel)
  int mat_ln(const std::string& wname, const std::string& aname, const std::string& bname, const std::string& biasname,
             const float* gamma, const float* beta, long O, long I, void** w_out, float** bias_out, float** colsum_out) {
    SF_CHECK(load_f32(wname, aname, bname, O, I));
    float* bias_raw = nullptr;   // config.qkv_bias=False: the Linear has no bias (ln_fold_kernel takes nullptr)
    if (find(biasname, false)) SF_CHECK(vec(biasname, O, &bias_raw));
    *w_out = arena.take(O * I * 2);
    *bias_out = static_cast<float*>(arena.take(O * sizeof(float)));
    *colsum_out = static_cast<float*>(arena.take(O * sizeof(float)));
    if (!*w_out || !*bias_out || !*colsum_out) { set_error("sf_bind_weights: arena exhausted at '%s'", wname.c_str()); return SF_ERR_STATE; }
    const unsigned blocks = static_cast<unsigned>((O + 7) / 8);
    if (c->cfg.dtype == kBF16)
      ln_fold_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(scratch[0], gamma, beta, bias_raw, static_cast<__nv_bfloat16*>(*w_out),
                                                            *colsum_out, *bias_out, static_cast<int>(O), static_cast<int>(I));
    else
      ln_fold_kernel<__half><<<blocks, 256, 0, st>>>(scratch[0], gamma, beta, bias_raw, static_cast<__half*>(*w_out),
                                                     *colsum_out, *bias_out, static_cast<int>(O), static_cast<int>(I));
    count_launch();
    return 0;
  }
};

size_t arena_bytes_for(const sf_ctx* c) {
  const size_t D = c->D, I = c->I, L = c->L, K = c->Kp, S0 = c->S0, F = c->cfg.num_frames;
  size_t per_layer = (3 * D * D + D * D * 3 + 3 * D * D + D * D + 2 * D * I) * 2  // matrices (incl. folded)
                     + (3 * D + D * 3 + 3 * D + D + I + D + 6 * D + 1 + 2 * (3 * D + 3 * D + I)) * 4 + 60 * 256;
  size_t other = D * K * 2 + (D + S0 * D + F * D + 2 * D) * 4 + (2 * D * D + D * D + 2 * D * I) * 2 +
                 (2 * D + D + I + D + D + 2 * D) * 4 + 40 * 256;
  return per_layer * L + other + (1 << 20);
}

}  // namespace

// =================================================================================== C ABI
extern "C" {

const char* sf_last_error(void) { return last_error(); }
const char* sf_version(void) { return "streamformer_b200 0.1 (sm_100a)"; }
uint64_t sf_launch_count(void) { return launch_count(); }
int sf_set_option(const char* name, int value) {
  if (name && strcmp(name, "gemm_chain") == 0) { set_gemm_chain(value); return 0; }
  if (name && strcmp(name, "stream_graph") == 0) { g_stream_graph_opt = value; return 0; }
  if (name && strcmp(name, "dual_stream") == 0) { g_dual_stream_opt = value; return 0; }
  if (name && strcmp(name, "spatial_row") == 0) { set_spatial_row(value); return 0; }
  if (name && strcmp(name, "decode_tma") == 0) { set_decode_tma(value); return 0; }
  set_error("sf_set_option: unknown option '%s'", name ? name : "(null)");
  return SF_ERR_INVALID;
}
int sf_profile(int mode) { prof_set_mode(mode); return 0; }
int sf_profile_collect_phases(double* ms, long long* count, int n_phases) { return phase_collect(ms, count, n_phases); }
int sf_profile_collect(double* ms, double* flops, double* bytes, long long* launches, int n_classes) {
  return prof_collect(ms, flops, bytes, launches, n_classes);
}

int sf_create(const sf_config* cfg, int device, sf_ctx** out) {
  if (!cfg || !out) { set_error("sf_create: null argument"); return SF_ERR_INVALID; }
  if (cfg->hidden_size % cfg->num_attention_heads || cfg->hidden_size / cfg->num_attention_heads != 64) {
    set_error("sf_create: head dim must be 64 (hidden %d / heads %d)", cfg->hidden_size, cfg->num_attention_heads);
    return SF_ERR_INVALID;
  }
  if (cfg->dtype != SF_BF16 && cfg->dtype != SF_F16) { set_error("sf_create: dtype must be bf16 or f16"); return SF_ERR_INVALID; }
  if (cfg->hidden_act != SF_ACT_GELU && cfg->hidden_act != SF_ACT_GELU_TANH) {
    set_error("sf_create: hidden_act must be gelu or gelu_pytorch_tanh");
    return SF_ERR_INVALID;
  }
  const int Kp = cfg->num_channels * cfg->patch_size * cfg->patch_size;
  if ((cfg->patch_size % 8) || (cfg->image_size % cfg->patch_size) || (cfg->hidden_size % 8) ||
      (cfg->intermediate_size % 8) || Kp > cfg->intermediate_size) {
    set_error("sf_create: unsupported geometry (patch %d image %d hidden %d mlp %d)", cfg->patch_size,
              cfg->image_size, cfg->hidden_size, cfg->intermediate_size);
    return SF_ERR_INVALID;
  }
  DeviceGuard dg(device);
  int major = 0;
  SF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) {
    set_error("sf_create: device %d is sm_%dx; this library contains sm_100a code only", device, major);
    return SF_ERR_CUDA;
  }
  sf_ctx* c = new sf_ctx();
  c->cfg = *cfg;
  c->device = device;
  c->D = cfg->hidden_size; c->H = cfg->num_attention_heads; c->I = cfg->intermediate_size;
  c->L = cfg->num_hidden_layers; c->Kp = Kp;
  const int g = cfg->image_size / cfg->patch_size;
  c->S0 = g * g;
  c->layers.resize(c->L);
  c->have_layer.assign(c->L, 0);
  *out = c;
  return 0;
}

int sf_destroy(sf_ctx* c) {
  if (!c) return 0;
  DeviceGuard dg(c->device);
  if (c->arena) cudaFree(c->arena);
  if (c->pos_alt) cudaFree(c->pos_alt);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  delete c;
  return 0;
}

int sf_bind_weights(sf_ctx* c, void* stream, const sf_weight_desc* w, int n) {
  if (!c || !w) { set_error("sf_bind_weights: null argument"); return SF_ERR_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard dg(c->device);
  Binder b;
  b.c = c; b.st = st;
  for (int i = 0; i < n; ++i) {
    std::string name = w[i].name ? w[i].name : "";
    if (name.rfind("timesformer.", 0) == 0) name = name.substr(12);
    b.map[name] = &w[i];
  }
  if (!c->arena) {
    c->arena_bytes = arena_bytes_for(c);
    SF_CUDA(cudaMalloc(&c->arena, c->arena_bytes));
  }
  b.arena.base = c->arena; b.arena.size = c->arena_bytes;
  const long D = c->D, I = c->I, K = c->Kp;
  b.scratch_elems = static_cast<size_t>(3 * D > I ? 3 * D : I) * D;
  if (b.scratch_elems < static_cast<size_t>(D * K)) b.scratch_elems = D * K;
  float* scratch_mem = nullptr;
  SF_CUDA(cudaMalloc(&scratch_mem, b.scratch_elems * 3 * sizeof(float)));
  for (int i = 0; i < 3; ++i) b.scratch[i] = scratch_mem + i * b.scratch_elems;

  // Parameter groups are bound independently: the full model brings all of them; a stand-alone
  // TimesformerEmbeddingsSigLIP / TimesformerEncoder / TimesformerLayerSigLIP / pooling head composed
  // inside someone else's PreTrainedModel (downstream/AR/models/modeling_timesformer_video_classification.py:42-56,
  // models/modeling_timesformer_siglip_adapter.py:481-482) brings only its own.
  c->bound = false;
  c->have_embed = b.find("embeddings.patch_embeddings.projection.weight", false) != nullptr;
  c->have_post = b.find("post_layernorm.weight", false) != nullptr;
  c->have_head = b.find("head.probe", false) != nullptr;
  bool any = c->have_embed || c->have_post || c->have_head;
  for (int l = 0; l < c->L; ++l) {
    c->have_layer[l] = b.find("encoder.layer." + std::to_string(l) + ".attention.attention.qkv.weight", false) != nullptr;
    any = any || c->have_layer[l];
  }
  if (!any) { set_error("sf_bind_weights: none of the %d tensors belongs to the encoder (names as in the reference state dict)", n); cudaFree(scratch_mem); return SF_ERR_INVALID; }

  auto body = [&]() -> int {
    if (c->have_embed) {
    SF_CHECK(b.mat("embeddings.patch_embeddings.projection.weight", D, K, &c->patch_w));
    SF_CHECK(b.vec("embeddings.patch_embeddings.projection.bias", D, &c->patch_b));
    SF_CHECK(b.vec("embeddings.position_embeddings", static_cast<long>(c->S0) * D, &c->pos));
    SF_CHECK(b.vec("embeddings.time_embeddings", static_cast<long>(c->cfg.num_frames) * D, &c->time_emb));
    }
    for (int l = 0; l < c->L; ++l) {
      if (!c->have_layer[l]) continue;
      LayerW& lw = c->layers[l];
      const std::string p = "encoder.layer." + std::to_string(l) + ".";
      SF_CHECK(b.vec(p + "temporal_attention_gating", 1, &lw.gate));
      SF_CHECK(b.vec(p + "temporal_layernorm.weight", D, &lw.ln_t_g));
      SF_CHECK(b.vec(p + "temporal_layernorm.bias", D, &lw.ln_t_b));
      SF_CHECK(b.vec(p + "layernorm_before.weight", D, &lw.ln_s_g));
      SF_CHECK(b.vec(p + "layernorm_before.bias", D, &lw.ln_s_b));
      SF_CHECK(b.vec(p + "layernorm_after.weight", D, &lw.ln_a_g));
      SF_CHECK(b.vec(p + "layernorm_after.bias", D, &lw.ln_a_b));
      SF_CHECK(b.mat_ln(p + "temporal_attention.attention.qkv.weight", "", "", p + "temporal_attention.attention.qkv.bias",
                        lw.ln_t_g, lw.ln_t_b, 3 * D, D, &lw.t_qkv_w, &lw.t_qkv_b, &lw.t_qkv_cs));
      SF_CHECK(b.mat(p + "temporal_attention.output.dense.weight", D, D, &lw.t_out_w));
      SF_CHECK(b.vec(p + "temporal_attention.output.dense.bias", D, &lw.t_out_b));
      SF_CHECK(b.mat(p + "temporal_dense.weight", D, D, &lw.t_dense_w));
      SF_CHECK(b.vec(p + "temporal_dense.bias", D, &lw.t_dense_b));
      {
        // folded projection: W_f = W_td . W_o ; b_f = W_td . b_o + b_td   (fp32, then one rounding)
        const sf_weight_desc* wo = b.find(p + "temporal_attention.output.dense.weight");
        const sf_weight_desc* wtd = b.find(p + "temporal_dense.weight");
        SF_CHECK(b.to_scratch(0, wtd, D * D));
        SF_CHECK(b.to_scratch(1, wo, D * D));
        dim3 grid(static_cast<unsigned>((D + 15) / 16), static_cast<unsigned>((D + 15) / 16)), blk(16, 16);
        matmul_nn_kernel<<<grid, blk, 0, st>>>(b.scratch[2], b.scratch[0], b.scratch[1], static_cast<int>(D),
                                               static_cast<int>(D), static_cast<int>(D));
        count_launch();
        lw.t_fold_w = b.arena.take(D * D * 2);
        SF_CHECK(cast(st, kF32, b.scratch[2], c->cfg.dtype, lw.t_fold_w, D * D));
        lw.t_fold_b = static_cast<float*>(b.arena.take(D * sizeof(float)));
        matvec_kernel<<<static_cast<unsigned>((D + 7) / 8), 256, 0, st>>>(lw.t_fold_b, b.scratch[0], lw.t_out_b,
                                                                          lw.t_dense_b, static_cast<int>(D),
                                                                          static_cast<int>(D), 1.0f);
        count_launch();
      }
      SF_CHECK(b.mat_ln(p + "attention.attention.qkv.weight", p + "attention.attention.qkv_lora_a.weight",
                        p + "attention.attention.qkv_lora_b.weight", p + "attention.attention.qkv.bias", lw.ln_s_g,
                        lw.ln_s_b, 3 * D, D, &lw.s_qkv_w, &lw.s_qkv_b, &lw.s_qkv_cs));
      SF_CHECK(b.mat_lora(p + "attention.output.dense.weight", p + "attention.output.dense_lora_a.weight",
                          p + "attention.output.dense_lora_b.weight", D, D, &lw.s_out_w));
      SF_CHECK(b.vec(p + "attention.output.dense.bias", D, &lw.s_out_b));
      SF_CHECK(b.mat_ln(p + "intermediate.dense.weight", "", "", p + "intermediate.dense.bias", lw.ln_a_g, lw.ln_a_b, I,
                        D, &lw.fc1_w, &lw.fc1_b, &lw.fc1_cs));
      SF_CHECK(b.mat(p + "output.dense.weight", D, I, &lw.fc2_w));
      SF_CHECK(b.vec(p + "output.dense.bias", D, &lw.fc2_b));
    }
    if (c->have_post) {
    SF_CHECK(b.vec("post_layernorm.weight", D, &c->post_g));
    SF_CHECK(b.vec("post_layernorm.bias", D, &c->post_b));
    }
    // pooling head: in_proj rows [0,D) = W_q, [D,3D) = W_k;W_v (…siglip.py:1135-1137)
    if (c->have_head) {
      const sf_weight_desc* ipw = b.find("head.attention.in_proj_weight");
      const sf_weight_desc* ipb = b.find("head.attention.in_proj_bias");
      const sf_weight_desc* probe = b.find("head.probe");
      if (!ipw || !ipb || !probe) return SF_ERR_INVALID;
      SF_CHECK(b.expect(ipw, 3 * D * D));
      SF_CHECK(b.expect(ipb, 3 * D));
      SF_CHECK(b.expect(probe, D));
      const size_t es_in = dtype_size(ipw->dtype);
      c->head_kv_w = b.arena.take(2 * D * D * 2);
      SF_CHECK(cast(st, ipw->dtype, static_cast<const uint8_t*>(ipw->data) + static_cast<size_t>(D) * D * es_in,
                    c->cfg.dtype, c->head_kv_w, 2 * D * D));
      c->head_kv_b = static_cast<float*>(b.arena.take(2 * D * sizeof(float)));
      SF_CHECK(cast(st, ipb->dtype, static_cast<const uint8_t*>(ipb->data) + static_cast<size_t>(D) * dtype_size(ipb->dtype),
                    kF32, c->head_kv_b, 2 * D));
      // q = (W_q probe + b_q) / sqrt(64), constant per model
      SF_CHECK(b.to_scratch(0, ipw, D * D));   // first D rows only are used
      SF_CHECK(cast(st, ipb->dtype, ipb->data, kF32, b.scratch[1], D));
      SF_CHECK(cast(st, probe->dtype, probe->data, kF32, b.scratch[2], D));
      c->head_q = static_cast<float*>(b.arena.take(D * sizeof(float)));
      matvec_kernel<<<static_cast<unsigned>((D + 7) / 8), 256, 0, st>>>(c->head_q, b.scratch[0], b.scratch[2],
                                                                        b.scratch[1], static_cast<int>(D),
                                                                        static_cast<int>(D), 0.125f);
      count_launch();
      // u_h = W_k,h^T q_h (fp32): the collapsed pooling attention scores tokens directly
      SF_CHECK(cast(st, ipw->dtype, static_cast<const uint8_t*>(ipw->data) + static_cast<size_t>(D) * D * es_in, kF32,
                    b.scratch[0], D * D));
      c->head_u = static_cast<float*>(b.arena.take(static_cast<size_t>(c->H) * D * sizeof(float)));
      if (!c->head_u) { set_error("sf_bind_weights: arena exhausted at head_u"); return SF_ERR_STATE; }
      head_u_kernel<<<static_cast<unsigned>((c->H * D + 255) / 256), 256, 0, st>>>(c->head_u, b.scratch[0], c->head_q, c->H,
                                                                                   static_cast<int>(D));
      count_launch();
    }
    if (c->have_head) {
    SF_CHECK(b.mat("head.attention.out_proj.weight", D, D, &c->head_out_w));
    SF_CHECK(b.vec("head.attention.out_proj.bias", D, &c->head_out_b));
    SF_CHECK(b.vec("head.layernorm.weight", D, &c->head_ln_g));
    SF_CHECK(b.vec("head.layernorm.bias", D, &c->head_ln_b));
    SF_CHECK(b.mat("head.mlp.fc1.weight", I, D, &c->head_fc1_w));
    SF_CHECK(b.vec("head.mlp.fc1.bias", I, &c->head_fc1_b));
    SF_CHECK(b.mat("head.mlp.fc2.weight", D, I, &c->head_fc2_w));
    SF_CHECK(b.vec("head.mlp.fc2.bias", D, &c->head_fc2_b));
    }
    if (b.arena.overflow) { set_error("sf_bind_weights: internal arena too small (%zu > %zu)", b.arena.off, b.arena.size); return SF_ERR_STATE; }
    return 0;
  };
  int rc = body();
  cudaError_t e = cudaStreamSynchronize(st);  // scratch is freed below; binding is not a hot path
  cudaFree(scratch_mem);
  if (rc || e != cudaSuccess) {
    c->have_embed = c->have_post = c->have_head = false;
    c->have_layer.assign(c->L, 0);
  }
  if (rc) return rc;
  if (e != cudaSuccess) { set_error("sf_bind_weights: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
  c->bound = true;
  return 0;
}

int sf_set_pos_embed(sf_ctx* c, void* stream, const float* pos, int S) {
  if (!c || !pos || S <= 0) { set_error("sf_set_pos_embed: bad argument"); return SF_ERR_INVALID; }
  DeviceGuard dg(c->device);
  if (c->pos_alt_S != S) {
    // captured streaming graphs hold the old table's address: the epoch makes them re-capture; the old
    // table may still be read by work in flight on the stream, so it is released only after a sync
    if (c->pos_alt) { cudaStreamSynchronize(static_cast<cudaStream_t>(stream)); cudaFree(c->pos_alt); c->pos_alt = nullptr; }
    SF_CUDA(cudaMalloc(&c->pos_alt, static_cast<size_t>(S) * c->D * sizeof(float)));
    c->pos_alt_S = S;
    ++c->pos_epoch;
  }
  SF_CUDA(cudaMemcpyAsync(c->pos_alt, pos, static_cast<size_t>(S) * c->D * sizeof(float), cudaMemcpyDeviceToDevice,
                          static_cast<cudaStream_t>(stream)));
  return 0;
}

int sf_workspace_bytes(const sf_ctx* c, int B, int T, int Hh, int Ww, size_t* out) {
  if (!c || !out) { set_error("sf_workspace_bytes: null argument"); return SF_ERR_INVALID; }
  const int P = c->cfg.patch_size;
  const long S = static_cast<long>(Hh / P) * (Ww / P);
  const long M = static_cast<long>(B) * T * S;
  size_t layer = static_cast<size_t>(M) * c->D * 2 + 256 + layer_ws_bytes(c, M);
  size_t head = static_cast<size_t>(M) * c->D * 2 + 256 + head_ws_bytes(c, static_cast<long>(B) * T, S);
  *out = (layer > head ? layer : head) + 4096 + 65536;   // slack for the two-half carve-up of the dual-stream forward
  return 0;
}

int sf_forward(sf_ctx* c, void* stream, const void* pixels, int pixels_dtype, int B, int T, int Hh, int Ww,
               void* last_hidden, void* pooler, void* const* hidden_states, void* const* attentions,
               void* workspace, size_t workspace_bytes) {
  if (!c || !pixels || !last_hidden) { set_error("sf_forward: null argument"); return SF_ERR_INVALID; }
  DeviceGuard dg(c->device);
  return forward_impl(c, static_cast<cudaStream_t>(stream), nullptr, pixels, pixels_dtype, B, T, Hh, Ww, last_hidden,
                      pooler, hidden_states, attentions, workspace, workspace_bytes);
}

int sf_kv_create(sf_ctx* c, int B, int S, int max_frames, int time_horizon, sf_kv** out) {
  if (!c || !out || B <= 0 || S <= 0 || max_frames <= 0) { set_error("sf_kv_create: bad argument"); return SF_ERR_INVALID; }
  DeviceGuard dg(c->device);
  sf_kv* kv = new sf_kv();
  kv->ctx = c; kv->B = B; kv->S = S; kv->cap = max_frames; kv->horizon = time_horizon;
  kv->kv_bytes = static_cast<size_t>(B) * S * c->H * max_frames * 64 * 2;
  kv->layer_stride = 2 * kv->kv_bytes;
  cudaError_t e = cudaMalloc(&kv->mem, kv->layer_stride * c->L);
  // rows past the stream position are read as part of 16-row slabs (their probabilities are exactly 0): they must be finite
  if (e == cudaSuccess) e = cudaMemset(kv->mem, 0, kv->layer_stride * c->L);
  if (e == cudaSuccess) e = cudaMalloc(&kv->d_seen, sizeof(int));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&kv->cap_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("sf_kv_create: cudaMalloc(%zu bytes) failed: %s", kv->layer_stride * c->L, cudaGetErrorString(e));
    if (kv->mem) cudaFree(kv->mem);
    delete kv;
    return SF_ERR_CUDA;
  }
  *out = kv;
  return 0;
}
int sf_kv_reset(sf_kv* kv) {
  if (kv) { kv->seen = 0; kv->seen_dirty = true; }
  return 0;
}
int sf_kv_destroy(sf_kv* kv) {
  if (!kv) return 0;
  DeviceGuard dg(kv->ctx ? kv->ctx->device : 0);
  for (auto& g : kv->graphs) destroy_graph(g);
  if (kv->mem) cudaFree(kv->mem);
  if (kv->d_seen) cudaFree(kv->d_seen);
  if (kv->cap_stream) cudaStreamDestroy(kv->cap_stream);
  delete kv;
  return 0;
}
int sf_kv_advance(sf_kv* kv, int frames) {
  if (!kv || frames < 0 || kv->seen + frames > kv->cap) { set_error("sf_kv_advance: bad argument or cache overflow"); return SF_ERR_INVALID; }
  kv->seen += frames;
  kv->seen_dirty = true;
  return 0;
}
int sf_set_pixel_norm(sf_ctx* c, const float* mean, const float* std, int n) {
  if (!c || !mean || !std || n <= 0 || n > 4) { set_error("sf_set_pixel_norm: bad argument"); return SF_ERR_INVALID; }
  for (int i = 0; i < 4; ++i) {
    c->pix_mean[i] = mean[i < n ? i : n - 1];
    c->pix_std[i] = std[i < n ? i : n - 1];
    if (!(c->pix_std[i] > 0.f)) { set_error("sf_set_pixel_norm: std must be positive"); return SF_ERR_INVALID; }
  }
  ++c->pos_epoch;   // captured streaming graphs baked the old constants in: re-capture
  return 0;
}
int sf_kv_seq_len(const sf_kv* kv) { return kv ? kv->seen : 0; }
int sf_kv_capacity(const sf_kv* kv) { return kv ? kv->cap : 0; }
long long sf_kv_graph_launches(const sf_kv* kv) { return kv ? kv->graph_launches : 0; }

int sf_forward_stream(sf_ctx* c, void* stream, sf_kv* kv, const void* pixels, int pixels_dtype, int B, int T_new,
                      int Hh, int Ww, void* last_hidden, void* pooler, void* const* hidden_states, void* workspace,
                      size_t workspace_bytes) {
  if (!c || !kv || !pixels || !last_hidden) { set_error("sf_forward_stream: null argument"); return SF_ERR_INVALID; }
  if (kv->ctx != c) { set_error("sf_forward_stream: the cache belongs to another context"); return SF_ERR_INVALID; }
  DeviceGuard dg(c->device);
  {
    const int served = try_stream_graph(c, static_cast<cudaStream_t>(stream), kv, pixels, pixels_dtype, B, T_new, Hh, Ww,
                                        last_hidden, pooler, hidden_states, workspace, workspace_bytes);
    if (served != 0) return served < 0 ? served : 0;
  }
  return forward_impl(c, static_cast<cudaStream_t>(stream), kv, pixels, pixels_dtype, B, T_new, Hh, Ww, last_hidden,
                      pooler, hidden_states, nullptr, workspace, workspace_bytes);
}

int sf_embed_forward(sf_ctx* c, void* stream, const void* pixels, int pixels_dtype, int B, int T, int Hh, int Ww,
                     int time_off, int time_total, void* x_out, void* workspace, size_t workspace_bytes) {
  if (!c || !pixels || !x_out) { set_error("sf_embed_forward: null argument"); return SF_ERR_INVALID; }
  DeviceGuard dg(c->device);
  SF_CHECK(check_shape(c, B, T, Hh, Ww));
  const int P = c->cfg.patch_size;
  const long M = static_cast<long>(B) * T * (Hh / P) * (Ww / P);
  Bump b;
  b.base = static_cast<uint8_t*>(workspace); b.size = workspace_bytes;
  void* patches = b.take(static_cast<size_t>(M) * c->Kp * 2);
  if (b.overflow) { set_error("sf_embed_forward: workspace too small (need %zu)", b.off); return SF_ERR_WORKSPACE; }
  if (time_total < time_off + T) time_total = time_off + T;
  return run_embed(c, static_cast<cudaStream_t>(stream), pixels, pixels_dtype, B, T, Hh, Ww, time_off, time_total, x_out,
                   patches, nullptr);
}

int sf_layer_forward(sf_ctx* c, void* stream, int layer, const void* x_in, void* x_out, int B, int T, int S, sf_kv* kv,
                     float* attn_probs, void* workspace, size_t workspace_bytes) {
  if (!c || !x_in || !x_out) { set_error("sf_layer_forward: null argument"); return SF_ERR_INVALID; }
  if (!c->bound) { set_error("weights not bound"); return SF_ERR_STATE; }
  if (layer < 0 || layer >= c->L) { set_error("sf_layer_forward: layer %d out of range", layer); return SF_ERR_INVALID; }
  if (B <= 0 || T <= 0 || S <= 0) { set_error("sf_layer_forward: bad shape B=%d T=%d S=%d", B, T, S); return SF_ERR_INVALID; }
  DeviceGuard dg(c->device);
  SF_CHECK(check_groups(c, false, layer, layer + 1, false, false));
  if (kv) {
    // block-level streaming: every layer of a step appends at the same position kv->seen; the caller
    // advances the cache ONCE after the last layer with sf_kv_advance (the twin's DynamicCache grows per
    // layer instead, …timesformer_encoder.py:517-518)
    if (kv->ctx != c) { set_error("sf_layer_forward: the cache belongs to another context"); return SF_ERR_INVALID; }
    if (kv->B != B || kv->S != S) { set_error("KV cache was created for B=%d S=%d, got B=%d S=%d", kv->B, kv->S, B, S); return SF_ERR_INVALID; }
    if (kv->seen + T > kv->cap) { set_error("KV cache overflow: %d cached + %d new frames > capacity %d", kv->seen, T, kv->cap); return SF_ERR_STATE; }
  }
  Bump b;
  b.base = static_cast<uint8_t*>(workspace); b.size = workspace_bytes;
  WsPlan w;
  const long M = static_cast<long>(B) * T * S;
  SF_CHECK(carve_layer_ws(c, M, b, w));
  // x_in comes from the caller: one pass for its row statistics (inside sf_forward they are a
  // by-product of the GEMM that wrote the rows)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SF_CHECK(rowstats(st, c->cfg.dtype, x_in, c->D, static_cast<int>(M), c->D, w.stats[2]));
  return run_layer(c, st, layer, x_in, x_out, B, T, S, kv, attn_probs, w, w.stats[2], 1, w.stats[2]);
}

int sf_encoder_forward(sf_ctx* c, void* stream, const void* x_in, int B, int T, int S, sf_kv* kv, void* x_out,
                       void* const* hidden_states, void* const* attentions, void* workspace, size_t workspace_bytes) {
  if (!c || !x_in || (!x_out && !hidden_states)) { set_error("sf_encoder_forward: null argument"); return SF_ERR_INVALID; }
  if (!c->bound) { set_error("weights not bound"); return SF_ERR_STATE; }
  if (B <= 0 || T <= 0 || S <= 0) { set_error("sf_encoder_forward: bad shape B=%d T=%d S=%d", B, T, S); return SF_ERR_INVALID; }
  DeviceGuard dg(c->device);
  SF_CHECK(check_groups(c, false, 0, c->L, false, false));
  if (kv) {
    if (kv->ctx != c) { set_error("sf_encoder_forward: the cache belongs to another context"); return SF_ERR_INVALID; }
    if (kv->B != B || kv->S != S) { set_error("KV cache was created for B=%d S=%d, got B=%d S=%d", kv->B, kv->S, B, S); return SF_ERR_INVALID; }
    if (kv->seen + T > kv->cap) { set_error("KV cache overflow: %d cached + %d new frames > capacity %d", kv->seen, T, kv->cap); return SF_ERR_STATE; }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Bump b;
  b.base = static_cast<uint8_t*>(workspace); b.size = workspace_bytes;
  WsPlan w;
  const long M = static_cast<long>(B) * T * S;
  void* x = hidden_states ? nullptr : b.take(static_cast<size_t>(M) * c->D * 2);
  SF_CHECK(carve_layer_ws(c, M, b, w));
  // the rows come from the caller: one pass for their LayerNorm statistics, after that the statistics are
  // by-products of the GEMM epilogues exactly as inside sf_forward
  SF_CHECK(rowstats(st, c->cfg.dtype, x_in, c->D, static_cast<int>(M), c->D, w.stats[2]));
  int parts = 1;
  bool qkv_ready = false;
  const void* cur = x_in;
  for (int l = 0; l < c->L; ++l) {
    void* nxt = hidden_states ? hidden_states[l + 1] : (l + 1 == c->L ? x_out : x);
    const LayerW* next = (l + 1 < c->L && !phase_prof_enabled()) ? &c->layers[l + 1] : nullptr;
    bool next_done = false;
    SF_CHECK(run_layer(c, st, l, cur, nxt, B, T, S, kv, attentions ? static_cast<float*>(attentions[l]) : nullptr, w,
                       w.stats[2], parts, w.stats[2], nullptr, qkv_ready, next, &next_done, &parts));
    qkv_ready = next_done;
    cur = nxt;
  }
  if (kv) { kv->seen += T; kv->seen_dirty = true; }
  return 0;
}

int sf_final_norm(sf_ctx* c, void* stream, const void* x, int B, int T, int S, void* last_hidden) {
  if (!c || !x || !last_hidden) { set_error("sf_final_norm: null argument"); return SF_ERR_INVALID; }
  if (!c->bound) { set_error("weights not bound"); return SF_ERR_STATE; }
  DeviceGuard dg(c->device);
  SF_CHECK(check_groups(c, false, 0, 0, true, false));
  return layernorm(static_cast<cudaStream_t>(stream), c->cfg.dtype, x, c->D, c->post_g, c->post_b, c->cfg.layer_norm_eps,
                   last_hidden, c->D, static_cast<long>(B) * T * S, c->D, T > 1 ? kRowBNTtoBTN : kRowIdentity, T, S);
}

int sf_head_forward(sf_ctx* c, void* stream, const void* tokens, int frames, int S, void* pooled, void* workspace,
                    size_t workspace_bytes) {
  if (!c || !tokens || !pooled) { set_error("sf_head_forward: null argument"); return SF_ERR_INVALID; }
  if (!c->bound) { set_error("weights not bound"); return SF_ERR_STATE; }
  DeviceGuard dg(c->device);
  Bump b;
  b.base = static_cast<uint8_t*>(workspace); b.size = workspace_bytes;
  return run_head(c, static_cast<cudaStream_t>(stream), tokens, frames, S, pooled, b);
}

// ---- single kernels
int sf_op_gemm(void* stream, int dtype, const void* A, int lda, const void* W, int ldw, void* out, int ldo, int M, int N,
               int K, const sf_gemm_epilogue* e) {
  GemmEpilogue g;
  if (e) {
    g.bias = e->bias; g.act = e->act; g.residual = e->residual; g.ldr = e->ldr; g.gate = e->gate;
    g.row_map = e->row_map; g.T = e->T > 0 ? e->T : 1; g.S = e->S > 0 ? e->S : 1;
    g.pos = e->pos; g.time_emb = e->time_emb; g.time_len = e->time_len; g.time_total = e->time_total;
    g.time_off = e->time_off;
    g.ln_stats = static_cast<const float2*>(e->ln_stats); g.ln_parts = e->ln_parts; g.ln_colsum = e->ln_colsum;
    g.ln_eps = e->ln_eps; g.stats_out = static_cast<float2*>(e->stats_out);
  }
  return gemm(static_cast<cudaStream_t>(stream), dtype, A, lda, W, ldw, out, ldo, M, N, K, g);
}
int sf_op_layernorm(void* stream, int dtype, const void* x, int ldx, const float* gamma, const float* beta, float eps,
                    void* y, int ldy, int M, int D, int row_map, int T, int S) {
  return layernorm(static_cast<cudaStream_t>(stream), dtype, x, ldx, gamma, beta, eps, y, ldy, M, D, row_map,
                   T > 0 ? T : 1, S > 0 ? S : 1);
}
int sf_op_im2col(void* stream, int pix_dtype, const void* pixels, int act_dtype, void* out, int BT, int C, int Hh, int Ww,
                 int P) {
  static const float half[4] = {0.5f, 0.5f, 0.5f, 0.5f};   // uint8 pixels: Normalize(0.5, 0.5)
  return im2col_patches(static_cast<cudaStream_t>(stream), pix_dtype, pixels, act_dtype, out, BT, C, Hh, Ww, P, half, half);
}
int sf_op_temporal_attention(void* stream, int dtype, const void* qkv, int ld_qkv, const void* kcache, const void* vcache,
                             int Tcap, void* out, int ld_out, int sites, int heads, int Tq, int Tk, int q_off, int causal,
                             float scale) {
  return temporal_attention(static_cast<cudaStream_t>(stream), dtype, qkv, ld_qkv, kcache, vcache, Tcap, out, ld_out, sites,
                            heads, Tq, Tk, q_off, causal, scale);
}
int sf_op_temporal_decode(void* stream, int dtype, const void* qkv, int ld_qkv, void* kcache, void* vcache, int Tcap,
                          void* out, int ld_out, int sites, int heads, int seen, float scale) {
  return temporal_decode(static_cast<cudaStream_t>(stream), dtype, qkv, ld_qkv, kcache, vcache, Tcap, out, ld_out, sites,
                         heads, seen, scale);
}
int sf_op_siglip_head(void* stream, int dtype, const void* image, int ld_i, const void* text, int ld_t, int B, int L, int D,
                      const float* logit_scale, const float* logit_bias, int normalize_image, int normalize_text,
                      const int64_t* targets, int diag_offset, float loss_div, float* logits, int ld_logits, float* loss,
                      void* dlogits, int ld_dlogits, float* dparams) {
  return siglip_head(static_cast<cudaStream_t>(stream), dtype, image, ld_i, text, ld_t, B, L, D, logit_scale, logit_bias,
                     normalize_image, normalize_text, reinterpret_cast<const long long*>(targets), diag_offset, loss_div, logits,
                     ld_logits, loss, dlogits, ld_dlogits, dparams);
}
int sf_op_l2norm_backward(void* stream, int dtype, const void* x, int ldx, const void* dxhat, int ldg, const float* gscale,
                          void* dx, int ldo, int B, int D) {
  return l2norm_backward(static_cast<cudaStream_t>(stream), dtype, x, ldx, dxhat, ldg, gscale, dx, ldo, B, D);
}
int sf_export_packed(sf_ctx* c, void* stream, int layer, const char* name, int transpose, void* dst, size_t dst_bytes) {
  if (!c || !name || !dst) { set_error("sf_export_packed: null argument"); return SF_ERR_INVALID; }
  if (!c->bound) { set_error("weights not bound"); return SF_ERR_STATE; }
  DeviceGuard dg(c->device);
  const long D = c->D, I = c->I, K = c->Kp;
  const void* mat = nullptr; const float* vec = nullptr;
  long rows = 0, cols = 0, n = 0;
  const std::string nm(name);
  if (layer >= 0) {
    if (layer >= c->L || !c->have_layer[layer]) { set_error("sf_export_packed: layer %d is not bound", layer); return SF_ERR_STATE; }
    const LayerW& lw = c->layers[layer];
    if (nm == "t_qkv") { mat = lw.t_qkv_w; rows = 3 * D; cols = D; }
    else if (nm == "t_out") { mat = lw.t_out_w; rows = D; cols = D; }
    else if (nm == "t_dense") { mat = lw.t_dense_w; rows = D; cols = D; }
    else if (nm == "s_qkv") { mat = lw.s_qkv_w; rows = 3 * D; cols = D; }
    else if (nm == "s_out") { mat = lw.s_out_w; rows = D; cols = D; }
    else if (nm == "fc1") { mat = lw.fc1_w; rows = I; cols = D; }
    else if (nm == "fc2") { mat = lw.fc2_w; rows = D; cols = I; }
    else if (nm == "t_qkv_b") { vec = lw.t_qkv_b; n = 3 * D; }
    else if (nm == "t_out_b") { vec = lw.t_out_b; n = D; }
    else if (nm == "t_dense_b") { vec = lw.t_dense_b; n = D; }
    else if (nm == "s_qkv_b") { vec = lw.s_qkv_b; n = 3 * D; }
    else if (nm == "s_out_b") { vec = lw.s_out_b; n = D; }
    else if (nm == "fc1_b") { vec = lw.fc1_b; n = I; }
    else if (nm == "fc2_b") { vec = lw.fc2_b; n = D; }
  } else {
    if (nm.rfind("head", 0) == 0 && !c->have_head) { set_error("sf_export_packed: the pooling head is not bound"); return SF_ERR_STATE; }
    if (nm.rfind("patch", 0) == 0 && !c->have_embed) { set_error("sf_export_packed: the embeddings are not bound"); return SF_ERR_STATE; }
    if (nm == "head_kv") { mat = c->head_kv_w; rows = 2 * D; cols = D; }
    else if (nm == "head_out") { mat = c->head_out_w; rows = D; cols = D; }
    else if (nm == "head_fc1") { mat = c->head_fc1_w; rows = I; cols = D; }
    else if (nm == "head_fc2") { mat = c->head_fc2_w; rows = D; cols = I; }
    else if (nm == "patch") { mat = c->patch_w; rows = D; cols = K; }
    else if (nm == "head_kv_b") { vec = c->head_kv_b; n = 2 * D; }
    else if (nm == "head_out_b") { vec = c->head_out_b; n = D; }
    else if (nm == "head_fc1_b") { vec = c->head_fc1_b; n = I; }
    else if (nm == "head_fc2_b") { vec = c->head_fc2_b; n = D; }
    else if (nm == "patch_b") { vec = c->patch_b; n = D; }
    else if (nm == "head_q") { vec = c->head_q; n = D; }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec) {
    if (dst_bytes < static_cast<size_t>(n) * 4) { set_error("sf_export_packed: destination too small for '%s'", name); return SF_ERR_INVALID; }
    SF_CUDA(cudaMemcpyAsync(dst, vec, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  if (!mat) { set_error("sf_export_packed: unknown tensor '%s' (layer %d)", name, layer); return SF_ERR_INVALID; }
  if (dst_bytes < static_cast<size_t>(rows) * cols * 2) { set_error("sf_export_packed: destination too small for '%s'", name); return SF_ERR_INVALID; }
  if (!transpose) {
    SF_CUDA(cudaMemcpyAsync(dst, mat, static_cast<size_t>(rows) * cols * 2, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  return transpose2d(st, c->cfg.dtype, mat, static_cast<int>(cols), dst, static_cast<int>(rows), static_cast<int>(rows), static_cast<int>(cols));
}
int sf_op_transpose(void* stream, int dtype, const void* in, int ld_in, void* out, int ld_out, int M, int N) {
  return transpose2d(static_cast<cudaStream_t>(stream), dtype, in, ld_in, out, ld_out, M, N);
}
int sf_op_colsum(void* stream, int dtype, const void* x, int ld, int M, int N, float* out) {
  return colsum(static_cast<cudaStream_t>(stream), dtype, x, ld, M, N, out);
}
int sf_op_ln_backward(void* stream, int dtype, const void* x, int ldx, const void* dn, int ld_dn, float eps, const void* dres,
                      int ld_dres, void* dx, int ld_dx, int M, int D) {
  return ln_backward(static_cast<cudaStream_t>(stream), dtype, x, ldx, dn, ld_dn, eps, dres, ld_dres, dx, ld_dx, M, D);
}
int sf_op_ln_affine_backward(void* stream, int dtype, const void* x, int ldx, const void* dy, int ld_dy, const float* gamma, float eps,
                             void* dx, int ld_dx, int M, int D, int row_map, int T, int S, float* dgamma, float* dbeta) {
  return ln_affine_backward(static_cast<cudaStream_t>(stream), dtype, x, ldx, dy, ld_dy, gamma, eps, dx, ld_dx, M, D, row_map, T, S, dgamma, dbeta);
}
int sf_op_gelu(void* stream, int dtype, const void* a, void* h, long long n, int act) {
  return gelu_forward(static_cast<cudaStream_t>(stream), dtype, a, h, static_cast<long>(n), act);
}
int sf_op_gelu_backward(void* stream, int dtype, void* a_h, void* dh_dpre, long long n, int act) {
  return gelu_backward(static_cast<cudaStream_t>(stream), dtype, a_h, dh_dpre, static_cast<long>(n), act);
}
int sf_op_gate_backward(void* stream, int dtype, const void* dx, const void* y, const float* gate, void* dy, long long n, float* dgate) {
  return gate_backward(static_cast<cudaStream_t>(stream), dtype, dx, y, gate, dy, static_cast<long>(n), dgate);
}
int sf_op_wfold_finish(void* stream, int dtype, const void* G, int ldg, const void* Wp, int ldw, const float* gamma, const float* beta,
                       const float* db, void* dW, int out_dtype, int ld_dw, int O, int I, float* dgamma, float* dbeta) {
  return wfold_finish(static_cast<cudaStream_t>(stream), dtype, G, ldg, Wp, ldw, gamma, beta, db, dW, out_dtype, ld_dw, O, I, dgamma, dbeta);
}
int sf_op_wgrad_splits(int M, int O, int I) { return wgrad_splits(M, O, I); }
int sf_op_wgrad(void* stream, int dtype, const void* dY, int ldy, const void* X, int ldx, int M, int O, int I, float* partials) {
  return wgrad(static_cast<cudaStream_t>(stream), dtype, dY, ldy, X, ldx, M, O, I, partials);
}
int sf_op_wfold_finish_partials(void* stream, int dtype, const float* partials, int splits, const void* Wp, int ldw, const float* gamma,
                                const float* beta, const float* db, void* dW, int out_dtype, int ld_dw, int O, int I, float* dgamma,
                                float* dbeta) {
  return wfold_finish(static_cast<cudaStream_t>(stream), dtype, nullptr, I, Wp, ldw, gamma, beta, db, dW, out_dtype, ld_dw, O, I, dgamma,
                      dbeta, partials, splits);
}
int sf_op_embed_table_grad(void* stream, int dtype, const void* dx, int ld, int B, int T, int S, int D, int mode, const int* tidx, float* out) {
  return embed_table_grad(static_cast<cudaStream_t>(stream), dtype, dx, ld, B, T, S, D, mode, tidx, out);
}
int sf_op_rowperm(void* stream, const void* in, void* out, long long M, int row_bytes, int row_map, int T, int S) {
  return rowperm(static_cast<cudaStream_t>(stream), in, out, static_cast<long>(M), row_bytes, row_map, T, S);
}
int sf_op_attention_backward(void* stream, int dtype, int mode, const void* qkv, int ld_qkv, const void* out, int ld_out, const void* dout,
                             int ld_dout, void* dqkv, int ld_dqkv, int groups, int heads, int L, int T_inner, int causal, float scale) {
  return attention_backward(static_cast<cudaStream_t>(stream), dtype, mode, qkv, ld_qkv, out, ld_out, dout, ld_dout, dqkv, ld_dqkv, groups,
                            heads, L, T_inner, causal, scale);
}
int sf_op_pool_attention_backward(void* stream, int dtype, const void* kv, int ld_kv, const float* q, const void* dout, int ld_dout,
                                  void* dkv, int ld_dkv, float* dq, int frames, int heads, int S) {
  return pool_attention_backward(static_cast<cudaStream_t>(stream), dtype, kv, ld_kv, q, dout, ld_dout, dkv, ld_dkv, dq, frames, heads, S);
}
int sf_op_kv_append(void* stream, int dtype, const void* qkv, int ld_qkv, void* kcache, void* vcache, int Tcap, int sites,
                    int heads, int Tq, int pos0) {
  return kv_append(static_cast<cudaStream_t>(stream), dtype, qkv, ld_qkv, kcache, vcache, Tcap, sites, heads, Tq, pos0);
}
int sf_op_spatial_attention(void* stream, int dtype, const void* qkv, int ld_qkv, void* out, int ld_out, int frames,
                            int heads, int S, int T_inner, float scale, float* probs) {
  return spatial_attention(static_cast<cudaStream_t>(stream), dtype, qkv, ld_qkv, out, ld_out, frames, heads, S, T_inner,
                           scale, probs);
}
int sf_op_rowstats(void* stream, int dtype, const void* x, int ldx, int M, int D, void* stats) {
  return rowstats(static_cast<cudaStream_t>(stream), dtype, x, ldx, M, D, static_cast<float2*>(stats));
}
int sf_op_gemm_stats_parts(int M, int N) { return gemm_stats_parts(M, N); }
int sf_op_pool_probe(void* stream, int dtype, const void* tokens, int ld, const float* u, const void* wv, const float* bv,
                     void* out, int ld_out, int frames, int heads, int S) {
  return pool_probe(static_cast<cudaStream_t>(stream), dtype, tokens, ld, u, wv, bv, out, ld_out, frames, heads, S);
}
int sf_op_pool_attention(void* stream, int dtype, const void* kv, int ld_kv, const float* q, void* out, int ld_out,
                         int frames, int heads, int S) {
  return pool_attention(static_cast<cudaStream_t>(stream), dtype, kv, ld_kv, q, out, ld_out, frames, heads, S);
}

}  // extern "C"
