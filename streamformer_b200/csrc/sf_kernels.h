// sf_kernels.h — internal launcher interface between the StreamFormer runtime (runtime.cu /
// c_abi.cu) and the hand-written sm_100a kernels. All launchers are asynchronous on `stream`,
// never allocate, and return 0 or a negative sf_status.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace sf {

// kU8 / kU8HWC are PIXEL formats only: planar [.., C, H, W] and interleaved [.., H, W, C] uint8 frames
enum DType : int { kBF16 = 0, kF16 = 1, kF32 = 2, kU8 = 3, kU8HWC = 4 };
inline size_t dtype_size(int dt) { return dt == kF32 ? 4 : (dt == kU8 || dt == kU8HWC) ? 1 : 2; }

enum RowMap : int {
  kRowIdentity = 0,
  kRowBTNtoBNT = 1,  // GEMM row m=(b*T+t)*S+n  -> output row (b*S+n)*T+t
  kRowBNTtoBTN = 2,  // GEMM row m=(b*S+n)*T+t  -> output row (b*T+t)*S+n
};
enum Act : int { kActNone = 0, kActGeluErf = 1, kActGeluTanh = 2 };

// Epilogue of  out[r, :] = f(A[m, :] . W^T)  with r = row_map(m).
//   v = acc + bias ; v = act(v) ; v += pos[n] + time[tidx(t)] ; v = residual[r] + tanh(*gate) * v
struct GemmEpilogue {
  const float* bias = nullptr;      // [N] fp32
  int act = kActNone;
  const void* residual = nullptr;   // activation dtype, row stride ldr, indexed by OUTPUT row
  int ldr = 0;
  const float* gate = nullptr;      // device scalar (pre-tanh); nullptr => scale 1
  int row_map = kRowIdentity;
  int T = 1;                        // frames per clip   (row_map / pos / time decomposition)
  int S = 1;                        // sites (patches) per frame
  const float* pos = nullptr;       // [S, N] fp32, added per site n   (rows in (b,t,n) order)
  const float* time_emb = nullptr;  // [time_len, N] fp32, added per frame t
  int time_len = 0;                 // rows in the time table (config.num_frames)
  int time_total = 0;               // total frames the table is stretched over (>= time_off+T)
  int time_off = 0;                 // frames already seen (streaming)
  // streaming under a CUDA graph: time_off = *time_off_dev and
  // time_total = max(time_horizon, time_off + T), read on the device at run time
  const int* time_off_dev = nullptr;
  int time_horizon = 0;
  // LayerNorm folded into this GEMM: A holds the RAW rows x, W was pre-scaled by gamma at bind
  // time, bias is b + W.beta, and the epilogue finishes the normalisation per row:
  //   v = rstd[m] * (acc - mean[m] * ln_colsum[n]) + bias[n]
  // with mean/rstd from the partial (sum, sum of squares) of row m written by the producer of x.
  const float2* ln_stats = nullptr; // [ln_parts][M]
  int ln_parts = 0;
  const float* ln_colsum = nullptr; // [N] fp32: sum_k W'[n,k] of the packed (rounded) matrix
  float ln_eps = 0.f;
  // Partial row statistics of the OUTPUT rows (values as rounded to the activation dtype), one
  // (sum, sumsq) per row and column group, for the next folded LayerNorm: [gemm_stats_parts(M,N)][M]
  float2* stats_out = nullptr;
  // W is a bound (packed) weight that no kernel launched right before this GEMM writes: the producer warp may start
  // the weight half of its first operand stages BEFORE griddepcontrol.wait, under the previous kernel's tail.  The
  // launcher drops the flag when the previous launch of the process was a weight-writing (bind / fold) kernel.
  bool w_static = false;
};

// C[M,N] = A[M,K] . W[N,K]^T on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
// A: row-major, leading dim lda (elements); W: row-major [N,K], leading dim ldw; out leading dim ldo.
// Requirements: K % 8 == 0, N % 8 == 0, lda/ldw/ldo/ldr % 8 == 0, 16-byte aligned pointers.
int gemm(cudaStream_t stream, int dtype, const void* A, int lda, const void* W, int ldw, void* out,
         int ldo, int M, int N, int K, const GemmEpilogue& epi);

// Several dependent GEMMs over the same M rows as ONE persistent launch (see gemm_tcgen05.cu,
// "GEMM chains"): call i may read the output rows and the row statistics (stats_out -> ln_stats,
// gemm_chain_stats_parts(calls, n, N) partials per row) of call i-1 and, as residual, rows written by any
// earlier call of the chain.  `counters` is gemm_chain_counter_bytes(M) of device memory that must
// be zero before the first launch; the kernel leaves it zero again (no host-side state, so a
// captured launch can be replayed).
struct GemmCall {
  const void* A; int lda;
  const void* W; int ldw;
  void* out; int ldo;
  int M, N, K;
  GemmEpilogue epi;
};
bool gemm_chain_supported(int dtype, const GemmCall* calls, int n);
void set_gemm_chain(int on);   // -1: environment default (SF_GEMM_CHAIN, off), 0 / 1: force
void set_spatial_row(int on);  // -1: environment default (SF_SPATIAL_ROW, on): row-in-registers spatial kernel at S = 196
void set_decode_tma(int on);   // -1: environment default (SF_DECODE_TMA, off): TMA-ring / mma.sync streaming decode kernel
int gemm_chain_stats_parts(const GemmCall* calls, int n, int N);
size_t gemm_chain_counter_bytes(int M);
int gemm_chain(cudaStream_t stream, int dtype, const GemmCall* calls, int n, void* counters);

// Number of column groups (partials per row) gemm() writes to GemmEpilogue::stats_out for an M x N output.
int gemm_stats_parts(int M, int N);
// Upper bound of gemm_stats_parts over all tile shapes (columns per partial >= 64).
inline int gemm_stats_parts_max(int N) { return (N + 63) / 64; }

// stats[m] = (sum_d x[m,d], sum_d x[m,d]^2): a one-partial row-statistics table for rows that were
// not produced by gemm() (block-level API entry points).
int rowstats(cudaStream_t stream, int dtype, const void* x, int ldx, int M, int D, float2* stats);

// y = LayerNorm(x) * gamma + beta over the last dim D (fp32 statistics, biased variance).
// row_map as in GemmEpilogue (input row m -> output row r), T/S for the decomposition.
int layernorm(cudaStream_t stream, int dtype, const void* x, int ldx, const float* gamma,
              const float* beta, float eps, void* y, int ldy, int M, int D, int row_map, int T,
              int S);

// pixels [BT, C, H, W] (pix_dtype: bf16/f16/f32) -> patch-major A[(bt*S + n), C*P*P] in act dtype,
// K ordered (c, kh, kw) like Conv2d.weight.reshape(D, -1).  uint8 pixels (kU8: [BT, C, H, W], kU8HWC:
// [BT, H, W, C], C <= 4) are normalised on the way: (x / 255 - mean[c]) / std[c], evaluated in fp32 in
// exactly that order (ClipToTensor + Normalize of the reference's loaders), mean/std host arrays of 4.
int im2col_patches(cudaStream_t stream, int pix_dtype, const void* pixels, int act_dtype, void* out,
                   int BT, int C, int H, int W, int P, const float* mean = nullptr, const float* std = nullptr);

// Temporal attention over frames, one (site, head) per warp task.
//   q rows: qkv[(site*Tq + i), h*64 + d]                (row stride ld_qkv; q at col 0)
//   keys/values: if kcache == nullptr, K/V come from the same qkv buffer (cols D.., 2D..) with
//   Tk == Tq rows per site; else from the cache [site][head][Tcap][64] with Tk valid rows
//   (the new rows must already be appended).  q row i attends to keys j <= q_off + i when
//   causal, all Tk keys otherwise.  out[(site*Tq + i), h*64 + d], row stride ld_out.
//   seen_dev (optional, cache only): device int holding the frames cached before this call; when
//   given the kernel takes q_off = *seen_dev and Tk = q_off + Tq from it, so one captured CUDA
//   graph serves every streaming step.
int temporal_attention(cudaStream_t stream, int dtype, const void* qkv, int ld_qkv, const void* kcache,
                       const void* vcache, int Tcap, void* out, int ld_out, int sites, int heads,
                       int Tq, int Tk, int q_off, int causal, float scale, const int* seen_dev = nullptr);

// Streaming decode (Tq == 1 with a cache): kv_append + temporal_attention in one kernel (attention.cu): by default
// the register-direct FMA kernel (histories global -> registers with 16-byte loads), with set_decode_tma(1) /
// SF_DECODE_TMA=1 the TMA-ring + mma.sync kernel.  Supported for caches of up to 96 frames.
bool temporal_decode_supported(int Tcap, int Tq);
int temporal_decode(cudaStream_t stream, int dtype, const void* qkv, int ld_qkv, void* kcache, void* vcache, int Tcap,
                    void* out, int ld_out, int sites, int heads, int seen, float scale, const int* seen_dev = nullptr);

// Append the K and V slices of qkv (rows (site*Tq + i)) into cache[site][head][pos0 + i][64].
int kv_append(cudaStream_t stream, int dtype, const void* qkv, int ld_qkv, void* kcache, void* vcache,
              int Tcap, int sites, int heads, int Tq, int pos0, const int* seen_dev = nullptr);

// Spatial attention inside each frame: q|k|v column blocks of width heads*64; full (non-causal)
// softmax over the S keys of the frame.  Row of token n of frame f (same for qkv and out):
//   T_inner <= 1 : f*S + n                      (frames contiguous, (b,t,n) order)
//   T_inner  > 1 : (b*S + n)*T_inner + t, f = b*T_inner + t   (the residual stream's (b,n,t) order:
//                  the frame is read in place with a row stride of T_inner, no permute copy)
// probs (optional, fp32 [frames, heads, S, S]) receives the attention probabilities.
int spatial_attention(cudaStream_t stream, int dtype, const void* qkv, int ld_qkv, void* out,
                      int ld_out, int frames, int heads, int S, int T_inner, float scale, float* probs);

// tcgen05 implementation of spatial_attention (attention_tc.cu) for frames of up to 208 tokens;
// spatial_attention() dispatches to it when supported (SF_SPATIAL_TC=0 forces the mma.sync kernel).
bool spatial_attention_tc_supported(int ld_qkv, int S);
int spatial_attention_tc(cudaStream_t stream, int dtype, const void* qkv, int ld_qkv, void* out, int ld_out,
                         int frames, int heads, int S, int T_inner, float scale);

// Attention-pooling core of the SigLIP head: for each frame and head, softmax_n(q_h . K[n,h]) V[n,h].
//   kv rows (frame*S + n): [K (heads*64) | V (heads*64)], q: [heads*64] fp32 (already scaled).
//   out [frames, heads*64] in act dtype.
int pool_attention(cudaStream_t stream, int dtype, const void* kv, int ld_kv, const float* q,
                   void* out, int ld_out, int frames, int heads, int S);

// The same pooling attention with the key / value projections collapsed into it (the query is one
// learned probe): tokens [frames*S, D] (row stride ld) -> out [frames, D] = concat_h(W_v,h s_h + b_v,h)
// with s_h = softmax_n(x_n . u_h)-weighted token sum; u [heads, D] fp32 = W_k,h^T q_h (q scaled),
// wv [D, D] row-major in the activation dtype, bv [D] fp32.
int pool_probe(cudaStream_t stream, int dtype, const void* tokens, int ld, const float* u, const void* wv, const float* bv,
               void* out, int ld_out, int frames, int heads, int S);

// SigLIP task head on the (gathered) last-frame pooler_output: logits = exp(*logit_scale) * <x^, t^> + *logit_bias,
// loss += sum(-logsigmoid(label * logit)) / loss_div, optionally d loss / d logits and the two scalar parameter
// gradients (attention.cu).  Labels: +1 at targets[i] (classification) or at column i + diag_offset (contrastive;
// diag_offset < 0: negatives only), -1 elsewhere.  `loss` and `dparams` are accumulated into (zero them first).
int siglip_head(cudaStream_t stream, int dtype, const void* image, int ld_i, const void* text, int ld_t, int B, int L, int D,
                const float* logit_scale, const float* logit_bias, int norm_image, int norm_text, const long long* targets,
                int diag_offset, float loss_div, float* logits, int ld_l, float* loss, void* dlogits, int ld_d, float* dparams);

// dx = (g - x^ (x^ . g)) / |x| with g = exp(*gscale) * dxhat: backward of x^ = x / |x| (rows of [B, D])
int l2norm_backward(cudaStream_t stream, int dtype, const void* x, int ldx, const void* dxhat, int ldg, const float* gscale,
                    void* dx, int ldo, int B, int D);

// ---- backward pass (backward.cu, attention.cu): see the kernel comments for the formulas
int transpose2d(cudaStream_t st, int dtype, const void* in, int ld_in, void* out, int ld_out, int M, int N);
int colsum(cudaStream_t st, int dtype, const void* x, int ld, int M, int N, float* out);
int ln_backward(cudaStream_t st, int dtype, const void* x, int ldx, const void* dn, int ldn, float eps, const void* dres, int ldr, void* dx,
                int ldo, int M, int D);
int ln_affine_backward(cudaStream_t st, int dtype, const void* x, int ldx, const void* dy, int ldy, const float* gamma, float eps, void* dx,
                       int ldo, int M, int D, int row_map, int Tn, int Sn, float* dgamma, float* dbeta);
int gelu_forward(cudaStream_t st, int dtype, const void* a, void* h, long n, int act);
int gelu_backward(cudaStream_t st, int dtype, void* a_h, void* dh_dpre, long n, int act);
int gate_backward(cudaStream_t st, int dtype, const void* dx, const void* y, const float* gate, void* dy, long n, float* dgate);
int wfold_finish(cudaStream_t st, int dtype, const void* G, int ldg, const void* Wp, int ldw, const float* gamma, const float* beta,
                 const float* db, void* dW, int out_dtype, int ldo, int O, int I, float* dgamma, float* dbeta,
                 const float* Gf = nullptr, int splits = 0);   // Gf: fp32 split-K partials [splits][O][ldg] instead of G
// G[O, I] = dY^T . X on tcgen05 with MN-major operands straight from the row-major activations (wgrad_tcgen05.cu);
// partials: fp32 [wgrad_splits(M, O, I)][O][I], summed by wfold_finish
int wgrad_splits(int M, int O, int I);
int wgrad(cudaStream_t stream, int dtype, const void* dY, int ldy, const void* X, int ldx, int M, int O, int I, float* partials);
int embed_table_grad(cudaStream_t st, int dtype, const void* dx, int ld, int B, int Tn, int Sn, int D, int mode, const int* tidx, float* out);
int rowperm(cudaStream_t st, const void* in, void* out, long M, int row_bytes, int row_map, int Tn, int Sn);
int attention_backward(cudaStream_t stream, int dtype, int mode, const void* qkv, int ld_qkv, const void* out, int ld_o, const void* dout,
                       int ld_do, void* dqkv, int ld_dq, int groups, int heads, int L, int T_inner, int causal, float scale);
int pool_attention_backward(cudaStream_t stream, int dtype, const void* kv, int ld_kv, const float* q, const void* dout, int ld_do,
                            void* dkv, int ld_dkv, float* dq, int frames, int heads, int S);

// out = cast(in)  (weight / bias packing helpers; n elements)
int cast(cudaStream_t stream, int src_dtype, const void* src, int dst_dtype, void* dst, size_t n);

// Number of kernel launches issued by this library since load (all launchers increment it).
uint64_t launch_count();
void count_launch(int n = 1);

// Optional in-situ profiling: when enabled every launcher brackets its kernel with CUDA events on
// the launching stream; collect() synchronises and sums elapsed time per kernel class.
enum ProfClass : int { kProfGemm = 0, kProfLayerNorm, kProfIm2col, kProfTemporalAttn, kProfSpatialAttn,
                       kProfPoolAttn, kProfKvAppend, kProfOther, kProfNumClasses };
void prof_enable(bool on);
void prof_set_mode(int mode);   // 0 off, 1 per-kernel events, 2 per-phase events only
bool prof_enabled();
// Phase timing (mode 2): one event pair around each phase of the forward, kernels inside run
// back to back exactly as in production (no per-kernel events, PDL overlap intact).
enum Phase : int { kPhaseEmbed = 0, kPhaseAttnBlock, kPhaseMlp, kPhaseHead, kPhaseNum };
bool phase_prof_enabled();
void phase_begin(cudaStream_t st, int phase);
void phase_end(cudaStream_t st);
int phase_collect(double* ms, long long* count, int n);
struct PhaseScope {
  cudaStream_t st; bool on;
  PhaseScope(cudaStream_t s, int phase) : st(s), on(phase_prof_enabled()) { if (on) phase_begin(st, phase); }
  ~PhaseScope() { if (on) phase_end(st); }
};
void prof_begin(cudaStream_t st, int cls, double flops, double bytes);
void prof_end(cudaStream_t st);
int prof_collect(double* ms, double* flops, double* bytes, long long* launches, int n);
struct ProfScope {
  cudaStream_t st; bool on;
  ProfScope(cudaStream_t s, int cls, double flops = 0.0, double bytes = 0.0) : st(s), on(prof_enabled()) {
    if (on) prof_begin(st, cls, flops, bytes);
  }
  ~ProfScope() { if (on) prof_end(st); }
};

// Launch configuration carrying the library-wide attributes: programmatic stream serialization
// (PDL; SF_PDL=0 disables it) and, optionally, a CTA-cluster dimension.
struct LaunchCfg {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  LaunchCfg(dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster = 1);
};

const char* last_error();
void set_error(const char* fmt, ...);

}  // namespace sf
