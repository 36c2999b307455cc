/* streamformer_b200.h — C ABI of the B200-native StreamFormer video-encoder hot path.
 *
 * The reference (Go2Heart/StreamFormer) has no FFI on this path: the boundary is the Python class
 * TimesformerMultiTaskingModelSigLIP (models/modeling_timesformer_siglip.py:1241-1354) and its
 * KV-cache twin (downstream/VideoQA/llava/model/multimodal_encoder/timesformer_encoder.py:1255-1392).
 * This header is what that class binds instead of its eager torch ops: plain C, raw device pointers,
 * sizes and a cudaStream_t; no ATen / pybind types.  The ctypes stub that binds it lives in
 * streamformer_b200/_native.py and is reproduced in INTEGRATION.md.
 *
 * Conventions
 *  - every function returns 0 on success or a negative sf_status; sf_last_error() returns a
 *    thread-local message describing the last failure;
 *  - all device work is enqueued on the `stream` argument (a cudaStream_t passed as void*);
 *    nothing synchronises the host and nothing allocates inside the forward calls;
 *  - a context is bound to one device and is NOT thread-safe (same as the reference: one model
 *    replica per process / GPU);
 *  - activations are bf16 or fp16 (sf_config.dtype); residual stream rows are ordered (b, n, t)
 *    exactly like the reference's hidden_states [B, N*T, D] (…siglip.py:452-454), and stay in that
 *    order through every kernel of a layer (the spatial attention reads its frame with a row stride
 *    of T instead of the reference's permute copies, …siglip.py:962-991).
 */
#ifndef STREAMFORMER_B200_H_
#define STREAMFORMER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum sf_status {
  SF_OK = 0,
  SF_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  SF_ERR_CUDA = -2,      /* a CUDA runtime call or kernel launch failed */
  SF_ERR_DRIVER = -3,    /* tensor-map encode / driver entry point failure */
  SF_ERR_STATE = -4,     /* weights not bound, cache overflow, ... */
  SF_ERR_WORKSPACE = -5  /* caller-provided workspace too small */
} sf_status;

/* SF_U8 / SF_U8_HWC are pixel formats only (planar [B,T,C,H,W] / interleaved [B,T,H,W,C] uint8 frames,
 * normalised on the GPU: the input edge of extract_oad_feature.py:42-48, 124-130 and
 * datasets/kinetics_sparse.py:110-118 moved behind the boundary) */
typedef enum sf_dtype { SF_BF16 = 0, SF_F16 = 1, SF_F32 = 2, SF_U8 = 3, SF_U8_HWC = 4 } sf_dtype;
typedef enum sf_act { SF_ACT_NONE = 0, SF_ACT_GELU = 1, SF_ACT_GELU_TANH = 2 } sf_act;
typedef enum sf_rowmap { SF_ROW_IDENTITY = 0, SF_ROW_BTN_TO_BNT = 1, SF_ROW_BNT_TO_BTN = 2 } sf_rowmap;

/* Mirrors StreamformerConfig (models/configuration_streamformer.py:92-137). */
typedef struct sf_config {
  int image_size;            /* 224 */
  int patch_size;            /* 16  */
  int num_channels;          /* 3   */
  int num_frames;            /* rows of embeddings.time_embeddings (16) */
  int hidden_size;           /* 768 */
  int num_hidden_layers;     /* 12  */
  int num_attention_heads;   /* 12 (head dim must be 64) */
  int intermediate_size;     /* 3072 */
  int hidden_act;            /* sf_act: SF_ACT_GELU ("gelu", erf) or SF_ACT_GELU_TANH */
  float layer_norm_eps;      /* 1e-6 */
  int causal_temporal;       /* config.enable_causal_temporal */
  int dtype;                 /* sf_dtype of activations and packed matrices: SF_BF16 or SF_F16 */
  int fold_temporal_proj;    /* 1: pre-multiply temporal_dense . temporal_attention.output.dense */
} sf_config;

/* One reference state-dict tensor (names exactly as in SURVEY.md 8(b), without any
 * "timesformer." prefix). Pointers are BORROWED for the duration of sf_bind_weights only: the
 * library packs (casts, merges LoRA, folds) into context-owned memory. */
typedef struct sf_weight_desc {
  const char* name;
  const void* data;   /* device pointer */
  int dtype;          /* sf_dtype */
  int ndim;
  int64_t shape[4];
} sf_weight_desc;

typedef struct sf_ctx sf_ctx;
typedef struct sf_kv sf_kv;

const char* sf_last_error(void);
const char* sf_version(void);
/* kernels launched by this library since it was loaded */
uint64_t sf_launch_count(void);

/* Library-wide switches.  "gemm_chain": 1 runs GEMMs that follow each other over the same rows
 * (out-proj -> fc1 -> fc2 -> next QKV ...) as one persistent launch with in-kernel row dependencies,
 * 0 one launch per GEMM, -1 the default (SF_GEMM_CHAIN environment variable, off).
 * "stream_graph": 0 serves streaming steps with direct launches, 1 from the captured CUDA graph, -1 the
 * default (SF_STREAM_GRAPH environment variable, on).
 * "dual_stream": 1 runs a one-shot forward of an even batch as two half batches on two streams (the caller's and
 * a context-owned one, forked / joined with events) so that one half's kernel fill and drain overlap the other
 * half's steady state; 0 single stream; -1 the default (SF_DUAL_STREAM environment variable, off: measured
 * no gain on B200, 19.5 k vs 19.6 k frames/s at cfg2).  Profiling
 * modes, hidden-state / attention outputs, the KV cache and stream capture always use the single-stream schedule.
 * "spatial_row": kernel choice of the spatial attention at 193..200 tokens per frame (224 x 224): 1 the kernel that
 * keeps whole score rows in registers, 0 the general two-pass tcgen05 kernel, -1 the default (SF_SPATIAL_ROW, on).
 * "decode_tma": streaming decode kernel: 1 TMA-staged histories + mma.sync, 0 register-direct FMA kernel, -1 the
 * default (SF_DECODE_TMA, off). */
int sf_set_option(const char* name, int value);

/* In-situ profiling: when enabled every kernel launch is bracketed by CUDA events on its stream.
 * sf_profile_collect synchronises the device and returns, per kernel class (0 gemm, 1 layernorm,
 * 2 im2col, 3 temporal attention, 4 spatial attention, 5 pooling attention, 6 kv append, 7 other),
 * the summed duration [ms], executed FLOPs, algorithmic bytes and launch count since the last call. */
#define SF_PROFILE_CLASSES 8
int sf_profile(int mode);   /* 0 off, 1 per-kernel events, 2 per-phase events */
int sf_profile_collect(double* ms, double* flops, double* bytes, long long* launches, int n_classes);
/* Phase timing (mode 2): one event pair around each phase of sf_forward, the kernels inside run back
 * to back as in production.  Phases: 0 embedding (im2col + patch GEMM), 1 space-time attention block
 * of a layer (temporal QKV, temporal attention, out-proj.temporal_dense + gate, spatial QKV, spatial
 * attention, out-proj), 2 MLP of a layer, 3 post-LN + pooling head.  ms / count are summed since the
 * last call (count = number of phase instances, e.g. layers). */
#define SF_PROFILE_PHASES 4
int sf_profile_collect_phases(double* ms, long long* count, int n_phases);

/* ---- model lifetime ---------------------------------------------------------------------- */
/* replaces TimesformerMultiTaskingModelSigLIP.__init__ (…siglip.py:1244-1258) */
int sf_create(const sf_config* cfg, int device, sf_ctx** out);
int sf_destroy(sf_ctx* ctx);
/* replaces load_state_dict / from_pretrained weight materialisation; may be called again after an
 * optimiser step to re-pack.  Parameter groups bind independently — embeddings.*, encoder.layer.{i}.*,
 * post_layernorm.*, head.* — so a sub-module composed inside another model (downstream/AR/models/
 * modeling_timesformer_video_classification.py:42-56) binds only what it owns; an entry point whose
 * group is absent fails with SF_ERR_STATE. Optional LoRA tensors (…qkv_lora_{a,b}.weight, …dense_lora_{a,b}.weight,
 * …siglip.py:632-647, 731-746) are merged W + B.A. */
int sf_bind_weights(sf_ctx* ctx, void* stream, const sf_weight_desc* w, int n);
/* replaces the loaders' ClipToTensor + Normalize(mean, std) for uint8 pixels (extract_oad_feature.py:42-48):
 * pixel -> (x / 255 - mean[c]) / std[c], evaluated in fp32 in that order.  Default 0.5 / 0.5; n <= 4. */
int sf_set_pixel_norm(sf_ctx* ctx, const float* mean, const float* std, int n);
/* interpolated position table for a non-default resolution (…siglip.py:380-411): fp32 [S, D] */
int sf_set_pos_embed(sf_ctx* ctx, void* stream, const float* pos, int S);

/* ---- full forward ------------------------------------------------------------------------ */
int sf_workspace_bytes(const sf_ctx* ctx, int B, int T, int H, int W, size_t* out);
/* replaces TimesformerMultiTaskingModelSigLIP.forward (…siglip.py:1299-1354).
 *   pixels        [B, T, C, H, W], pixels_dtype in {bf16, f16, f32, u8}; or [B, T, H, W, C] with SF_U8_HWC
 *   last_hidden   [B, T, S, D]  activation dtype            (last_hidden_state)
 *   pooler        [B, T, D]     activation dtype            (pooler_output)
 *   hidden_states NULL or L+1 device pointers, each [B, S*T, D] (rows (b,n,t)), filled in order
 *   attentions    NULL or L device pointers, each fp32 [B*T, heads, S, S] (spatial probabilities) */
int sf_forward(sf_ctx* ctx, void* stream, const void* pixels, int pixels_dtype, int B, int T, int H,
               int W, void* last_hidden, void* pooler, void* const* hidden_states,
               void* const* attentions, void* workspace, size_t workspace_bytes);

/* ---- streaming (temporal KV cache) ------------------------------------------------------- */
/* replaces transformers.DynamicCache as used by the KV twin (…timesformer_encoder.py:517-518,
 * 1340-1349): pre-allocated [layer][K|V][B*S sites][heads][max_frames][64], appended in place.
 * time_horizon: 0 = reference rule (time table stretched over frames seen so far when that exceeds
 * num_frames, …timesformer_encoder.py:336-366); >0 = fixed horizon, making streaming identical to
 * a one-shot forward of `time_horizon` frames. */
int sf_kv_create(sf_ctx* ctx, int B, int S, int max_frames, int time_horizon, sf_kv** out);
int sf_kv_reset(sf_kv* kv);
int sf_kv_destroy(sf_kv* kv);
int sf_kv_seq_len(const sf_kv* kv);
/* block-level streaming (sf_embed_forward + sf_layer_forward with a cache): every layer of a step appends
 * at the same position; the caller advances the cache by the step's frames once, after the last layer
 * (sf_forward_stream does this itself) */
int sf_kv_advance(sf_kv* kv, int frames);
int sf_kv_capacity(const sf_kv* kv);
/* steps of this cache that were served by replaying a captured CUDA graph (see sf_forward_stream) */
long long sf_kv_graph_launches(const sf_kv* kv);
/* forward of T_new frames that attend to every cached frame (+ causal order among the new ones);
 * outputs as sf_forward with T = T_new. Advances the cache by T_new.
 * From the second call with the same (B, T_new, H, W, dtype, workspace) on, the step is replayed
 * from a CUDA graph captured once: its kernels read the stream position from a device counter the
 * graph itself advances, so every step of every stream reuses one executable graph and the host
 * cost per step is the D2D staging copies (pixels in; last_hidden, pooler and, when asked for, the L+1
 * hidden states out) + one graph launch (SF_STREAM_GRAPH=0 / sf_set_option("stream_graph", 0) disable;
 * the profiling modes use direct launches). */
int sf_forward_stream(sf_ctx* ctx, void* stream, sf_kv* kv, const void* pixels, int pixels_dtype,
                      int B, int T_new, int H, int W, void* last_hidden, void* pooler,
                      void* const* hidden_states, void* workspace, size_t workspace_bytes);

/* ---- block-level API (what downstream/AR and the OVIS ViT-adapter call) ------------------- */
/* TimesformerEmbeddingsSigLIP.forward (…siglip.py:413-457): pixels -> x [B, S*T, D] */
int sf_embed_forward(sf_ctx* ctx, void* stream, const void* pixels, int pixels_dtype, int B, int T,
                     int H, int W, int time_off, int time_total, void* x_out, void* workspace,
                     size_t workspace_bytes);
/* TimesformerLayerSigLIP.forward (…siglip.py:900-1004): x_in/x_out [B, S*T, D]; x_out may alias
 * x_in; kv may be NULL; attn_probs NULL or fp32 [B*T, heads, S, S] */
int sf_layer_forward(sf_ctx* ctx, void* stream, int layer, const void* x_in, void* x_out, int B,
                     int T, int S, sf_kv* kv, float* attn_probs, void* workspace,
                     size_t workspace_bytes);
/* TimesformerEncoder.forward (…siglip.py:1019-1063): all layers over x_in [B, S*T, D].  hidden_states NULL
 * (result in x_out, which may not alias x_in unless the model has one layer) or L+1 pointers of which
 * [1..L] receive the layer outputs ([0] is the caller's x_in, not written; x_out is ignored then);
 * attentions as sf_forward; kv NULL or a cache, advanced by T once all layers ran. */
int sf_encoder_forward(sf_ctx* ctx, void* stream, const void* x_in, int B, int T, int S, sf_kv* kv,
                       void* x_out, void* const* hidden_states, void* const* attentions, void* workspace,
                       size_t workspace_bytes);
/* post_layernorm + (b,n,t)->(b,t,n) (…siglip.py:1330-1346): x [B, S*T, D] -> [B, T, S, D] */
int sf_final_norm(sf_ctx* ctx, void* stream, const void* x, int B, int T, int S, void* last_hidden);
/* TimesformerSiglipMultiheadAttentionPoolingHead.forward (…siglip.py:1141-1154):
 * tokens [frames, S, D] -> pooled [frames, D] */
int sf_head_forward(sf_ctx* ctx, void* stream, const void* tokens, int frames, int S, void* pooled,
                    void* workspace, size_t workspace_bytes);

/* ---- single kernels (parity tests, profiling) -------------------------------------------- */
typedef struct sf_gemm_epilogue {
  const float* bias; int act;
  const void* residual; int ldr; const float* gate;
  int row_map; int T; int S;
  const float* pos; const float* time_emb; int time_len; int time_total; int time_off;
  /* LayerNorm folded into the GEMM (nn.LayerNorm + nn.Linear pairs, …siglip.py:943+578, 974+691, 997+820):
   * A holds the raw rows, W is pre-scaled by gamma, bias = b + W.beta, and
   *   out = rstd[m] * (A.W^T - mean[m] * ln_colsum[n]) + bias[n]
   * with mean/rstd from ln_parts partial (sum, sumsq) float pairs per row: ln_stats[part][M][2]. */
  const void* ln_stats; int ln_parts; const float* ln_colsum; float ln_eps;
  /* optional by-product of the residual / embed epilogues: partial (sum, sumsq) of every output row,
   * [sf_op_gemm_stats_parts(M, N)][M][2] floats, in the layout ln_stats consumes */
  void* stats_out;
} sf_gemm_epilogue;
int sf_op_gemm(void* stream, int dtype, const void* A, int lda, const void* W, int ldw, void* out,
               int ldo, int M, int N, int K, const sf_gemm_epilogue* epi);
int sf_op_layernorm(void* stream, int dtype, const void* x, int ldx, const float* gamma,
                    const float* beta, float eps, void* y, int ldy, int M, int D, int row_map, int T,
                    int S);
int sf_op_im2col(void* stream, int pix_dtype, const void* pixels, int act_dtype, void* out, int BT,
                 int C, int H, int W, int P);
int sf_op_temporal_attention(void* stream, int dtype, const void* qkv, int ld_qkv, const void* kcache,
                             const void* vcache, int Tcap, void* out, int ld_out, int sites, int heads,
                             int Tq, int Tk, int q_off, int causal, float scale);
/* streaming decode: one new frame per site; kv_append + temporal attention over the cache in one kernel
 * (cache capacity <= 96 frames; `seen` frames are cached before the call, the new row lands at index `seen`) */
int sf_op_temporal_decode(void* stream, int dtype, const void* qkv, int ld_qkv, void* kcache, void* vcache,
                          int Tcap, void* out, int ld_out, int sites, int heads, int seen, float scale);
int sf_op_kv_append(void* stream, int dtype, const void* qkv, int ld_qkv, void* kcache, void* vcache,
                    int Tcap, int sites, int heads, int Tq, int pos0);
/* T_inner <= 1: token n of frame f at row f*S + n; T_inner > 1: at row (b*S + n)*T_inner + t with
 * f = b*T_inner + t (the residual stream's (b,n,t) order, read in place) */
int sf_op_spatial_attention(void* stream, int dtype, const void* qkv, int ld_qkv, void* out, int ld_out,
                            int frames, int heads, int S, int T_inner, float scale, float* probs);
/* stats[m] = (sum_d x[m,d], sum_d x[m,d]^2): the one-partial table a folded LayerNorm consumes */
int sf_op_rowstats(void* stream, int dtype, const void* x, int ldx, int M, int D, void* stats);
int sf_op_gemm_stats_parts(int M, int N);
int sf_op_pool_attention(void* stream, int dtype, const void* kv, int ld_kv, const float* q, void* out,
                         int ld_out, int frames, int heads, int S);
/* The pooling attention of the SigLIP head with its key / value projections collapsed into it (the
 * query is one learned probe, …siglip.py:1141-1148): tokens [frames*S, D] -> out [frames, D] =
 * concat_h(W_v,h s_h + b_v,h), s_h = softmax_n(x_n . u_h)-weighted token sum.  u [heads, D] fp32 =
 * W_k,h^T q_h with q the scaled probe query, wv [D, D] (activation dtype, nn.Linear layout), bv [D] fp32. */
int sf_op_pool_probe(void* stream, int dtype, const void* tokens, int ld, const float* u, const void* wv,
                     const float* bv, void* out, int ld_out, int frames, int heads, int S);

/* ---- task head on the gathered pooler_output (the step after the path) -------- */
/* replaces TimesformerVideoClassificationHead.forward (…siglip.py:1704-1726: image = pooler_output[:, -1],
 * L2-normalised; logits_per_image = exp(logit_scale) * image . text^T + logit_bias; loss =
 * -logsigmoid(target_labels * logits).sum() / B) and SigLipLoss._loss / the retrieval head (…siglip.py:220-243,
 * 2324-2351; both sides normalised, labels 2*eye - 1; with image rows = this rank's clips and text rows = every
 * rank's captions, diag_offset = rank * B_local reproduces the reference's ring exchange of negatives).
 *   image [B, D] / text [L, D] in `dtype` (row strides ld_i / ld_t elements), logit_scale (pre-exp) and
 *   logit_bias device scalars; targets int64 [B] or NULL (then +1 sits at column i + diag_offset; < 0: none);
 *   logits fp32 [B, L] or NULL; *loss += result (zero it first); dlogits [B, L] in `dtype` or NULL receives
 *   d loss / d logits, dparams NULL or float[2] += (d loss / d logit_scale, d loss / d logit_bias). */
int sf_op_siglip_head(void* stream, int dtype, const void* image, int ld_i, const void* text, int ld_t, int B,
                      int L, int D, const float* logit_scale, const float* logit_bias, int normalize_image,
                      int normalize_text, const int64_t* targets, int diag_offset, float loss_div, float* logits,
                      int ld_logits, float* loss, void* dlogits, int ld_dlogits, float* dparams);

/* backward of the head's image normalisation: dx = (g - x^ (x^ . g)) / |x|, g = exp(*gscale) * dxhat */
int sf_op_l2norm_backward(void* stream, int dtype, const void* x, int ldx, const void* dxhat, int ldg,
                          const float* gscale, void* dx, int ldo, int B, int D);

/* ---- backward pass: what torch.autograd does for the reference's training step
 * (tools/finetune_tools.py:543-573) on the graph of …siglip.py:900-1004.  The contractions (dgrad = dY . W,
 * wgrad = dY^T . X) run on sf_op_gemm, fed by sf_op_transpose; the rest are the row-wise / attention kernels
 * below.  streamformer_b200/autograd.py composes them into one torch.autograd.Function around the forward. */
/* packed matrices / vectors of the bound context copied (transposed when `transpose`) into caller memory:
 * name in {"t_qkv","t_out","t_dense","s_qkv","s_out","fc1","fc2"} (layer matrices as the kernels consume them:
 * LayerNorm gamma folded in, LoRA merged; [out, in] row-major, activation dtype), the same with suffix "_b"
 * (fp32 bias vectors, folded where the matrix is), layer = -1 for {"head_kv","head_out","head_fc1","head_fc2",
 * "patch"} and their "_b", and "head_q" (fp32 [D], the scaled probe query). */
int sf_export_packed(sf_ctx* ctx, void* stream, int layer, const char* name, int transpose, void* dst, size_t dst_bytes);
/* out[n, m] = in[m, n]; out row stride ld_out >= round_up(M, 8), the padding columns are zeroed */
int sf_op_transpose(void* stream, int dtype, const void* in, int ld_in, void* out, int ld_out, int M, int N);
/* out[n] = sum_m x[m, n]  (fp32) */
int sf_op_colsum(void* stream, int dtype, const void* x, int ld, int M, int N, float* out);
/* LayerNorm without affine, n = (x - mean) rstd:  dx = rstd (dn - mean(dn) - n mean(dn n)) + dres (dres nullable) */
int sf_op_ln_backward(void* stream, int dtype, const void* x, int ldx, const void* dn, int ld_dn, float eps,
                      const void* dres, int ld_dres, void* dx, int ld_dx, int M, int D);
/* LayerNorm with gamma / beta, y[row_map(m)] = LN(x[m]):  dx, dgamma += sum dy n, dbeta += sum dy (fp32, accumulated) */
int sf_op_ln_affine_backward(void* stream, int dtype, const void* x, int ldx, const void* dy, int ld_dy,
                             const float* gamma, float eps, void* dx, int ld_dx, int M, int D, int row_map, int T, int S,
                             float* dgamma, float* dbeta);
/* h = GELU(a), out of place (the training forward keeps the pre-activation for the backward) */
int sf_op_gelu(void* stream, int dtype, const void* a, void* h, long long n, int act);
/* in place: a -> h = GELU(a), dh -> dpre = dh GELU'(a)   (n elements, multiple of 8; act as sf_act) */
int sf_op_gelu_backward(void* stream, int dtype, void* a_h, void* dh_dpre, long long n, int act);
/* dy = tanh(*gate) dx;  *dgate += (1 - tanh^2(*gate)) <dx, y>   (…siglip.py:954-958) */
int sf_op_gate_backward(void* stream, int dtype, const void* dx, const void* y, const float* gate, void* dy, long long n,
                        float* dgate);
/* parameter gradients of a Linear whose input LayerNorm is folded into it, from G = dY^T . n [O, I] and db [O]:
 * dW = G gamma + db (x) beta (out_dtype), dgamma += colsum(G o W), dbeta += W^T db with W = Wp / gamma;
 * gamma == NULL: plain Linear (dW = G cast to out_dtype) */
int sf_op_wfold_finish(void* stream, int dtype, const void* G, int ldg, const void* Wp, int ldw, const float* gamma,
                       const float* beta, const float* db, void* dW, int out_dtype, int ld_dw, int O, int I,
                       float* dgamma, float* dbeta);
/* weight gradient G = dY^T . X (dY [M, O], X [M, I], row-major activations) on the tcgen05 tensor cores, both operands
 * read MN-major in place (no transposed copies), token rows split over CTAs: fp32 partials
 * [sf_op_wgrad_splits(M, O, I)][O][I], to be summed by sf_op_wfold_finish (pass them as `partials` there) */
int sf_op_wgrad_splits(int M, int O, int I);
int sf_op_wgrad(void* stream, int dtype, const void* dY, int ldy, const void* X, int ldx, int M, int O, int I, float* partials);
/* sf_op_wfold_finish with the gradient matrix given as the fp32 split partials of sf_op_wgrad */
int sf_op_wfold_finish_partials(void* stream, int dtype, const float* partials, int splits, const void* Wp, int ldw,
                                const float* gamma, const float* beta, const float* db, void* dW, int out_dtype, int ld_dw,
                                int O, int I, float* dgamma, float* dbeta);
/* gradients of the embedding tables from dx [B, S*T, D] (rows (b,n,t)): mode 0 position table [S, D],
 * mode 1 time table (out[tidx[t]] += ...), fp32, accumulated */
int sf_op_embed_table_grad(void* stream, int dtype, const void* dx, int ld, int B, int T, int S, int D, int mode,
                           const int* tidx, float* out);
/* out[row_map(m)] = in[m] (rows of row_bytes bytes, multiple of 16) */
int sf_op_rowperm(void* stream, const void* in, void* out, long long M, int row_bytes, int row_map, int T, int S);
/* backward of the temporal (mode 0: groups = sites, L = T) or spatial (mode 1: groups = frames, L = S, rows as in
 * sf_op_spatial_attention) attention core: dqkv (dq|dk|dv, layout of qkv) from qkv, the forward output and dout */
int sf_op_attention_backward(void* stream, int dtype, int mode, const void* qkv, int ld_qkv, const void* out, int ld_out,
                             const void* dout, int ld_dout, void* dqkv, int ld_dqkv, int groups, int heads, int L,
                             int T_inner, int causal, float scale);
/* backward of sf_op_pool_attention: dkv [frames*S, 2*heads*64], dq fp32 [heads*64] accumulated (nullable) */
int sf_op_pool_attention_backward(void* stream, int dtype, const void* kv, int ld_kv, const float* q, const void* dout,
                                  int ld_dout, void* dkv, int ld_dkv, float* dq, int frames, int heads, int S);

#ifdef __cplusplus
}
#endif
#endif /* STREAMFORMER_B200_H_ */
