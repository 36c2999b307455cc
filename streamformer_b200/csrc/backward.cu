// backward.cu — the row-wise / element-wise half of the encoder's backward pass (what
// torch.autograd does for the reference in tools/finetune_tools.py:543-573 on the graph of
// models/modeling_timesformer_siglip.py:900-1004).  The contraction half (dgrad = dY . W, wgrad = dY^T . X)
// runs on the tcgen05 GEMM of gemm_tcgen05.cu, fed by the transposes below; the attention backward
// lives in attention_bwd.cu.  Everything here is HBM-bound: 16-byte accesses, one warp per row where a
// row reduction is involved, fp32 arithmetic.
//
//   transpose           [M, N] -> [N, Mpad] (wgrad operands: the reduction dimension M becomes contiguous)
//   colsum              bias gradients: sum over rows
//   ln_backward         LayerNorm (no affine: the affine part is folded into the next GEMM's weights) + residual
//   ln_affine_backward  LayerNorm with gamma / beta (post_layernorm, pooling-head LN), optional row permutation
//   gelu_backward       h = GELU(a), dpre = dh * GELU'(a)   (both in place)
//   gate_backward       dy = tanh(g) * dx, dg += (1 - tanh^2 g) * <dx, y>     (…siglip.py:954-958)
//   wfold_finish        parameter gradients of a LayerNorm-folded Linear from G = dY^T . n
//   pos / time sums     gradients of the position / time embedding tables
#include <math.h>
#include <stdint.h>

#include <cuda.h>

#include "sf_kernels.h"
#include "sf_ptx.cuh"
#include "sf_tma.h"

namespace sf {
namespace {

template <typename T>
__device__ __forceinline__ void unpack8b(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = Pack2<T>::unpack(w[j]);
    f[2 * j] = t.x; f[2 * j + 1] = t.y;
  }
}
template <typename T>
__device__ __forceinline__ uint4 pack8b(const float (&f)[8]) {
  uint4 o;
  o.x = Pack2<T>::pack(f[0], f[1]); o.y = Pack2<T>::pack(f[2], f[3]);
  o.z = Pack2<T>::pack(f[4], f[5]); o.w = Pack2<T>::pack(f[6], f[7]);
  return o;
}

int done(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s launch failed: %s", what, cudaGetErrorString(e)); return -2; }
  return 0;
}

// ------------------------------------------------------------------------------- transpose
// out[n, m] = in[m, n] for m < M (zeros for M <= m < Mpad).  64 x 64 tiles through shared memory as 32-bit words
// holding (in[m][n], in[m+1][n]) — two consecutive m of one column, i.e. one word of the transposed row: a thread
// reads the same 16-byte chunk of rows m and m+1, interleaves them (PRMT) and stores eight words; the transposed
// rows are read back as words and leave as 16-byte stores.  Row stride 33 words: the word stores see a 2-way bank
// conflict, the word loads none.
__global__ void __launch_bounds__(256) transpose_kernel(const uint16_t* __restrict__ in, long ld_in, uint16_t* __restrict__ out,
                                                        long ld_out, int M, int N, int Mpad) {
  __shared__ uint32_t tile[64][33];
  griddep_wait();
  griddep_launch_dependents();
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  {
    const int mp = threadIdx.x >> 3, c8 = threadIdx.x & 7;     // row pair, 8-column chunk
    const int m = m0 + 2 * mp, n = n0 + c8 * 8;
    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = make_uint4(0u, 0u, 0u, 0u);
    if (n < N) {                                               // N % 8 == 0: whole chunks
      if (m < M) a = *reinterpret_cast<const uint4*>(in + m * ld_in + n);
      if (m + 1 < M) b = *reinterpret_cast<const uint4*>(in + (m + 1) * ld_in + n);
    }
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tile[c8 * 8 + 2 * j][mp] = __byte_perm(aw[j], bw[j], 0x5410);       // (a.lo, b.lo): column n + 2j
      tile[c8 * 8 + 2 * j + 1][mp] = __byte_perm(aw[j], bw[j], 0x7632);   // (a.hi, b.hi): column n + 2j + 1
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int r = i >> 3, q = i & 7;            // r: transposed row (input column), q: group of 8 m values
    const int n = n0 + r, m = m0 + q * 8;
    if (n < N && m < Mpad) {
      uint4 v;
      v.x = tile[r][q * 4]; v.y = tile[r][q * 4 + 1]; v.z = tile[r][q * 4 + 2]; v.w = tile[r][q * 4 + 3];
      *reinterpret_cast<uint4*>(out + n * ld_out + m) = v;
    }
  }
}

// ------------------------------------------------------------------------------- column sums
// out[n] += sum over this CTA's row range; a thread owns one 16-byte chunk (8 columns) and every 8th row of the
// range, so a warp reads 4 rows x 128 contiguous bytes per step.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long ld, int M, int N, float* __restrict__ out) {
  __shared__ float red[8][256 + 8];
  griddep_wait();
  griddep_launch_dependents();
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;     // column chunk (of 32 per CTA), row lane
  const int n = blockIdx.x * 256 + cl * 8;
  const long rows_per = (M + gridDim.y - 1) / gridDim.y;
  const long r0 = rows_per * blockIdx.y, r1 = (r0 + rows_per < M) ? r0 + rows_per : M;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (n < N) {
    for (long m = r0 + rl; m < r1; m += 8) {
      float f[8];
      unpack8b<T>(*reinterpret_cast<const uint4*>(x + m * ld + n), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cl * 8 + j] = acc[j];
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col < N) atomicAdd(out + col, s);
}

// ------------------------------------------------------------------------------- LayerNorm backward
// y = n = (x - mean) * rstd (affine folded elsewhere):  dx = rstd * (dn - mean(dn) - n * mean(dn * n)) [+ dres]
template <typename T, int NCH>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ x, long ldx, const T* __restrict__ dn, long ldn, float eps,
                                                     const T* __restrict__ dres, long ldr, T* __restrict__ dx, long ldo, int M, int D) {
  griddep_wait();
  griddep_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long m = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int nch = D >> 3;
  float v[NCH][8], g[NCH][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nch) {
      unpack8b<T>(*reinterpret_cast<const uint4*>(x + m * ldx + c * 8), v[i]);
      unpack8b<T>(*reinterpret_cast<const uint4*>(dn + m * ldn + c * 8), g[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  }
  const float mean = warp_sum(s) / D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i)
    if (lane + 32 * i < nch)
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; sq = fmaf(d, d, sq); }
  const float rstd = rsqrtf(warp_sum(sq) / D + eps);
  float c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i)
    if (lane + 32 * i < nch)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] = (v[i][j] - mean) * rstd;       // n
        c1 += g[i][j];
        c2 = fmaf(g[i][j], v[i][j], c2);
      }
  c1 = warp_sum(c1) / D;
  c2 = warp_sum(c2) / D;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nch) {
      float o[8], r[8];
      if (dres) unpack8b<T>(*reinterpret_cast<const uint4*>(dres + m * ldr + c * 8), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - c1 - v[i][j] * c2) + (dres ? r[j] : 0.f);
      *reinterpret_cast<uint4*>(dx + m * ldo + c * 8) = pack8b<T>(o);
    }
  }
}

__device__ __forceinline__ long map_row_b(long m, int row_map, int T, int S) {
  if (row_map == kRowBTNtoBNT) { const long n = m % S, bt = m / S; const long t = bt % T, b = bt / T; return (b * S + n) * T + t; }
  if (row_map == kRowBNTtoBTN) { const long t = m % T, bn = m / T; const long n = bn % S, b = bn / S; return (b * T + t) * S + n; }
  return m;
}

// y[map(m)] = n[m] * gamma + beta:  dn = dy * gamma, dx as above, dgamma += sum dy * n, dbeta += sum dy.
// Persistent warps (each lane owns fixed columns, so the parameter gradients accumulate in registers).
template <typename T, int NCH>
__global__ void __launch_bounds__(256) ln_affine_bwd_kernel(const T* __restrict__ x, long ldx, const T* __restrict__ dy, long ldy,
                                                            const float* __restrict__ gamma, float eps, T* __restrict__ dx, long ldo,
                                                            int M, int D, int row_map, int Tn, int Sn, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta) {
  extern __shared__ float red[];      // [2][D]
  griddep_wait();
  griddep_launch_dependents();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nch = D >> 3;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float ag[NCH][8], ab[NCH][8], gm[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ag[i][j] = ab[i][j] = 0.f;
      const int c = lane + 32 * i;
      gm[i][j] = c < nch ? gamma[c * 8 + j] : 0.f;
    }
  for (long m = static_cast<long>(blockIdx.x) * 8 + warp; m < M; m += static_cast<long>(gridDim.x) * 8) {
    const long r = map_row_b(m, row_map, Tn, Sn);
    float v[NCH][8], g[NCH][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nch) {
        unpack8b<T>(*reinterpret_cast<const uint4*>(x + m * ldx + c * 8), v[i]);
        unpack8b<T>(*reinterpret_cast<const uint4*>(dy + r * ldy + c * 8), g[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
    }
    const float mean = warp_sum(s) / D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
      if (lane + 32 * i < nch)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; sq = fmaf(d, d, sq); }
    const float rstd = rsqrtf(warp_sum(sq) / D + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
      if (lane + 32 * i < nch)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[i][j] = (v[i][j] - mean) * rstd;
          ag[i][j] = fmaf(g[i][j], v[i][j], ag[i][j]);
          ab[i][j] += g[i][j];
          g[i][j] *= gm[i][j];                   // dn
          c1 += g[i][j];
          c2 = fmaf(g[i][j], v[i][j], c2);
        }
    c1 = warp_sum(c1) / D;
    c2 = warp_sum(c2) / D;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nch) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - c1 - v[i][j] * c2);
        *reinterpret_cast<uint4*>(dx + m * ldo + c * 8) = pack8b<T>(o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nch)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&red[c * 8 + j], ag[i][j]);
        atomicAdd(&red[D + c * 8 + j], ab[i][j]);
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, red[i]);
    if (dbeta) atomicAdd(dbeta + i, red[D + i]);
  }
}

// ------------------------------------------------------------------------------- GELU
__device__ __forceinline__ void gelu_fwd_bwd(float a, int act, float& h, float& d) {
  if (act == kActGeluErf) {
    const float cdf = 0.5f * (1.0f + erff(a * 0.70710678118654752f));
    h = a * cdf;
    d = cdf + a * 0.3989422804014327f * __expf(-0.5f * a * a);
  } else {   // tanh approximation
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float u = k0 * (a + k1 * a * a * a);
    const float t = tanhf(u);
    h = 0.5f * a * (1.0f + t);
    d = 0.5f * (1.0f + t) + 0.5f * a * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * a * a);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) gelu_bwd_kernel(T* __restrict__ a_h, T* __restrict__ dh_dpre, long n8, int act) {
  griddep_wait();
  griddep_launch_dependents();
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float a[8], g[8];
    unpack8b<T>(reinterpret_cast<const uint4*>(a_h)[i], a);
    unpack8b<T>(reinterpret_cast<const uint4*>(dh_dpre)[i], g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float h, d;
      gelu_fwd_bwd(a[j], act, h, d);
      a[j] = h;
      g[j] *= d;
    }
    reinterpret_cast<uint4*>(a_h)[i] = pack8b<T>(a);
    reinterpret_cast<uint4*>(dh_dpre)[i] = pack8b<T>(g);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const T* __restrict__ a, T* __restrict__ h, long n8, int act) {
  griddep_wait();
  griddep_launch_dependents();
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float v[8];
    unpack8b<T>(reinterpret_cast<const uint4*>(a)[i], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float hh, d;
      gelu_fwd_bwd(v[j], act, hh, d);
      v[j] = hh;
    }
    reinterpret_cast<uint4*>(h)[i] = pack8b<T>(v);
  }
}

// ------------------------------------------------------------------------------- temporal gate
template <typename T>
__global__ void __launch_bounds__(256) gate_bwd_kernel(const T* __restrict__ dx, const T* __restrict__ y, const float* __restrict__ gate,
                                                       T* __restrict__ dy, long n8, float* __restrict__ dgate) {
  __shared__ float red[8];
  griddep_wait();
  griddep_launch_dependents();
  const float tg = tanhf(*gate);
  float acc = 0.f;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float a[8], b[8];
    unpack8b<T>(reinterpret_cast<const uint4*>(dx)[i], a);
    unpack8b<T>(reinterpret_cast<const uint4*>(y)[i], b);
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc = fmaf(a[j], b[j], acc); a[j] *= tg; }
    reinterpret_cast<uint4*>(dy)[i] = pack8b<T>(a);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(dgate, s * (1.0f - tg * tg));
  }
}

// ------------------------------------------------------------------------------- LN-folded Linear: parameter gradients
// Forward: y = n . W'^T + b' with W'[o,i] = W[o,i] gamma[i], b'[o] = b[o] + sum_i W[o,i] beta[i]  (runtime.cu ln_fold_kernel).
// Given G = dY^T . n [O, I] and db' [O]:  dW[o,i] = G[o,i] gamma[i] + db'[o] beta[i],
//   dgamma[i] = sum_o G[o,i] W[o,i],  dbeta[i] = sum_o db'[o] W[o,i]  with W = W' / gamma (the packed, rounded matrix).
// gamma == nullptr: plain Linear, dW = G.
template <typename T, typename TO>
__global__ void __launch_bounds__(256) wfold_finish_kernel(const T* __restrict__ G, const float* __restrict__ Gf, int splits, long ldg,
                                                           const T* __restrict__ Wp, long ldw,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ db, TO* __restrict__ dW, long ldo, int O, int I,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta) {
  griddep_wait();
  griddep_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;       // column
  if (i >= I) return;
  const long rows_per = 16;                                  // gridDim.y = ceil(O / 16): enough CTAs for weight-sized work
  const long o0 = rows_per * blockIdx.y, o1 = (o0 + rows_per < O) ? o0 + rows_per : O;
  const float gm = gamma ? gamma[i] : 1.f, bt = beta ? beta[i] : 0.f;
  const float inv = (gamma && fabsf(gm) > 1e-20f) ? 1.0f / gm : 0.f;
  float ag = 0.f, ab = 0.f;
  for (long o = o0; o < o1; ++o) {
    float g = 0.f;
    if (Gf) {                                   // fp32 split-K partials [splits][O][ldg] of the wgrad kernel
      for (int sp = 0; sp < splits; ++sp) g += Gf[(static_cast<long>(sp) * O + o) * ldg + i];
    } else {
      g = static_cast<float>(G[o * ldg + i]);
    }
    const float dbo = db ? db[o] : 0.f;
    dW[o * ldo + i] = static_cast<TO>(fmaf(g, gm, dbo * bt));
    if (gamma) {
      const float w = static_cast<float>(Wp[o * ldw + i]) * inv;
      ag = fmaf(g, w, ag);
      ab = fmaf(dbo, w, ab);
    }
  }
  if (gamma) {
    if (dgamma) atomicAdd(dgamma + i, ag);
    if (dbeta) atomicAdd(dbeta + i, ab);
  }
}

// ------------------------------------------------------------------------------- embedding tables
// dx rows in (b, n, t) order.  mode 0: out[n, :] += sum_{b,t} dx[(b S + n) T + t, :]   (position table)
//                              mode 1: out[tidx[t], :] += sum_{b,n} dx[(b S + n) T + t, :]   (time table)
template <typename T>
__global__ void __launch_bounds__(96) embed_table_grad_kernel(const T* __restrict__ dx, long ld, int B, int Tn, int Sn, int D, int mode,
                                                              const int* __restrict__ tidx, float* __restrict__ out) {
  griddep_wait();
  griddep_launch_dependents();
  const int c = threadIdx.x;                  // 8-column chunk
  if (c * 8 >= D) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int key = blockIdx.x;                 // n (mode 0) or t (mode 1)
  const int part = blockIdx.y, parts = gridDim.y;
  if (mode == 0) {
    for (int bt = part; bt < B * Tn; bt += parts) {
      const long row = (static_cast<long>(bt / Tn) * Sn + key) * Tn + bt % Tn;
      float f[8];
      unpack8b<T>(*reinterpret_cast<const uint4*>(dx + row * ld + c * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  } else {
    for (int bn = part; bn < B * Sn; bn += parts) {
      const long row = static_cast<long>(bn) * Tn + key;
      float f[8];
      unpack8b<T>(*reinterpret_cast<const uint4*>(dx + row * ld + c * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
  const int dst = mode == 0 ? key : tidx[key];
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(out + static_cast<long>(dst) * D + c * 8 + j, acc[j]);
}

// rows permuted between the (b,t,n) and (b,n,t) orders (row_map as in GemmEpilogue), 16-byte chunks
__global__ void __launch_bounds__(256) rowperm_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long M, int chunks, int row_map,
                                                      int Tn, int Sn) {
  griddep_wait();
  griddep_launch_dependents();
  const long total = M * chunks;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long m = i / chunks;
    const int c = static_cast<int>(i % chunks);
    out[map_row_b(m, row_map, Tn, Sn) * chunks + c] = in[i];
  }
}

}  // namespace

// =================================================================================== launchers
static bool act_dtype_ok(int dtype, const char* what) {
  if (dtype != kBF16 && dtype != kF16) { set_error("%s: dtype must be bf16 or f16", what); return false; }
  return true;
}

int transpose2d(cudaStream_t st, int dtype, const void* in, int ld_in, void* out, int ld_out, int M, int N) {
  if (M <= 0 || N <= 0) return 0;
  if (!act_dtype_ok(dtype, "transpose")) return -1;
  if ((N % 8) || (ld_in % 8) || (ld_out % 8) || ld_out < M) { set_error("transpose: N, ld_in, ld_out must be multiples of 8 and ld_out >= M (M=%d N=%d ld_in=%d ld_out=%d)", M, N, ld_in, ld_out); return -1; }
  const int Mpad = (M + 7) / 8 * 8;
  ProfScope ps(st, kProfOther, 0.0, 4.0 * M * N);
  LaunchCfg lc(dim3(static_cast<unsigned>((Mpad + 63) / 64), static_cast<unsigned>((N + 63) / 64)), dim3(256), 0, st);
  cudaLaunchKernelEx(&lc.cfg, transpose_kernel, reinterpret_cast<const uint16_t*>(in), static_cast<long>(ld_in),
                     reinterpret_cast<uint16_t*>(out), static_cast<long>(ld_out), M, N, Mpad);
  return done("transpose");
}

int colsum(cudaStream_t st, int dtype, const void* x, int ld, int M, int N, float* out) {
  if (N <= 0) return 0;
  if (!act_dtype_ok(dtype, "colsum")) return -1;
  if ((N % 8) || (ld % 8)) { set_error("colsum: N and ld must be multiples of 8"); return -1; }
  cudaMemsetAsync(out, 0, static_cast<size_t>(N) * sizeof(float), st);
  if (M <= 0) return 0;
  const int col_blocks = (N + 255) / 256;
  int parts = (4 * num_sms() + col_blocks - 1) / col_blocks;        // ~4 CTAs per SM in total
  if (parts > (M + 31) / 32) parts = (M + 31) / 32;
  if (parts < 1) parts = 1;
  ProfScope ps(st, kProfOther, 0.0, 2.0 * M * N);
  LaunchCfg lc(dim3(static_cast<unsigned>(col_blocks), static_cast<unsigned>(parts)), dim3(256), 0, st);
  if (dtype == kBF16) cudaLaunchKernelEx(&lc.cfg, colsum_kernel<__nv_bfloat16>, reinterpret_cast<const __nv_bfloat16*>(x), static_cast<long>(ld), M, N, out);
  else cudaLaunchKernelEx(&lc.cfg, colsum_kernel<__half>, reinterpret_cast<const __half*>(x), static_cast<long>(ld), M, N, out);
  return done("colsum");
}

template <typename T>
static int launch_ln_bwd(cudaStream_t st, const void* x, int ldx, const void* dn, int ldn, float eps, const void* dres, int ldr, void* dx,
                         int ldo, int M, int D) {
  LaunchCfg lc(dim3(static_cast<unsigned>((M + 7) / 8)), dim3(256), 0, st);
#define SF_ARGS reinterpret_cast<const T*>(x), static_cast<long>(ldx), reinterpret_cast<const T*>(dn), static_cast<long>(ldn), eps, \
                reinterpret_cast<const T*>(dres), static_cast<long>(ldr), reinterpret_cast<T*>(dx), static_cast<long>(ldo), M, D
  if (D <= 256) cudaLaunchKernelEx(&lc.cfg, ln_bwd_kernel<T, 1>, SF_ARGS);
  else if (D <= 512) cudaLaunchKernelEx(&lc.cfg, ln_bwd_kernel<T, 2>, SF_ARGS);
  else if (D <= 768) cudaLaunchKernelEx(&lc.cfg, ln_bwd_kernel<T, 3>, SF_ARGS);
  else cudaLaunchKernelEx(&lc.cfg, ln_bwd_kernel<T, 4>, SF_ARGS);
#undef SF_ARGS
  return done("ln_backward");
}
int ln_backward(cudaStream_t st, int dtype, const void* x, int ldx, const void* dn, int ldn, float eps, const void* dres, int ldr, void* dx,
                int ldo, int M, int D) {
  if (M <= 0) return 0;
  if (!act_dtype_ok(dtype, "ln_backward")) return -1;
  if ((D % 8) || D > 1024 || (ldx % 8) || (ldn % 8) || (ldo % 8) || (dres && (ldr % 8))) { set_error("ln_backward: D must be a multiple of 8 up to 1024, 16-byte aligned rows"); return -1; }
  ProfScope ps(st, kProfLayerNorm, 0.0, (dres ? 8.0 : 6.0) * M * D);
  return dtype == kBF16 ? launch_ln_bwd<__nv_bfloat16>(st, x, ldx, dn, ldn, eps, dres, ldr, dx, ldo, M, D)
                        : launch_ln_bwd<__half>(st, x, ldx, dn, ldn, eps, dres, ldr, dx, ldo, M, D);
}

template <typename T>
static int launch_ln_affine_bwd(cudaStream_t st, const void* x, int ldx, const void* dy, int ldy, const float* gamma, float eps, void* dx,
                                int ldo, int M, int D, int row_map, int Tn, int Sn, float* dgamma, float* dbeta) {
  long blocks = (static_cast<long>(M) + 7) / 8;
  if (blocks > 2L * num_sms()) blocks = 2L * num_sms();
  LaunchCfg lc(dim3(static_cast<unsigned>(blocks)), dim3(256), static_cast<size_t>(2 * D) * sizeof(float), st);
#define SF_ARGS reinterpret_cast<const T*>(x), static_cast<long>(ldx), reinterpret_cast<const T*>(dy), static_cast<long>(ldy), gamma, eps, \
                reinterpret_cast<T*>(dx), static_cast<long>(ldo), M, D, row_map, Tn, Sn, dgamma, dbeta
  if (D <= 256) cudaLaunchKernelEx(&lc.cfg, ln_affine_bwd_kernel<T, 1>, SF_ARGS);
  else if (D <= 512) cudaLaunchKernelEx(&lc.cfg, ln_affine_bwd_kernel<T, 2>, SF_ARGS);
  else if (D <= 768) cudaLaunchKernelEx(&lc.cfg, ln_affine_bwd_kernel<T, 3>, SF_ARGS);
  else cudaLaunchKernelEx(&lc.cfg, ln_affine_bwd_kernel<T, 4>, SF_ARGS);
#undef SF_ARGS
  return done("ln_affine_backward");
}
int ln_affine_backward(cudaStream_t st, int dtype, const void* x, int ldx, const void* dy, int ldy, const float* gamma, float eps, void* dx,
                       int ldo, int M, int D, int row_map, int Tn, int Sn, float* dgamma, float* dbeta) {
  if (M <= 0) return 0;
  if (!act_dtype_ok(dtype, "ln_affine_backward")) return -1;
  if ((D % 8) || D > 1024 || (ldx % 8) || (ldy % 8) || (ldo % 8)) { set_error("ln_affine_backward: D must be a multiple of 8 up to 1024, 16-byte aligned rows"); return -1; }
  ProfScope ps(st, kProfLayerNorm, 0.0, 6.0 * M * D);
  return dtype == kBF16 ? launch_ln_affine_bwd<__nv_bfloat16>(st, x, ldx, dy, ldy, gamma, eps, dx, ldo, M, D, row_map, Tn > 0 ? Tn : 1, Sn > 0 ? Sn : 1, dgamma, dbeta)
                        : launch_ln_affine_bwd<__half>(st, x, ldx, dy, ldy, gamma, eps, dx, ldo, M, D, row_map, Tn > 0 ? Tn : 1, Sn > 0 ? Sn : 1, dgamma, dbeta);
}

int gelu_backward(cudaStream_t st, int dtype, void* a_h, void* dh_dpre, long n, int act) {
  if (n <= 0) return 0;
  if (!act_dtype_ok(dtype, "gelu_backward")) return -1;
  if (n % 8) { set_error("gelu_backward: element count must be a multiple of 8"); return -1; }
  if (act != kActGeluErf && act != kActGeluTanh) { set_error("gelu_backward: unknown activation %d", act); return -1; }
  long blocks = (n / 8 + 255) / 256;
  if (blocks > 16L * num_sms()) blocks = 16L * num_sms();
  ProfScope ps(st, kProfOther, 0.0, 8.0 * n);
  LaunchCfg lc(dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st);
  if (dtype == kBF16) cudaLaunchKernelEx(&lc.cfg, gelu_bwd_kernel<__nv_bfloat16>, reinterpret_cast<__nv_bfloat16*>(a_h), reinterpret_cast<__nv_bfloat16*>(dh_dpre), n / 8, act);
  else cudaLaunchKernelEx(&lc.cfg, gelu_bwd_kernel<__half>, reinterpret_cast<__half*>(a_h), reinterpret_cast<__half*>(dh_dpre), n / 8, act);
  return done("gelu_backward");
}

int gelu_forward(cudaStream_t st, int dtype, const void* a, void* h, long n, int act) {
  if (n <= 0) return 0;
  if (!act_dtype_ok(dtype, "gelu_forward")) return -1;
  if (n % 8) { set_error("gelu_forward: element count must be a multiple of 8"); return -1; }
  if (act != kActGeluErf && act != kActGeluTanh) { set_error("gelu_forward: unknown activation %d", act); return -1; }
  long blocks = (n / 8 + 255) / 256;
  if (blocks > 16L * num_sms()) blocks = 16L * num_sms();
  ProfScope ps(st, kProfOther, 0.0, 4.0 * n);
  LaunchCfg lc(dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st);
  if (dtype == kBF16) cudaLaunchKernelEx(&lc.cfg, gelu_fwd_kernel<__nv_bfloat16>, reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<__nv_bfloat16*>(h), n / 8, act);
  else cudaLaunchKernelEx(&lc.cfg, gelu_fwd_kernel<__half>, reinterpret_cast<const __half*>(a), reinterpret_cast<__half*>(h), n / 8, act);
  return done("gelu_forward");
}

int gate_backward(cudaStream_t st, int dtype, const void* dx, const void* y, const float* gate, void* dy, long n, float* dgate) {
  if (n <= 0) return 0;
  if (!act_dtype_ok(dtype, "gate_backward")) return -1;
  if (n % 8) { set_error("gate_backward: element count must be a multiple of 8"); return -1; }
  long blocks = (n / 8 + 255) / 256;
  if (blocks > 8L * num_sms()) blocks = 8L * num_sms();
  ProfScope ps(st, kProfOther, 0.0, 6.0 * n);
  LaunchCfg lc(dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st);
  if (dtype == kBF16) cudaLaunchKernelEx(&lc.cfg, gate_bwd_kernel<__nv_bfloat16>, reinterpret_cast<const __nv_bfloat16*>(dx), reinterpret_cast<const __nv_bfloat16*>(y), gate, reinterpret_cast<__nv_bfloat16*>(dy), n / 8, dgate);
  else cudaLaunchKernelEx(&lc.cfg, gate_bwd_kernel<__half>, reinterpret_cast<const __half*>(dx), reinterpret_cast<const __half*>(y), gate, reinterpret_cast<__half*>(dy), n / 8, dgate);
  return done("gate_backward");
}

template <typename T>
static int launch_wfold(cudaStream_t st, const void* G, const float* Gf, int splits, int ldg, const void* Wp, int ldw, const float* gamma,
                        const float* beta, const float* db, void* dW, int out_dtype, int ldo, int O, int I, float* dgamma, float* dbeta) {
  const int parts = (O + 15) / 16;
  LaunchCfg lc(dim3(static_cast<unsigned>((I + 255) / 256), static_cast<unsigned>(parts)), dim3(256), 0, st);
#define SF_ARGS(TO) reinterpret_cast<const T*>(G), Gf, splits, static_cast<long>(ldg), reinterpret_cast<const T*>(Wp), static_cast<long>(ldw), gamma, beta, db, \
                    reinterpret_cast<TO*>(dW), static_cast<long>(ldo), O, I, dgamma, dbeta
  if (out_dtype == kF32) cudaLaunchKernelEx(&lc.cfg, wfold_finish_kernel<T, float>, SF_ARGS(float));
  else if (out_dtype == kBF16) cudaLaunchKernelEx(&lc.cfg, wfold_finish_kernel<T, __nv_bfloat16>, SF_ARGS(__nv_bfloat16));
  else cudaLaunchKernelEx(&lc.cfg, wfold_finish_kernel<T, __half>, SF_ARGS(__half));
#undef SF_ARGS
  return done("wfold_finish");
}
int wfold_finish(cudaStream_t st, int dtype, const void* G, int ldg, const void* Wp, int ldw, const float* gamma, const float* beta,
                 const float* db, void* dW, int out_dtype, int ldo, int O, int I, float* dgamma, float* dbeta, const float* Gf, int splits) {
  if (O <= 0 || I <= 0) return 0;
  if (!G && !Gf) { set_error("wfold_finish: no gradient matrix"); return -1; }
  if (!act_dtype_ok(dtype, "wfold_finish")) return -1;
  if (gamma && !Wp) { set_error("wfold_finish: the packed matrix is needed for the LayerNorm parameter gradients"); return -1; }
  ProfScope ps(st, kProfOther, 0.0, 6.0 * O * I);
  return dtype == kBF16 ? launch_wfold<__nv_bfloat16>(st, G, Gf, splits, ldg, Wp, ldw, gamma, beta, db, dW, out_dtype, ldo, O, I, dgamma, dbeta)
                        : launch_wfold<__half>(st, G, Gf, splits, ldg, Wp, ldw, gamma, beta, db, dW, out_dtype, ldo, O, I, dgamma, dbeta);
}

int embed_table_grad(cudaStream_t st, int dtype, const void* dx, int ld, int B, int Tn, int Sn, int D, int mode, const int* tidx, float* out) {
  if (!act_dtype_ok(dtype, "embed_table_grad")) return -1;
  if ((D % 8) || D > 768 || (ld % 8)) { set_error("embed_table_grad: D must be a multiple of 8 up to 768"); return -1; }
  if (mode == 1 && !tidx) { set_error("embed_table_grad: the time table needs its frame -> row map"); return -1; }
  const int keys = mode == 0 ? Sn : Tn;
  const int span = mode == 0 ? B * Tn : B * Sn;
  int parts = span < 16 ? span : 16;
  if (parts < 1) parts = 1;
  ProfScope ps(st, kProfOther, 0.0, 2.0 * B * Tn * Sn * D);
  LaunchCfg lc(dim3(static_cast<unsigned>(keys), static_cast<unsigned>(parts)), dim3(96), 0, st);
  if (dtype == kBF16) cudaLaunchKernelEx(&lc.cfg, embed_table_grad_kernel<__nv_bfloat16>, reinterpret_cast<const __nv_bfloat16*>(dx), static_cast<long>(ld), B, Tn, Sn, D, mode, tidx, out);
  else cudaLaunchKernelEx(&lc.cfg, embed_table_grad_kernel<__half>, reinterpret_cast<const __half*>(dx), static_cast<long>(ld), B, Tn, Sn, D, mode, tidx, out);
  return done("embed_table_grad");
}

int rowperm(cudaStream_t st, const void* in, void* out, long M, int row_bytes, int row_map, int Tn, int Sn) {
  if (M <= 0) return 0;
  if (row_bytes % 16) { set_error("rowperm: rows must be multiples of 16 bytes"); return -1; }
  const int chunks = row_bytes / 16;
  long blocks = (M * chunks + 255) / 256;
  if (blocks > 16L * num_sms()) blocks = 16L * num_sms();
  ProfScope ps(st, kProfOther, 0.0, 2.0 * M * row_bytes);
  LaunchCfg lc(dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st);
  cudaLaunchKernelEx(&lc.cfg, rowperm_kernel, reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), M, chunks, row_map, Tn > 0 ? Tn : 1, Sn > 0 ? Sn : 1);
  return done("rowperm");
}

}  // namespace sf
