// rowwise.cu — the HBM-bound row kernels of the encoder:
//   * LayerNorm (reference: nn.LayerNorm at models/modeling_timesformer_siglip.py:860-880, 943, 974,
//     997, 1251, 1330, 1138) — one warp per token row, fp32 statistics, 16-byte loads/stores,
//     optional (b,n,t)<->(b,t,n) row permutation on the way out (replaces the reference's
//     materialised permute copies at :962-971, 982-991, 1332-1346).
//   * im2col for the 16x16/s16 patch-embedding conv (reference :336-350): pixels -> patch-major
//     GEMM operand with K ordered (c, kh, kw), cast to the activation dtype on the fly.
#include <stdlib.h>

#include "sf_kernels.h"
#include "sf_ptx.cuh"

namespace sf {
namespace {

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = Pack2<T>::unpack(u.x), b = Pack2<T>::unpack(u.y);
  const float2 c = Pack2<T>::unpack(u.z), d = Pack2<T>::unpack(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

__device__ __forceinline__ long map_row(long m, int row_map, int Tn, int Sn) {
  if (row_map == kRowBNTtoBTN) {
    const long t = m % Tn, bn = m / Tn;
    const long n = bn % Sn, b = bn / Sn;
    return (b * Tn + t) * Sn + n;
  }
  if (row_map == kRowBTNtoBNT) {
    const long n = m % Sn, bt = m / Sn;
    const long t = bt % Tn, b = bt / Tn;
    return (b * Sn + n) * Tn + t;
  }
  return m;
}

template <typename T>
__device__ __forceinline__ uint4 normalize8(const float (&v)[8], float mean, float rstd, const float* gamma,
                                            const float* beta) {
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma));
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + 4));
  uint4 o;
  o.x = Pack2<T>::pack((v[0] - mean) * rstd * g0.x + b0.x, (v[1] - mean) * rstd * g0.y + b0.y);
  o.y = Pack2<T>::pack((v[2] - mean) * rstd * g0.z + b0.z, (v[3] - mean) * rstd * g0.w + b0.w);
  o.z = Pack2<T>::pack((v[4] - mean) * rstd * g1.x + b1.x, (v[5] - mean) * rstd * g1.y + b1.y);
  o.w = Pack2<T>::pack((v[6] - mean) * rstd * g1.z + b1.z, (v[7] - mean) * rstd * g1.w + b1.w);
  return o;
}

// One warp per row, the whole row cached in registers: NCH 16-byte chunks per lane (D <= NCH*256).
template <typename T, int NCH>
__global__ void __launch_bounds__(256, 5)
layernorm_kernel(const T* __restrict__ x, int ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, T* __restrict__ y, int ldy, int M, int D,
                 int row_map, int Tn, int Sn) {
  griddep_wait();
  griddep_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int nchunks = D >> 3;
  const long m = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const T* xr = x + m * ldx;
  float v[NCH][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      unpack8<T>(*reinterpret_cast<const uint4*>(xr + c * 8), v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
  }
  const float mean = warp_sum(sum) / static_cast<float>(D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    if (lane + 32 * i < nchunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(D) + eps);
  T* yr = y + map_row(m, row_map, Tn, Sn) * ldy;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks)
      *reinterpret_cast<uint4*>(yr + c * 8) = normalize8<T>(v[i], mean, rstd, gamma + c * 8, beta + c * 8);
  }
}

// Any D (multiple of 8): three passes over the row through L1/L2.
template <typename T>
__global__ void __launch_bounds__(256)
layernorm_wide_kernel(const T* __restrict__ x, int ldx, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, T* __restrict__ y, int ldy, int M, int D,
                      int row_map, int Tn, int Sn) {
  griddep_wait();
  griddep_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int nchunks = D >> 3;
  const long m = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const T* xr = x + m * ldx;
  float f[8];
  float sum = 0.f;
  for (int c = lane; c < nchunks; c += 32) {
    unpack8<T>(*reinterpret_cast<const uint4*>(xr + c * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += f[j];
  }
  const float mean = warp_sum(sum) / static_cast<float>(D);
  float sq = 0.f;
  for (int c = lane; c < nchunks; c += 32) {
    unpack8<T>(*reinterpret_cast<const uint4*>(xr + c * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) sq += (f[j] - mean) * (f[j] - mean);
  }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(D) + eps);
  T* yr = y + map_row(m, row_map, Tn, Sn) * ldy;
  for (int c = lane; c < nchunks; c += 32) {
    unpack8<T>(*reinterpret_cast<const uint4*>(xr + c * 8), f);
    *reinterpret_cast<uint4*>(yr + c * 8) = normalize8<T>(f, mean, rstd, gamma + c * 8, beta + c * 8);
  }
}

template <typename T>
void launch_layernorm(cudaStream_t stream, const void* x, int ldx, const float* gamma, const float* beta, float eps,
                      void* y, int ldy, int M, int D, int row_map, int Tn, int Sn) {
  const int threads = 256, wpb = threads / 32;
  const unsigned blocks = static_cast<unsigned>((static_cast<long>(M) + wpb - 1) / wpb);
  const T* xi = reinterpret_cast<const T*>(x);
  T* yo = reinterpret_cast<T*>(y);
  LaunchCfg lc(dim3(blocks), dim3(threads), 0, stream);
#define SF_LN_ARGS xi, ldx, gamma, beta, eps, yo, ldy, M, D, row_map, Tn, Sn
  if (D <= 256) cudaLaunchKernelEx(&lc.cfg, layernorm_kernel<T, 1>, SF_LN_ARGS);
  else if (D <= 512) cudaLaunchKernelEx(&lc.cfg, layernorm_kernel<T, 2>, SF_LN_ARGS);
  else if (D <= 768) cudaLaunchKernelEx(&lc.cfg, layernorm_kernel<T, 3>, SF_LN_ARGS);
  else if (D <= 1024) cudaLaunchKernelEx(&lc.cfg, layernorm_kernel<T, 4>, SF_LN_ARGS);
  else cudaLaunchKernelEx(&lc.cfg, layernorm_wide_kernel<T>, SF_LN_ARGS);
#undef SF_LN_ARGS
}

// (sum, sum of squares) of every row, one warp per row: the one-partial statistics table a folded
// LayerNorm consumes when its input was not produced by gemm() (GemmEpilogue::ln_stats).
template <typename T>
__global__ void __launch_bounds__(256)
rowstats_kernel(const T* __restrict__ x, int ldx, int M, int D, float2* __restrict__ stats) {
  griddep_wait();
  griddep_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long m = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const T* xr = x + m * ldx;
  float s1 = 0.f, s2 = 0.f, f[8];
  for (int c = lane; c < (D >> 3); c += 32) {
    unpack8<T>(*reinterpret_cast<const uint4*>(xr + c * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1 += f[j];
      s2 = fmaf(f[j], f[j], s2);
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) stats[m] = make_float2(s1, s2);
}

template <typename PixT>
__device__ __forceinline__ void load8(const PixT* p, float (&f)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = Pack2<__nv_bfloat16>::unpack(w[j]);
    f[2 * j] = t.x; f[2 * j + 1] = t.y;
  }
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = Pack2<__half>::unpack(w[j]);
    f[2 * j] = t.x; f[2 * j + 1] = t.y;
  }
}

template <>
__device__ __forceinline__ void load8<uint8_t>(const uint8_t* p, float (&f)[8]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[j] = static_cast<float>((u.x >> (8 * j)) & 0xffu);
    f[4 + j] = static_cast<float>((u.y >> (8 * j)) & 0xffu);
  }
}

// uint8 frames are normalised on the way in: (x / 255 - mean[c]) / std[c], in fp32 and in exactly this
// order — ClipToTensor (x / 255) then Normalize (sub mean, div std) of the reference's loaders
// (extract_oad_feature.py:42-48, datasets/kinetics_sparse.py:110-118) — so the result is bit-identical
// to handing the loader's fp32 tensor to the float path.  IEEE division, no reciprocal shortcuts.
struct PixNorm {
  float mean[4], std[4];
};
template <typename PixT>
__device__ __forceinline__ void normalise8(float (&f)[8], const PixNorm& nm, int c) {
  if constexpr (sizeof(PixT) == 1) {
    const float m = nm.mean[c], sd = nm.std[c];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __fdiv_rn(__fsub_rn(__fdiv_rn(f[j], 255.0f), m), sd);
  }
}

// one thread = 8 consecutive kw of one (patch row m, channel c, kernel row kh)
template <typename PixT, typename T>
__global__ void __launch_bounds__(256)
im2col_kernel(const PixT* __restrict__ pix, T* __restrict__ out, int BT, int C, int H, int W, int P, PixNorm nm) {
  griddep_wait();
  griddep_launch_dependents();
  const int gw = W / P, gh = H / P;
  const int S = gw * gh;
  const int K = C * P * P;
  const int chunks_per_row = K >> 3;
  const long total = static_cast<long>(BT) * S * chunks_per_row;
  const int pc = P >> 3;  // chunks per kernel row
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % chunks_per_row);
    const long m = i / chunks_per_row;
    const int n = static_cast<int>(m % S);
    const long bt = m / S;
    const int ph = n / gw, pw = n % gw;
    const int part = j % pc;
    const int kh = (j / pc) % P;
    const int c = j / (pc * P);
    const PixT* src = pix + ((bt * C + c) * H + (ph * P + kh)) * static_cast<long>(W) + pw * P + part * 8;
    float f[8];
    load8<PixT>(src, f);
    normalise8<PixT>(f, nm, c);
    uint4 o;
    o.x = Pack2<T>::pack(f[0], f[1]);
    o.y = Pack2<T>::pack(f[2], f[3]);
    o.z = Pack2<T>::pack(f[4], f[5]);
    o.w = Pack2<T>::pack(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + m * K + j * 8) = o;
  }
}

// Tiled variant: one CTA per (frame, patch row).  The P image rows x C channels of the strip are read
// as whole pixel rows (fully coalesced, converted to T on the way into shared memory), then every
// patch's K = C*P*P values leave as one contiguous output row — both sides of the copy move whole
// cache lines (the gather kernel above reads 32-byte pieces 16 rows apart: 2.1 TB/s).
template <typename PixT, typename T>
__global__ void __launch_bounds__(256)
im2col_strip_kernel(const PixT* __restrict__ pix, T* __restrict__ out, int C, int H, int W, int P, PixNorm nm) {
  extern __shared__ __align__(16) uint8_t strip_raw[];
  T* strip = reinterpret_cast<T*>(strip_raw);                 // [C][P][W]
  griddep_wait();
  griddep_launch_dependents();
  const int gw = W / P, gh = H / P;
  const long bt = blockIdx.x / gh;
  const int ph = blockIdx.x % gh;
  const int wc = W >> 3;                                      // 8-pixel pieces per image row
  const int pieces = C * P * wc;
  for (int i = threadIdx.x; i < pieces; i += blockDim.x) {
    const int x8 = i % wc, rowi = i / wc;                     // rowi = c * P + kh
    const int c = rowi / P, kh = rowi % P;
    const PixT* src = pix + ((bt * C + c) * H + (ph * P + kh)) * static_cast<long>(W) + x8 * 8;
    float f[8];
    load8<PixT>(src, f);
    normalise8<PixT>(f, nm, c);
    uint4 o;
    o.x = Pack2<T>::pack(f[0], f[1]);
    o.y = Pack2<T>::pack(f[2], f[3]);
    o.z = Pack2<T>::pack(f[4], f[5]);
    o.w = Pack2<T>::pack(f[6], f[7]);
    *reinterpret_cast<uint4*>(strip + static_cast<long>(rowi) * W + x8 * 8) = o;
  }
  __syncthreads();
  const int K = C * P * P;
  const int kc = K >> 3;                                      // 8-element pieces per output row
  const int pc = P >> 3;
  const long m0 = (bt * gh + ph) * static_cast<long>(gw);
  for (int i = threadIdx.x; i < gw * kc; i += blockDim.x) {
    const int pw = i / kc, j = i % kc;
    const int part = j % pc, rowi = j / pc;                   // rowi = c * P + kh
    const uint4 v = *reinterpret_cast<const uint4*>(strip + static_cast<long>(rowi) * W + pw * P + part * 8);
    *reinterpret_cast<uint4*>(out + (m0 + pw) * K + j * 8) = v;
  }
}

// Interleaved uint8 frames [BT, H, W, C] (what a video decoder hands over): the P image rows of a patch
// row are ONE contiguous run of P*W*C bytes, read with 16-byte loads, normalised and scattered into the
// same [C][P][W] shared-memory strip; the write side is identical to the planar kernel.
template <typename T>
__global__ void __launch_bounds__(256)
im2col_strip_hwc_kernel(const uint8_t* __restrict__ pix, T* __restrict__ out, int C, int H, int W, int P, PixNorm nm) {
  extern __shared__ __align__(16) uint8_t strip_raw[];
  T* strip = reinterpret_cast<T*>(strip_raw);                 // [C][P][W]
  griddep_wait();
  griddep_launch_dependents();
  const int gw = W / P, gh = H / P;
  const long bt = blockIdx.x / gh;
  const int ph = blockIdx.x % gh;
  const int row_bytes = W * C;
  const int total = P * row_bytes;                            // multiple of 16: P % 8 == 0 and W % 8 == 0
  const uint8_t* src = pix + (bt * H + static_cast<long>(ph) * P) * row_bytes;
  for (int i = threadIdx.x * 16; i < total; i += blockDim.x * 16) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + i);
    const uint32_t wds[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int e = i + j;
      const int kh = e / row_bytes, r = e - kh * row_bytes;
      const int x = r / C, c = r - x * C;
      const float raw = static_cast<float>((wds[j >> 2] >> (8 * (j & 3))) & 0xffu);
      const float v = __fdiv_rn(__fsub_rn(__fdiv_rn(raw, 255.0f), nm.mean[c]), nm.std[c]);
      strip[(static_cast<long>(c) * P + kh) * W + x] = static_cast<T>(v);
    }
  }
  __syncthreads();
  const int K = C * P * P;
  const int kc = K >> 3;
  const int pc = P >> 3;
  const long m0 = (bt * gh + ph) * static_cast<long>(gw);
  for (int i = threadIdx.x; i < gw * kc; i += blockDim.x) {
    const int pw = i / kc, j = i % kc;
    const int part = j % pc, rowi = j / pc;
    const uint4 v = *reinterpret_cast<const uint4*>(strip + static_cast<long>(rowi) * W + pw * P + part * 8);
    *reinterpret_cast<uint4*>(out + (m0 + pw) * K + j * 8) = v;
  }
}

constexpr size_t kStripMaxBytes = 200 * 1024;   // opt-in dynamic shared memory (227 KB per CTA on sm_100)

template <typename T>
int launch_im2col_hwc(cudaStream_t st, const void* pix, void* out, int BT, int C, int H, int W, int P, const PixNorm& nm) {
  const size_t strip_bytes = static_cast<size_t>(C) * P * W * sizeof(T);
  if (strip_bytes > kStripMaxBytes || (W % 8) || (P % 8) || C > 4) {
    set_error("im2col: interleaved uint8 frames need C <= 4 and a %zu-byte strip <= %zu bytes (W=%d)", strip_bytes, kStripMaxBytes, W);
    return -1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(im2col_strip_hwc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kStripMaxBytes));
    attr_set = true;
  }
  const long total = static_cast<long>(BT) * (H / P) * (W / P) * (C * P * P / 8);
  {
    ProfScope ps(st, kProfIm2col, 0.0, static_cast<double>(total) * 8 * (1 + sizeof(T)));
    LaunchCfg lc(dim3(static_cast<unsigned>(BT * (H / P))), dim3(256), strip_bytes, st);
    cudaLaunchKernelEx(&lc.cfg, im2col_strip_hwc_kernel<T>, reinterpret_cast<const uint8_t*>(pix), reinterpret_cast<T*>(out), C, H, W, P, nm);
  }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("im2col launch: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}

template <typename PixT, typename T>
int launch_im2col(cudaStream_t st, const void* pix, void* out, int BT, int C, int H, int W, int P, const PixNorm& nm) {
  const size_t strip_bytes = static_cast<size_t>(C) * P * W * sizeof(T);
  static const bool strip_on = [] { const char* e = getenv("SF_IM2COL_STRIP"); return !(e && e[0] == '0'); }();
  if (strip_on && strip_bytes <= 48 * 1024 && (W % 8) == 0 && (P % 8) == 0) {
    const long total = static_cast<long>(BT) * (H / P) * (W / P) * (C * P * P / 8);
    {
      ProfScope ps(st, kProfIm2col, 0.0, static_cast<double>(total) * 8 * (sizeof(PixT) + sizeof(T)));
      LaunchCfg lc(dim3(static_cast<unsigned>(BT * (H / P))), dim3(256), strip_bytes, st);
      cudaLaunchKernelEx(&lc.cfg, im2col_strip_kernel<PixT, T>, reinterpret_cast<const PixT*>(pix), reinterpret_cast<T*>(out),
                         C, H, W, P, nm);
    }
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("im2col launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
  }
  const long total = static_cast<long>(BT) * (H / P) * (W / P) * (C * P * P / 8);
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 32) blocks = 148L * 32;
  if (blocks < 1) blocks = 1;
  {
    ProfScope ps(st, kProfIm2col, 0.0, static_cast<double>(total) * 8 * (sizeof(PixT) + sizeof(T)));
    LaunchCfg lc(dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st);
    cudaLaunchKernelEx(&lc.cfg, im2col_kernel<PixT, T>, reinterpret_cast<const PixT*>(pix), reinterpret_cast<T*>(out),
                       BT, C, H, W, P, nm);
  }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("im2col launch: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}

}  // namespace

int layernorm(cudaStream_t stream, int dtype, const void* x, int ldx, const float* gamma,
              const float* beta, float eps, void* y, int ldy, int M, int D, int row_map, int T, int S) {
  if (M <= 0) return 0;
  if ((D % 8) || (ldx % 8) || (ldy % 8)) {
    set_error("layernorm: D, ldx, ldy must be multiples of 8 (D=%d ldx=%d ldy=%d)", D, ldx, ldy);
    return -1;
  }
  ProfScope ps(stream, kProfLayerNorm, 0.0, 4.0 * static_cast<double>(M) * D);
  if (dtype == kBF16) {
    launch_layernorm<__nv_bfloat16>(stream, x, ldx, gamma, beta, eps, y, ldy, M, D, row_map, T, S);
  } else if (dtype == kF16) {
    launch_layernorm<__half>(stream, x, ldx, gamma, beta, eps, y, ldy, M, D, row_map, T, S);
  } else {
    set_error("layernorm: dtype must be bf16 or f16");
    return -1;
  }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("layernorm launch: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}

int rowstats(cudaStream_t stream, int dtype, const void* x, int ldx, int M, int D, float2* stats) {
  if (M <= 0) return 0;
  if ((D % 8) || (ldx % 8)) { set_error("rowstats: D and ldx must be multiples of 8"); return -1; }
  if (dtype != kBF16 && dtype != kF16) { set_error("rowstats: dtype must be bf16 or f16"); return -1; }
  const unsigned blocks = static_cast<unsigned>((static_cast<long>(M) + 7) / 8);
  {
    ProfScope ps(stream, kProfLayerNorm, 0.0, 2.0 * static_cast<double>(M) * D);
    LaunchCfg lc(dim3(blocks), dim3(256), 0, stream);
    if (dtype == kBF16) cudaLaunchKernelEx(&lc.cfg, rowstats_kernel<__nv_bfloat16>, reinterpret_cast<const __nv_bfloat16*>(x), ldx, M, D, stats);
    else cudaLaunchKernelEx(&lc.cfg, rowstats_kernel<__half>, reinterpret_cast<const __half*>(x), ldx, M, D, stats);
  }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("rowstats launch: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}

int im2col_patches(cudaStream_t stream, int pix_dtype, const void* pixels, int act_dtype, void* out,
                   int BT, int C, int H, int W, int P, const float* mean, const float* std) {
  if (BT <= 0) return 0;
  PixNorm nm;
  for (int i = 0; i < 4; ++i) { nm.mean[i] = mean ? mean[i] : 0.5f; nm.std[i] = std ? std[i] : 0.5f; }
  if ((pix_dtype == kU8 || pix_dtype == kU8HWC) && C > 4) { set_error("im2col: uint8 frames support at most 4 channels"); return -1; }
  if ((P % 8) || (H % P) || (W % P)) {
    set_error("im2col: patch size must be a multiple of 8 and divide H, W (H=%d W=%d P=%d)", H, W, P);
    return -1;
  }
#define SF_IM2COL(PT, AT) return launch_im2col<PT, AT>(stream, pixels, out, BT, C, H, W, P, nm)
  if (pix_dtype == kU8HWC) {
    if (act_dtype == kBF16) return launch_im2col_hwc<__nv_bfloat16>(stream, pixels, out, BT, C, H, W, P, nm);
    if (act_dtype == kF16) return launch_im2col_hwc<__half>(stream, pixels, out, BT, C, H, W, P, nm);
  }
  if (act_dtype == kBF16) {
    if (pix_dtype == kU8) SF_IM2COL(uint8_t, __nv_bfloat16);
    if (pix_dtype == kF32) SF_IM2COL(float, __nv_bfloat16);
    if (pix_dtype == kBF16) SF_IM2COL(__nv_bfloat16, __nv_bfloat16);
    if (pix_dtype == kF16) SF_IM2COL(__half, __nv_bfloat16);
  } else if (act_dtype == kF16) {
    if (pix_dtype == kU8) SF_IM2COL(uint8_t, __half);
    if (pix_dtype == kF32) SF_IM2COL(float, __half);
    if (pix_dtype == kBF16) SF_IM2COL(__nv_bfloat16, __half);
    if (pix_dtype == kF16) SF_IM2COL(__half, __half);
  }
#undef SF_IM2COL
  set_error("im2col: unsupported dtypes pix=%d act=%d", pix_dtype, act_dtype);
  return -1;
}

}  // namespace sf
