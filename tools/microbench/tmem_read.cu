// TMEM read-bandwidth microbenchmark: W warps of one CTA per SM loop over tcgen05.ld 32x32b.x32 / .x16
// of their lane quarter; reports bytes per SM clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I streamformer_b200/csrc tools/microbench/tmem_read.cu -o gpurun_out/tmem_read
#include <cstdio>
#include <cuda_runtime.h>
#include "sf_ptx.cuh"
using namespace sf;

template <int X>
__global__ void __launch_bounds__(512, 1) k(int iters, int warps, long long* out, unsigned* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < warps) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < 512; c += 64) {
        if constexpr (X == 32) {
          uint32_t v[32], w[32];
          tmem_ld_32x32b_x32(base + c, v);
          tmem_ld_32x32b_x32(base + c + 32, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc ^= v[j] + w[j];
        } else {
          uint32_t v[16], w[16], x[16], y[16];
          tmem_ld_32x32b_x16(base + c, v);
          tmem_ld_32x32b_x16(base + c + 16, w);
          tmem_ld_32x32b_x16(base + c + 32, x);
          tmem_ld_32x32b_x16(base + c + 48, y);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc ^= v[j] + w[j] + x[j] + y[j];
        }
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* out; unsigned* sink;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 200;
  for (int x : {32, 16}) {
    for (int warps : {4, 8, 16}) {
      if (x == 32) k<32><<<148, 512>>>(iters, warps, out, sink); else k<16><<<148, 512>>>(iters, warps, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = double(iters) * warps * 512.0 * 32 * 4;   // per SM
      printf("tcgen05.ld.32x32b.x%d  warps=%2d : %lld clk, %.1f B/clk/SM (%s)\n", x, warps, mx, bytes / mx, cudaGetErrorString(e));
    }
  }
  return 0;
}
