// wgrad_tcgen05.cu — weight gradients G[O, I] = dY^T . X of every nn.Linear on the path (the wgrad half
// of what torch.autograd computes for the reference's training step, tools/finetune_tools.py:543-573).
//
// dY [M, O] and X [M, I] are the row-major activations the backward already holds; the contraction runs over the
// token rows M (25 088 .. 100 352), so BOTH operands are "MN-major" for the tensor core: TMA drops 64-token x
// 64-column boxes (128-byte rows, SWIZZLE_128B) into shared memory exactly as they lie in HBM and tcgen05.mma reads
// them through MN-major descriptors (a_major = b_major = MN in the instruction descriptor; leading byte offset = one
// box, stride byte offset = one 8-row group) — no transposed copies of the activations are ever made.
//
//   one CTA  = one 128 (O) x 256 (I) tile of G  x  one slice of the token rows (split-K: G is only 0.6 .. 2.4 M
//              elements, 18 .. 72 tiles, so the row range is split until ~148 CTAs are busy)
//   warp 0     TMA producer  (6 boxes = 48 KB per 64-token block, 4-stage mbarrier ring)
//   warp 1     MMA issuer    (tcgen05.mma kind::f16, M = 128, N = 256, K = 16; fp32 accumulator = 256 TMEM columns)
//   warps 2-5  epilogue      (tcgen05.ld 32x32b, thread == row of G, fp32 partial tile -> global memory)
//
// The fp32 partials [splits][O][I] are summed by wfold_finish (backward.cu), which also applies the LayerNorm fold.
#include <cuda.h>
#include <stdint.h>

#include <type_traits>

#include "sf_kernels.h"
#include "sf_ptx.cuh"
#include "sf_tma.h"

namespace sf {
namespace {

constexpr int kWM = 128, kWN = 256, kWK = 64;       // tile of G and token rows per pipeline stage
constexpr int kBoxBytes = 64 * 128;                 // one TMA box: 64 token rows x 64 columns x 2 B
constexpr int kWStageBytes = (kWM / 64 + kWN / 64) * kBoxBytes;   // 48 KB
constexpr int kWStages = 4;
constexpr int kWSmem = kWStages * kWStageBytes + 256 + 1024;

struct WgradParams {
  int M, O, I;
  int k_blocks;          // 64-row blocks over M
  int kb_per_split;
  int splits;
  float* out;            // [splits][O][I]
};

// MN-major, SWIZZLE_128B: k-rows of 128 bytes, 8-row groups 1024 B apart (SBO), 64-element atoms along M/N one box apart (LBO)
__device__ __forceinline__ uint64_t wdesc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(kBoxBytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <typename T>
__global__ void __launch_bounds__(192, 1)
wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
  extern __shared__ uint8_t wsm_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wsm_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kWStages * kWStageBytes);
  uint64_t* empty_bar = full_bar + kWStages;
  uint64_t* tmem_full = empty_bar + kWStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i_tiles = (p.I + kWN - 1) / kWN, o_tiles = (p.O + kWM - 1) / kWM;
  const int tile = blockIdx.x % (i_tiles * o_tiles), split = blockIdx.x / (i_tiles * o_tiles);
  const int o0 = (tile / i_tiles) * kWM, i0 = (tile % i_tiles) * kWN;
  const int kb0 = split * p.kb_per_split;
  int kb1 = kb0 + p.kb_per_split;
  if (kb1 > p.k_blocks) kb1 = p.k_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < kWStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, kWN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch_dependents();

  if (warp == 0) {
    if (elect_one_sync()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kWStageBytes;
        uint8_t* sb = sa + (kWM / 64) * kBoxBytes;
        mbar_arrive_expect_tx(&full_bar[stage], kWStageBytes);
#pragma unroll
        for (int a = 0; a < kWM / 64; ++a) tma_load_2d(sa + a * kBoxBytes, &tmY, &full_bar[stage], o0 + 64 * a, kb * kWK);
#pragma unroll
        for (int b = 0; b < kWN / 64; ++b) tma_load_2d(sb + b * kBoxBytes, &tmX, &full_bar[stage], i0 + 64 * b, kb * kWK);
        if (++stage == kWStages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr int fmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
      constexpr uint32_t idesc = umma_idesc_f16(kWM, kWN, fmt, 1) | (1u << 15);     // a_major = b_major = MN
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kWStageBytes);
        const uint32_t sb = sa + (kWM / 64) * kBoxBytes;
#pragma unroll
        for (int k = 0; k < kWK / 16; ++k)       // 16 token rows = two 8-row groups = 2048 B further into every box
          umma_f16(tmem_base, wdesc(sa + k * 2048), wdesc(sb + k * 2048), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == kWStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int o = o0 + q * 32 + lane;
    float* dst = p.out + (static_cast<size_t>(split) * p.O + o) * p.I + i0;
    if (kb1 > kb0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < kWN; c += 32) {
      uint32_t v[32];
      if (kb1 > kb0) {
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;  // empty slice (more splits than 64-row blocks): contributes zeros
      }
      if (o < p.O) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (i0 + c + j < p.I)                  // I % 8 == 0: whole float4 groups
            *reinterpret_cast<float4*>(dst + c + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                  __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kWN); }
}

int make_act_map(CUtensorMap* map, int dtype, const void* base, int rows, int cols, int ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -3;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64u, 64u};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(wgrad operand) failed (%d): rows=%d cols=%d ld=%d", (int)r, rows, cols, ld); return -3; }
  return 0;
}

}  // namespace

// Number of row slices wgrad() uses for an [O, I] gradient over M token rows (the caller sizes the partial buffer with it).
int wgrad_splits(int M, int O, int I) {
  const int tiles = ((O + kWM - 1) / kWM) * ((I + kWN - 1) / kWN);
  const int k_blocks = (M + kWK - 1) / kWK;
  int splits = num_sms() / (tiles > 0 ? tiles : 1);
  if (splits < 1) splits = 1;
  if (splits > 16) splits = 16;
  if (splits > k_blocks) splits = k_blocks > 0 ? k_blocks : 1;
  return splits;
}

int wgrad(cudaStream_t stream, int dtype, const void* dY, int ldy, const void* X, int ldx, int M, int O, int I, float* partials) {
  if (O <= 0 || I <= 0) return 0;
  if (dtype != kBF16 && dtype != kF16) { set_error("wgrad: dtype must be bf16 or f16"); return -1; }
  if ((O % 8) || (I % 8) || (ldy % 8) || (ldx % 8) || M <= 0) { set_error("wgrad: O, I and the leading dims must be multiples of 8 (M=%d O=%d I=%d)", M, O, I); return -1; }
  CUtensorMap tmY, tmX;
  int rc = make_act_map(&tmY, dtype, dY, M, O, ldy);
  if (rc) return rc;
  rc = make_act_map(&tmX, dtype, X, M, I, ldx);
  if (rc) return rc;
  WgradParams p;
  p.M = M; p.O = O; p.I = I;
  p.k_blocks = (M + kWK - 1) / kWK;
  p.splits = wgrad_splits(M, O, I);
  p.kb_per_split = (p.k_blocks + p.splits - 1) / p.splits;
  p.out = partials;
  const int tiles = ((O + kWM - 1) / kWM) * ((I + kWN - 1) / kWN);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(wgrad_tcgen05_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem);
    cudaFuncSetAttribute(wgrad_tcgen05_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem);
    attr_set = true;
  }
  cudaError_t e;
  {
    ProfScope ps(stream, kProfGemm, 2.0 * M * static_cast<double>(O) * I, 2.0 * M * (static_cast<double>(O) + I) + 4.0 * p.splits * O * I);
    LaunchCfg lc(dim3(static_cast<unsigned>(tiles * p.splits)), dim3(192), kWSmem, stream);
    if (dtype == kBF16) e = cudaLaunchKernelEx(&lc.cfg, wgrad_tcgen05_kernel<__nv_bfloat16>, tmY, tmX, p);
    else e = cudaLaunchKernelEx(&lc.cfg, wgrad_tcgen05_kernel<__half>, tmY, tmX, p);
  }
  count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("wgrad launch failed: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}

}  // namespace sf
