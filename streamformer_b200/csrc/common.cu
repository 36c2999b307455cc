// common.cu — error string, launch counter and small element-wise helpers shared by the library.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <vector>

#include "sf_kernels.h"
#include "sf_ptx.cuh"
#include "sf_tma.h"

namespace sf {

namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

const char* last_error() { return g_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
uint64_t launch_count() { return g_launches.load(); }

namespace {
struct ProfRec { cudaEvent_t a, b; int cls; double flops, bytes; };
bool g_prof_on = false;
std::vector<ProfRec> g_recs;
std::vector<cudaEvent_t> g_event_pool;
cudaEvent_t take_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace
// mode 0 off, 1 per-kernel events, 2 per-phase events only (kernels run back to back, PDL intact)
namespace {
int g_prof_mode = 0;
struct PhaseRec { cudaEvent_t a, b; int phase; };
std::vector<PhaseRec> g_phase_recs;
}  // namespace
void prof_set_mode(int mode) { g_prof_mode = mode; g_prof_on = (mode == 1); }
void prof_enable(bool on) { prof_set_mode(on ? 1 : 0); }
bool prof_enabled() { return g_prof_on; }
bool phase_prof_enabled() { return g_prof_mode == 2; }
void phase_begin(cudaStream_t st, int phase) {
  PhaseRec r; r.a = take_event(); r.b = take_event(); r.phase = phase;
  cudaEventRecord(r.a, st);
  g_phase_recs.push_back(r);
}
void phase_end(cudaStream_t st) { if (!g_phase_recs.empty()) cudaEventRecord(g_phase_recs.back().b, st); }
int phase_collect(double* ms, long long* count, int n) {
  for (int i = 0; i < n; ++i) { ms[i] = 0; count[i] = 0; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { set_error("phase_collect: %s", cudaGetErrorString(e)); return -2; }
  for (auto& r : g_phase_recs) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    if (r.phase >= 0 && r.phase < n) { ms[r.phase] += t; count[r.phase] += 1; }
    g_event_pool.push_back(r.a); g_event_pool.push_back(r.b);
  }
  g_phase_recs.clear();
  return 0;
}
void prof_begin(cudaStream_t st, int cls, double flops, double bytes) {
  ProfRec r; r.a = take_event(); r.b = take_event(); r.cls = cls; r.flops = flops; r.bytes = bytes;
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
}
void prof_end(cudaStream_t st) { if (!g_recs.empty()) cudaEventRecord(g_recs.back().b, st); }
int prof_collect(double* ms, double* flops, double* bytes, long long* launches, int n) {
  for (int i = 0; i < n; ++i) { ms[i] = 0; flops[i] = 0; bytes[i] = 0; launches[i] = 0; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { set_error("prof_collect: %s", cudaGetErrorString(e)); return -2; }
  for (auto& r : g_recs) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    if (r.cls >= 0 && r.cls < n) { ms[r.cls] += t; flops[r.cls] += r.flops; bytes[r.cls] += r.bytes; launches[r.cls] += 1; }
    g_event_pool.push_back(r.a); g_event_pool.push_back(r.b);
  }
  g_recs.clear();
  return 0;
}
void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n)); }

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

LaunchCfg::LaunchCfg(dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster) {
  static const bool pdl = [] { const char* e = getenv("SF_PDL"); return !(e && e[0] == '0'); }();
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  int n = 0;
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = static_cast<unsigned>(n);
}

namespace {

template <typename S>
__device__ __forceinline__ float to_f32(S v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename D>
__device__ __forceinline__ D from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, size_t n) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) dst[i] = from_f32<D>(to_f32<S>(src[i]));
}

template <typename S>
int cast_dispatch_dst(cudaStream_t st, const void* src, int dd, void* dst, size_t n) {
  const int threads = 256;
  size_t want = (n + threads - 1) / threads;
  int blocks = static_cast<int>(want < 148 * 16 ? (want ? want : 1) : 148 * 16);
  const S* s = reinterpret_cast<const S*>(src);
  switch (dd) {
    case kBF16: cast_kernel<S, __nv_bfloat16><<<blocks, threads, 0, st>>>(s, reinterpret_cast<__nv_bfloat16*>(dst), n); break;
    case kF16: cast_kernel<S, __half><<<blocks, threads, 0, st>>>(s, reinterpret_cast<__half*>(dst), n); break;
    case kF32: cast_kernel<S, float><<<blocks, threads, 0, st>>>(s, reinterpret_cast<float*>(dst), n); break;
    default: set_error("cast: bad dst dtype %d", dd); return -1;
  }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("cast launch: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}

}  // namespace

int cast(cudaStream_t stream, int sd, const void* src, int dd, void* dst, size_t n) {
  if (n == 0) return 0;
  switch (sd) {
    case kBF16: return cast_dispatch_dst<__nv_bfloat16>(stream, src, dd, dst, n);
    case kF16: return cast_dispatch_dst<__half>(stream, src, dd, dst, n);
    case kF32: return cast_dispatch_dst<float>(stream, src, dd, dst, n);
    default: set_error("cast: bad src dtype %d", sd); return -1;
  }
}

}  // namespace sf
