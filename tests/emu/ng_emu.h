// TEST INFRASTRUCTURE ONLY — never shipped, never loaded by the nanograd_b200 package.
//
// Host emulation of the small slice of CUDA that nanograd_b200/csrc uses, so the kernel
// *sources* can be compiled with g++ (-DNG_EMU) and executed in the GPU-less build container:
// every CUDA thread of a block runs as its own execution context (a user-level fiber by default, a host thread with
// NG_EMU_THREADS=1), __syncthreads() is a real barrier and warp shuffles exchange values between the 32 contexts of a
// warp; launches are checked against a real GPU's limits (grid dimensions, dynamic shared memory).  It exists to catch
// indexing / marshalling bugs before GPU time is spent; it is neither a fallback nor a product
// path (the shipped library is built by nvcc for sm_100a and ng_init fails without a GPU).
#pragma once

#include <algorithm>
#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __grid_constant__
#define __shared__ static
#define __align__(n) alignas(n)

struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
  dim3(int a) : x((unsigned)a), y(1), z(1) {}
  dim3(long a) : x((unsigned)a), y(1), z(1) {}
  dim3(long long a) : x((unsigned)a), y(1), z(1) {}
};

extern thread_local uint3 threadIdx;
extern thread_local uint3 blockIdx;
extern thread_local dim3 blockDim;
extern thread_local dim3 gridDim;

// ---- minimal runtime API ---------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef struct ng_emu_stream* cudaStream_t;
typedef struct ng_emu_event* cudaEvent_t;
typedef struct ng_emu_graph* cudaGraph_t;       // graphs are not emulated: the types exist so runtime.cu compiles
typedef struct ng_emu_graph_exec* cudaGraphExec_t;
struct ng_emu_event { std::chrono::steady_clock::time_point t; };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaEventDisableTiming = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
struct cudaDeviceProp { int multiProcessorCount, major, minor; size_t totalGlobalMem; };

inline const char* cudaGetErrorString(cudaError_t) { return "emulated error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  p->multiProcessorCount = 148; p->major = 10; p->minor = 0; p->totalGlobalMem = 8ull << 30; return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return cudaSuccess; }
// NG_EMU_POISON_NAN=1 (tools/kernel_initcheck.py): every block the caching allocator hands out — fresh or recycled —
// is filled with NaNs, so a kernel that consumes device memory nobody wrote turns its result into NaN.
inline void ng_emu_poison_nan(void* p, size_t n) {
  static const bool on = [] { const char* e = getenv("NG_EMU_POISON_NAN"); return e && e[0] == '1'; }();
  if (on && p) memset(p, 0xFF, n);
}
inline cudaError_t cudaMalloc(void** p, size_t n) {
#ifdef NG_EMU_EXACT_ALLOC
  // AddressSanitizer build (tools/kernel_memcheck.py): exactly n bytes, so that one element past the end is reported
  if (posix_memalign(p, 256, n ? n : 1) != 0) *p = nullptr;
  if (*p) memset(*p, 0xCD, n);
#else
  // 256-byte aligned like cudaMalloc, plus a poisoned tail to expose overruns in tests
  *p = aligned_alloc(256, (n + 255) / 256 * 256);
  if (*p) memset(*p, 0xCD, (n + 255) / 256 * 256);
#endif
  return *p ? cudaSuccess : 2;
}
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaHostAlloc(void** p, size_t n, int) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new ng_emu_event(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, int) { *e = new ng_emu_event(); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }

// ---- device intrinsics -----------------------------------------------------------------------
namespace ng_emu {
void sync_block();
float shfl(float v, int src_lane);
void* dyn_smem_ptr();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}  // namespace ng_emu

inline void __syncthreads() { ng_emu::sync_block(); }
inline float __shfl_xor_sync(unsigned, float v, int mask) {
  const int lane = (int)(threadIdx.x & 31);
  return ng_emu::shfl(v, lane ^ mask);
}
inline float __shfl_down_sync(unsigned, float v, int delta) {
  const int lane = (int)(threadIdx.x & 31);
  return ng_emu::shfl(v, lane + delta < 32 ? lane + delta : lane);
}
inline float __shfl_sync(unsigned, float v, int src) { return ng_emu::shfl(v, src & 31); }
inline float __ldg(const float* p) { return *p; }
inline float __ldcg(const float* p) { return *(const volatile float*)p; }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned int __float_as_uint(float v) { unsigned int u; memcpy(&u, &v, 4); return u; }
inline float __uint_as_float(unsigned int u) { float v; memcpy(&v, &u, 4); return v; }
inline unsigned int atomicMax(unsigned int* p, unsigned int v) {   // CUDA threads are host threads here
  unsigned int old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
inline float rsqrtf(float x) { return 1.f / sqrtf(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float x) { return expf(x); }
inline float __logf(float x) { return logf(x); }

#ifndef INFINITY
#define INFINITY (__builtin_inff())
#endif
