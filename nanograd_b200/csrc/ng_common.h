// Shared internals of the nanograd_b200 device library (not part of the C ABI).
#pragma once

#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>

#ifdef NG_EMU
// Test-only build: the kernels are compiled for the host and each CUDA thread is run as a
// host thread (tests/emu/ng_emu.h).  The shipped library never defines NG_EMU.
#include "ng_emu.h"
#else
#include <cuda_runtime.h>
#endif

#include "../../include/nanograd_b200.h"

namespace ng {

// ---- process-global state (runtime.cu) ---------------------------------------------------------
extern cudaStream_t g_stream;        // the single in-order compute stream
extern int g_device;                 // -1 until ng_init
extern int g_sm_count;
extern uint64_t g_launches;          // kernels launched by this library

int set_error(const char* fmt, ...);
int ensure_init();
// Workspace from the caching pool (freed with pool_free on the same stream → safe reuse).
int pool_alloc(void** out, uint64_t bytes);
int pool_free(void* p);
// Data parallel (nccl_dp.cu): world size / rank of this process, and in-line collectives on the
// compute stream used by synchronised BatchNorm.
int dp_world();
int dp_rank();
bool dp_sync_bn();
int dp_allreduce_inline(float* buf, int64_t count);
int dp_allgather_inline(const float* send, float* recv, int64_t count_per_rank);

// Optional per-class kernel timing (ng_prof_*): returns a token (< 0 when disabled).
int prof_begin(int cls, double flops, double bytes);
void prof_end(int token);

// prof_begin / prof_end around a C-ABI call, early returns included
struct ProfScope {
  int tok;
  ProfScope(int cls, double flops, double bytes) : tok(prof_begin(cls, flops, bytes)) {}
  ~ProfScope() { prof_end(tok); }
  ProfScope(const ProfScope&) = delete;
  ProfScope& operator=(const ProfScope&) = delete;
};

template <typename T>
static inline T* dptr(uint64_t p) { return reinterpret_cast<T*>(static_cast<uintptr_t>(p)); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace ng

#define NG_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ng::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,          \
                           cudaGetErrorString(_e));                                      \
  } while (0)

#define NG_REQUIRE(cond, ...)                                                            \
  do {                                                                                   \
    if (!(cond)) return ng::set_error(__VA_ARGS__);                                      \
  } while (0)

#define NG_ENSURE_INIT()                                                                 \
  do {                                                                                   \
    int _r = ng::ensure_init();                                                          \
    if (_r) return _r;                                                                   \
  } while (0)

// Kernel launch on the compute stream.  Template kernels go in parentheses:
//   NG_LAUNCH((kern<4, true>), grid, block, smem, args...)
#ifdef NG_EMU
#define NG_LAUNCH(kernel, grid, block, smem, ...)                                        \
  do {                                                                                   \
    ng_emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); }); \
    ng::g_launches++;                                                                    \
  } while (0)
#define NG_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(ng_emu::dyn_smem_ptr())
#else
#define NG_LAUNCH(kernel, grid, block, smem, ...)                                        \
  do {                                                                                   \
    kernel<<<(grid), (block), (smem), ng::g_stream>>>(__VA_ARGS__);                      \
    ng::g_launches++;                                                                    \
  } while (0)
#define NG_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw[]; \
  type* name = reinterpret_cast<type*>(name##_raw)
#endif

// Division of 0 <= n < 2^31 by a launch-invariant divisor as one wide multiply and a shift (the
// memory-bound kernels otherwise spend more issue slots on 64-bit index divisions than on data):
// m = ceil(2^s / d), s = 31 + ceil(log2 d)  =>  floor(n / d) == (n * m) >> s exactly for n < 2^31.
struct FastDiv {
  uint32_t d, m, s;   // d == 0: not set up (callers keep their generic path)
};
static inline FastDiv make_fastdiv(int64_t d) {
  FastDiv f;
  f.d = f.m = f.s = 0;
  if (d < 1 || d >= (1LL << 31)) return f;
  uint32_t l = 0;
  while ((1ull << l) < (uint64_t)d) ++l;
  f.d = (uint32_t)d;
  f.s = 31 + l;
  f.m = (uint32_t)(((1ull << f.s) + (uint64_t)d - 1) / (uint64_t)d);
  return f;
}
#if defined(__CUDACC__) || defined(NG_EMU)
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f) {
  return (uint32_t)(((unsigned long long)n * f.m) >> f.s);
}
#endif

// 16-byte aligned static shared array (float4 fragment reads)
#ifdef NG_EMU
#define NG_SHARED16 alignas(16) static
#else
#define NG_SHARED16 __shared__ __align__(16)
#endif

#define NG_CHECK_LAUNCH()                                                                \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess)                                                               \
      return ng::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,      \
                           cudaGetErrorString(_e));                                      \
  } while (0)

// ---- small device helpers ------------------------------------------------------------------------
namespace ng {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x a multiple of 32 (<= 1024).  `scratch` holds >= 32 floats.
// Every thread returns the total.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // scratch may still be read from a previous call
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = (lane < nwarp) ? scratch[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// BatchNorm's affine map, z = (x - mean) * (gamma * invstd) + beta, written ONE way (see batchnorm.cu)
__device__ __forceinline__ float bn_affine(float x, float mu, float a, float b) { return fmaf(x - mu, a, b); }

__device__ __forceinline__ bool aligned16(const void* p) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

}  // namespace ng
