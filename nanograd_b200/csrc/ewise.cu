// Elementwise kernels: unary, broadcasting binary, scalar-immediate binary, fill, accumulate.
// HBM-bound: 128-bit loads/stores on the contiguous paths, grid sized in multiples of the SM
// count with a grid-stride loop.
//
// Replaces the JIT-built OpenCL `binop` / `unop` programs of the reference
// (nanograd/nn/ops_gpu.py:51-85,114-131) and states the same maths as ops_cpu.py:165-287.
#include "ng_common.h"

namespace ng {

template <int OP>
__device__ __forceinline__ float unary_f(float a) {
  if (OP == NG_U_COPY) return a;
  if (OP == NG_U_NEG) return -a;
  if (OP == NG_U_EXP) return expf(a);
  if (OP == NG_U_LOG) return logf(a);
  if (OP == NG_U_RELU) return fmaxf(a, 0.f);
  if (OP == NG_U_SIGMOID) return 1.f / (1.f + expf(-a));
  if (OP == NG_U_TANH) return tanhf(a);
  if (OP == NG_U_SQRT) return sqrtf(a);
  if (OP == NG_U_RECIP) return 1.f / a;
  return a;
}

template <int OP>
__device__ __forceinline__ float binary_f(float a, float b) {
  if (OP == NG_B_ADD) return a + b;
  if (OP == NG_B_MUL) return a * b;
  if (OP == NG_B_POW) return powf(a, b);
  if (OP == NG_B_DIV) return a / b;
  if (OP == NG_B_SUB) return a - b;
  if (OP == NG_B_EQ) return (a == b) ? 1.f : 0.f;
  if (OP == NG_B_RELU_BWD) return (b >= 0.f) ? a : 0.f;  // relu'(0) = 1, ops_cpu.py:263
  if (OP == NG_B_EXP_BWD) return a * expf(b);
  if (OP == NG_B_LOG_BWD) return a / b;
  if (OP == NG_B_SIGMOID_BWD) {
    const float s = 1.f / (1.f + expf(-b));
    return a * s * (1.f - s);
  }
  if (OP == NG_B_TANH_BWD) {
    const float t = tanhf(b);
    return a * (1.f - t * t);
  }
  if (OP == NG_B_POW_DBASE) return b * powf(a, b - 1.f);
  if (OP == NG_B_POW_DEXP) return powf(a, b) * logf(a);
  return a;
}

template <int OP>
__global__ void __launch_bounds__(256) unary_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                    int64_t n, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (int64_t i = tid; i < n4; i += nthreads) {
      float4 v = a4[i];
      v.x = unary_f<OP>(v.x); v.y = unary_f<OP>(v.y); v.z = unary_f<OP>(v.z); v.w = unary_f<OP>(v.w);
      o4[i] = v;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nthreads) out[i] = unary_f<OP>(a[i]);
  } else {
    for (int64_t i = tid; i < n; i += nthreads) out[i] = unary_f<OP>(a[i]);
  }
}

template <int OP>
__global__ void __launch_bounds__(256) binary_same_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                          const float* __restrict__ b, int64_t n, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (int64_t i = tid; i < n4; i += nthreads) {
      const float4 x = a4[i], y = b4[i];
      float4 r;
      r.x = binary_f<OP>(x.x, y.x); r.y = binary_f<OP>(x.y, y.y);
      r.z = binary_f<OP>(x.z, y.z); r.w = binary_f<OP>(x.w, y.w);
      o4[i] = r;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nthreads) out[i] = binary_f<OP>(a[i], b[i]);
  } else {
    for (int64_t i = tid; i < n; i += nthreads) out[i] = binary_f<OP>(a[i], b[i]);
  }
}

template <int OP, bool SCALAR_LHS>
__global__ void __launch_bounds__(256) binary_scalar_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                            float s, int64_t n, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (int64_t i = tid; i < n4; i += nthreads) {
      const float4 x = a4[i];
      float4 r;
      if (SCALAR_LHS) {
        r.x = binary_f<OP>(s, x.x); r.y = binary_f<OP>(s, x.y); r.z = binary_f<OP>(s, x.z); r.w = binary_f<OP>(s, x.w);
      } else {
        r.x = binary_f<OP>(x.x, s); r.y = binary_f<OP>(x.y, s); r.z = binary_f<OP>(x.z, s); r.w = binary_f<OP>(x.w, s);
      }
      o4[i] = r;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nthreads)
      out[i] = SCALAR_LHS ? binary_f<OP>(s, a[i]) : binary_f<OP>(a[i], s);
  } else {
    for (int64_t i = tid; i < n; i += nthreads)
      out[i] = SCALAR_LHS ? binary_f<OP>(s, a[i]) : binary_f<OP>(a[i], s);
  }
}

constexpr int kMaxDims = 8;
struct BcastDesc {
  int ndim;
  int64_t shape[kMaxDims];
  int64_t sa[kMaxDims];
  int64_t sb[kMaxDims];
};

template <int OP>
__global__ void __launch_bounds__(256) binary_bcast_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                           const float* __restrict__ b, int64_t n, BcastDesc d) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n; i += nthreads) {
    int64_t rem = i, oa = 0, ob = 0;
#pragma unroll
    for (int k = kMaxDims - 1; k >= 0; --k) {
      if (k < d.ndim) {
        const int64_t q = rem / d.shape[k];
        const int64_t idx = rem - q * d.shape[k];
        rem = q;
        oa += idx * d.sa[k];
        ob += idx * d.sb[k];
      }
    }
    out[i] = binary_f<OP>(a[oa], b[ob]);
  }
}

__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ out, float v, int64_t n) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n; i += nthreads) out[i] = v;
}

// in-place out += a (no __restrict__: out is read and written)
__global__ void __launch_bounds__(256) accumulate_kernel(float* out, const float* a, int64_t n, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (int64_t i = tid; i < n4; i += nthreads) {
      float4 o = o4[i];
      const float4 x = a4[i];
      o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
      o4[i] = o;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nthreads) out[i] += a[i];
  } else {
    for (int64_t i = tid; i < n; i += nthreads) out[i] += a[i];
  }
}

// ---- fused activations: the reference composes them from primitives --------
//   LeakyReLU  relu(x) - alpha * relu(-x)        module.py:744-772   (3 nodes fwd, 1 + alpha at x = 0 bwd)
//   Mish       x * tanh(log(1 + exp(x)))         module.py:844-864   (6 nodes fwd)
//   Swish      x * sigmoid(beta * x), learnt beta module.py:818-841   (3 nodes fwd; dbeta is a full reduction)
// `pptr` (device, 1 float) overrides the immediate parameter when non-null (Swish's beta lives on the device).
template <int KIND>
__device__ __forceinline__ float act_f(float x, float p) {
  if (KIND == NG_ACT_LEAKY_RELU) return fmaxf(x, 0.f) - p * fmaxf(-x, 0.f);
  if (KIND == NG_ACT_MISH) return x * tanhf(logf(1.f + expf(x)));
  return x / (1.f + expf(-p * x));  // NG_ACT_SWISH
}
// d act / dx
template <int KIND>
__device__ __forceinline__ float act_df(float x, float p) {
  if (KIND == NG_ACT_LEAKY_RELU) return ((x >= 0.f) ? 1.f : 0.f) + ((x <= 0.f) ? p : 0.f);  // ReLU'(0) = 1 on both branches
  if (KIND == NG_ACT_MISH) {
    const float t = tanhf(logf(1.f + expf(x)));
    const float sg = 1.f / (1.f + expf(-x));   // d softplus / dx
    return t + x * (1.f - t * t) * sg;
  }
  const float sg = 1.f / (1.f + expf(-p * x));
  return sg + x * p * sg * (1.f - sg);
}

template <int KIND>
__global__ void __launch_bounds__(256) act_fwd_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                      const float* __restrict__ pptr, float pimm, int64_t n) {
  const float p = pptr ? pptr[0] : pimm;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = act_f<KIND>(x[i], p);
}

template <int KIND>
__global__ void __launch_bounds__(256) act_bwd_kernel(float* __restrict__ dx, const float* __restrict__ dy,
                                                      const float* __restrict__ x, const float* __restrict__ pptr,
                                                      float pimm, int64_t n) {
  const float p = pptr ? pptr[0] : pimm;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = dy[i] * act_df<KIND>(x[i], p);
}

// Swish: partial[b] = sum over the block's elements of dy * x^2 * s(1 - s), s = sigmoid(beta x)
__global__ void __launch_bounds__(256) swish_dbeta_kernel(float* __restrict__ partial, const float* __restrict__ dy,
                                                          const float* __restrict__ x, const float* __restrict__ beta,
                                                          int64_t n) {
  __shared__ float scratch[32];
  const float b = beta[0];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[i];
    const float sg = 1.f / (1.f + expf(-b * xv));
    acc += dy[i] * xv * xv * sg * (1.f - sg);
  }
  const float t = block_sum(acc, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// out[0] = scale * sum_i partial[i] (one block, fixed order)
__global__ void __launch_bounds__(256) sum_partials_kernel(float* __restrict__ out, const float* __restrict__ partial,
                                                           int m, float scale) {
  __shared__ float scratch[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < m; i += blockDim.x) acc += partial[i];
  const float t = block_sum(acc, scratch);
  if (threadIdx.x == 0) out[0] = t * scale;
}

static inline int ew_grid(int64_t work_items) {
  // enough CTAs for >= 2 waves at 8 resident CTAs/SM, never more than the work needs
  int64_t want = ceil_div(work_items, 256);
  int64_t cap = (int64_t)(g_sm_count > 0 ? g_sm_count : 148) * 16;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

static inline bool host_aligned16(uint64_t p) { return (p & 15u) == 0; }

}  // namespace ng

using namespace ng;

#define NG_UNARY_CASE(OPC)                                                                     \
  case OPC:                                                                                    \
    NG_LAUNCH((unary_kernel<OPC>), grid, 256, 0, dptr<float>(out), dptr<const float>(a), n, vec); \
    break;

#define NG_BIN_CASES(X)                                                                        \
  X(NG_B_ADD) X(NG_B_MUL) X(NG_B_POW) X(NG_B_DIV) X(NG_B_SUB) X(NG_B_EQ) X(NG_B_RELU_BWD)      \
  X(NG_B_EXP_BWD) X(NG_B_LOG_BWD) X(NG_B_SIGMOID_BWD) X(NG_B_TANH_BWD) X(NG_B_POW_DBASE)       \
  X(NG_B_POW_DEXP)

extern "C" {

int ng_unary(int op, uint64_t out, uint64_t a, int64_t n) {
  NG_ENSURE_INIT();
  if (n <= 0) return 0;
  const int vec = host_aligned16(out) && host_aligned16(a);
  const int grid = ew_grid(vec ? ceil_div(n, 4) : n);
  switch (op) {
    NG_UNARY_CASE(NG_U_COPY) NG_UNARY_CASE(NG_U_NEG) NG_UNARY_CASE(NG_U_EXP) NG_UNARY_CASE(NG_U_LOG)
    NG_UNARY_CASE(NG_U_RELU) NG_UNARY_CASE(NG_U_SIGMOID) NG_UNARY_CASE(NG_U_TANH)
    NG_UNARY_CASE(NG_U_SQRT) NG_UNARY_CASE(NG_U_RECIP)
    default:
      return set_error("ng_unary: unknown op %d", op);
  }
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_binary(int op, uint64_t out, uint64_t a, uint64_t b, int ndim, const int64_t* out_shape,
              const int64_t* a_strides, const int64_t* b_strides) {
  NG_ENSURE_INIT();
  NG_REQUIRE(ndim >= 0 && ndim <= kMaxDims, "ng_binary: ndim %d not in [0, %d]", ndim, kMaxDims);
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= out_shape[i];
  if (n <= 0) return 0;
  // contiguous same-shape operands → vector path
  bool same = true;
  {
    int64_t st = 1;
    for (int i = ndim - 1; i >= 0; --i) {
      if (out_shape[i] != 1 && (a_strides[i] != st || b_strides[i] != st)) same = false;
      st *= out_shape[i];
    }
  }
  if (same) {
    const int vec = host_aligned16(out) && host_aligned16(a) && host_aligned16(b);
    const int grid = ew_grid(vec ? ceil_div(n, 4) : n);
    switch (op) {
#define X(OPC)                                                                                   \
  case OPC:                                                                                      \
    NG_LAUNCH((binary_same_kernel<OPC>), grid, 256, 0, dptr<float>(out), dptr<const float>(a),   \
              dptr<const float>(b), n, vec);                                                     \
    break;
      NG_BIN_CASES(X)
#undef X
      default:
        return set_error("ng_binary: unknown op %d", op);
    }
  } else {
    BcastDesc d;
    memset(&d, 0, sizeof(d));
    d.ndim = ndim;
    for (int i = 0; i < ndim; ++i) {
      d.shape[i] = out_shape[i];
      d.sa[i] = a_strides[i];
      d.sb[i] = b_strides[i];
    }
    const int grid = ew_grid(n);
    switch (op) {
#define X(OPC)                                                                                   \
  case OPC:                                                                                      \
    NG_LAUNCH((binary_bcast_kernel<OPC>), grid, 256, 0, dptr<float>(out), dptr<const float>(a),  \
              dptr<const float>(b), n, d);                                                       \
    break;
      NG_BIN_CASES(X)
#undef X
      default:
        return set_error("ng_binary: unknown op %d", op);
    }
  }
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_binary_scalar(int op, uint64_t out, uint64_t a, float s, int scalar_is_lhs, int64_t n) {
  NG_ENSURE_INIT();
  if (n <= 0) return 0;
  const int vec = host_aligned16(out) && host_aligned16(a);
  const int grid = ew_grid(vec ? ceil_div(n, 4) : n);
  switch (op) {
#define X(OPC)                                                                                   \
  case OPC:                                                                                      \
    if (scalar_is_lhs)                                                                           \
      NG_LAUNCH((binary_scalar_kernel<OPC, true>), grid, 256, 0, dptr<float>(out),               \
                dptr<const float>(a), s, n, vec);                                                \
    else                                                                                         \
      NG_LAUNCH((binary_scalar_kernel<OPC, false>), grid, 256, 0, dptr<float>(out),              \
                dptr<const float>(a), s, n, vec);                                                \
    break;
    NG_BIN_CASES(X)
#undef X
    default:
      return set_error("ng_binary_scalar: unknown op %d", op);
  }
  NG_CHECK_LAUNCH();
  return 0;
}

#define NG_ACT_SWITCH(kind, STMT)                                         \
  switch (kind) {                                                        \
    case NG_ACT_LEAKY_RELU: { constexpr int KK = NG_ACT_LEAKY_RELU; STMT; break; } \
    case NG_ACT_MISH: { constexpr int KK = NG_ACT_MISH; STMT; break; }   \
    case NG_ACT_SWISH: { constexpr int KK = NG_ACT_SWISH; STMT; break; } \
    default: return set_error("ng_act: unknown activation %d", kind);    \
  }

int ng_act_fwd(int kind, uint64_t y, uint64_t x, uint64_t param_ptr, float param, int64_t n) {
  NG_ENSURE_INIT();
  if (n <= 0) return 0;
  const float* pp = param_ptr ? dptr<const float>(param_ptr) : nullptr;
  NG_ACT_SWITCH(kind, NG_LAUNCH((act_fwd_kernel<KK>), ew_grid(n), 256, 0, dptr<float>(y), dptr<const float>(x), pp, param, n));
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_act_bwd(int kind, uint64_t dx, uint64_t dy, uint64_t x, uint64_t param_ptr, float param, int64_t n) {
  NG_ENSURE_INIT();
  if (n <= 0) return 0;
  const float* pp = param_ptr ? dptr<const float>(param_ptr) : nullptr;
  NG_ACT_SWITCH(kind, NG_LAUNCH((act_bwd_kernel<KK>), ew_grid(n), 256, 0, dptr<float>(dx), dptr<const float>(dy),
                                dptr<const float>(x), pp, param, n));
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_swish_dbeta(uint64_t dbeta1, uint64_t dy, uint64_t x, uint64_t beta_ptr, int64_t n) {
  NG_ENSURE_INIT();
  NG_REQUIRE(beta_ptr != 0, "ng_swish_dbeta: beta must be a device pointer");
  int blocks = ew_grid(n > 0 ? n : 1);
  if (blocks > 1024) blocks = 1024;
  float* partial = nullptr;
  int r = pool_alloc((void**)&partial, (uint64_t)blocks * sizeof(float));
  if (r) return r;
  NG_LAUNCH(swish_dbeta_kernel, blocks, 256, 0, partial, dptr<const float>(dy), dptr<const float>(x),
            dptr<const float>(beta_ptr), n);
  NG_CHECK_LAUNCH();
  NG_LAUNCH(sum_partials_kernel, 1, 256, 0, dptr<float>(dbeta1), (const float*)partial, blocks, 1.f);
  NG_CHECK_LAUNCH();
  pool_free(partial);
  return 0;
}

int ng_fill(uint64_t out, float value, int64_t n) {
  NG_ENSURE_INIT();
  if (n <= 0) return 0;
  NG_LAUNCH(fill_kernel, ew_grid(n), 256, 0, dptr<float>(out), value, n);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_accumulate(uint64_t out, uint64_t a, int64_t n) {
  NG_ENSURE_INIT();
  if (n <= 0) return 0;
  const int vec = host_aligned16(out) && host_aligned16(a);
  const int grid = ew_grid(vec ? ceil_div(n, 4) : n);
  NG_LAUNCH(accumulate_kernel, grid, 256, 0, dptr<float>(out), dptr<const float>(a), n, vec);
  NG_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
