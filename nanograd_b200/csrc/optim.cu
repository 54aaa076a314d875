// Multi-tensor optimizer steps (SGD / Adam / AdamW) and gradient-bucket pack / unpack.
// One launch updates up to 48 parameter tensors; pointer tables ride in the kernel arguments
// (no host->device copy per step).  HBM-bound: SGD 20 B/param (12 with momentum == 0 in the
// reference's accounting), Adam/AdamW 28 B/param.
//
// The reference writes each step as 5-15 dispatched elementwise Tensor ops per parameter
// (nanograd/optim/optimizer.py:52-55, 79-89, 115-128) and rebinds p.data to a new buffer; here
// parameters and state are updated in place with the same arithmetic.
#include "ng_common.h"

#include <math.h>

namespace ng {

constexpr int kMT = 48;

struct MTArgs {
  float* p[kMT];
  const float* g[kMT];
  float* s1[kMT];
  float* s2[kMT];
  long long n[kMT];
};

struct SgdH { float lr, momentum, grad_scale; };
struct AdamH { float lr, beta1, beta2, eps, bc1, bc2, decay_mul, step_size, grad_scale; int adamw; };

__global__ void __launch_bounds__(256) multi_sgd_kernel(const __grid_constant__ MTArgs a, SgdH h) {
  const int t = blockIdx.y;
  float* p = a.p[t];
  const float* g = a.g[t];
  float* m = a.s1[t];
  const long long n = a.n[t];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // optimizer.py:54-55: m = m*momentum + lr*g ; p = p - m
    const float mi = m[i] * h.momentum + h.lr * (g[i] * h.grad_scale);
    m[i] = mi;
    p[i] = p[i] - mi;
  }
}

__global__ void __launch_bounds__(256) multi_adam_kernel(const __grid_constant__ MTArgs a, AdamH h) {
  const int t = blockIdx.y;
  float* p = a.p[t];
  const float* g = a.g[t];
  float* m = a.s1[t];
  float* v = a.s2[t];
  const long long n = a.n[t];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * h.grad_scale;
    const float mi = h.beta1 * m[i] + (1.f - h.beta1) * gi;
    const float vi = h.beta2 * v[i] + (1.f - h.beta2) * (gi * gi);
    m[i] = mi;
    v[i] = vi;
    if (h.adamw) {
      // optimizer.py:125-128: p*(1 - lr*wd) - (lr/bc1) * (m / (sqrt(v/bc2) + eps))
      const float denom = sqrtf(vi / h.bc2) + h.eps;
      p[i] = p[i] * h.decay_mul - h.step_size * (mi / denom);
    } else {
      // optimizer.py:86-89: p - lr * (m/bc1) / (sqrt(v/bc2) + eps)
      p[i] = p[i] - h.lr * (mi / h.bc1) / (sqrtf(vi / h.bc2) + h.eps);
    }
  }
}

// PACK: flat[off_t + i] = src_t[i];  UNPACK: dst_t[i] = flat[off_t + i]  (offsets in s2 slot as integers)
struct PackArgs {
  float* t[kMT];
  long long n[kMT];
  long long off[kMT];
};
template <bool PACK>
__global__ void __launch_bounds__(256) multi_pack_kernel(const __grid_constant__ PackArgs a, float* flat) {
  const int t = blockIdx.y;
  float* x = a.t[t];
  float* f = flat + a.off[t];
  const long long n = a.n[t];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (PACK) f[i] = x[i];
    else x[i] = f[i];
  }
}

static inline unsigned mt_grid_x(long long max_n) {
  long long g = ceil_div(max_n, 256 * 4);
  const long long cap = (long long)(g_sm_count > 0 ? g_sm_count : 148) * 4;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace ng

using namespace ng;

extern "C" {

int ng_multi_sgd(int count, const uint64_t* params, const uint64_t* grads, const uint64_t* moms,
                 const int64_t* numels, float lr, float momentum, float grad_scale) {
  NG_ENSURE_INIT();
  SgdH h{lr, momentum, grad_scale};
  for (int base = 0; base < count; base += kMT) {
    const int c = (count - base < kMT) ? count - base : kMT;
    MTArgs a;
    memset(&a, 0, sizeof(a));
    long long max_n = 0;
    for (int i = 0; i < c; ++i) {
      a.p[i] = dptr<float>(params[base + i]);
      a.g[i] = dptr<const float>(grads[base + i]);
      a.s1[i] = dptr<float>(moms[base + i]);
      a.n[i] = numels[base + i];
      if (a.n[i] > max_n) max_n = a.n[i];
    }
    if (max_n == 0) continue;
    NG_LAUNCH(multi_sgd_kernel, dim3(mt_grid_x(max_n), (unsigned)c), 256, 0, a, h);
    NG_CHECK_LAUNCH();
  }
  return 0;
}

int ng_multi_adam(int count, const uint64_t* params, const uint64_t* grads, const uint64_t* exp_avg,
                  const uint64_t* exp_avg_sq, const int64_t* numels, float lr, float beta1, float beta2, float eps,
                  float bc1, float bc2, float weight_decay, int adamw, float grad_scale) {
  NG_ENSURE_INIT();
  AdamH h;
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.bc1 = bc1; h.bc2 = bc2;
  h.decay_mul = 1.f - lr * weight_decay;
  h.step_size = lr / bc1;
  h.grad_scale = grad_scale;
  h.adamw = adamw;
  for (int base = 0; base < count; base += kMT) {
    const int c = (count - base < kMT) ? count - base : kMT;
    MTArgs a;
    memset(&a, 0, sizeof(a));
    long long max_n = 0;
    for (int i = 0; i < c; ++i) {
      a.p[i] = dptr<float>(params[base + i]);
      a.g[i] = dptr<const float>(grads[base + i]);
      a.s1[i] = dptr<float>(exp_avg[base + i]);
      a.s2[i] = dptr<float>(exp_avg_sq[base + i]);
      a.n[i] = numels[base + i];
      if (a.n[i] > max_n) max_n = a.n[i];
    }
    if (max_n == 0) continue;
    NG_LAUNCH(multi_adam_kernel, dim3(mt_grid_x(max_n), (unsigned)c), 256, 0, a, h);
    NG_CHECK_LAUNCH();
  }
  return 0;
}

static int pack_common(bool pack, int count, const uint64_t* ts, const int64_t* numels, uint64_t flat) {
  long long off = 0;
  for (int base = 0; base < count; base += kMT) {
    const int c = (count - base < kMT) ? count - base : kMT;
    PackArgs a;
    memset(&a, 0, sizeof(a));
    long long max_n = 0;
    for (int i = 0; i < c; ++i) {
      a.t[i] = dptr<float>(ts[base + i]);
      a.n[i] = numels[base + i];
      a.off[i] = off;
      off += numels[base + i];
      if (a.n[i] > max_n) max_n = a.n[i];
    }
    if (max_n == 0) continue;
    if (pack) NG_LAUNCH((multi_pack_kernel<true>), dim3(mt_grid_x(max_n), (unsigned)c), 256, 0, a, dptr<float>(flat));
    else NG_LAUNCH((multi_pack_kernel<false>), dim3(mt_grid_x(max_n), (unsigned)c), 256, 0, a, dptr<float>(flat));
    NG_CHECK_LAUNCH();
  }
  return 0;
}

int ng_multi_pack(int count, const uint64_t* srcs, const int64_t* numels, uint64_t dst_flat) {
  NG_ENSURE_INIT();
  return pack_common(true, count, srcs, numels, dst_flat);
}

int ng_multi_unpack(int count, const uint64_t* dsts, const int64_t* numels, uint64_t src_flat) {
  NG_ENSURE_INIT();
  return pack_common(false, count, dsts, numels, src_flat);
}

}  // extern "C"
