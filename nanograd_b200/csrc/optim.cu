// Multi-tensor optimizer steps (SGD / Adam / AdamW) and gradient-bucket pack / unpack.
// One launch updates up to 48 parameter tensors; pointer tables ride in the kernel arguments
// (no host->device copy per step); 128-bit loads/stores, grid = exact chunk count.  HBM-bound: SGD 20 B/param (12 with momentum == 0 in the
// reference's accounting), Adam/AdamW 28 B/param.
//
// The reference writes each step as 5-15 dispatched elementwise Tensor ops per parameter
// (nanograd/optim/optimizer.py:52-55, 79-89, 115-128) and rebinds p.data to a new buffer; here
// parameters and state are updated in place with the same arithmetic.
#include "ng_common.h"

#include <math.h>

namespace ng {

constexpr int kMT = 48;
// One CTA updates one 4096-element chunk of one tensor: 256 threads x four 128-bit vectors per array,
// all loads of a thread issued before the first use (Adam streams 4 arrays in, 3 out).  The CTA -> (tensor,
// chunk) map is a prefix sum over per-tensor chunk counts carried in the kernel arguments, so the grid is
// exactly the work (small bias vectors cost one CTA each instead of a full-width grid row).
constexpr int kChunk = 256 * 4 * 4;

struct MTArgs {
  float* p[kMT];
  const float* g[kMT];
  float* s1[kMT];
  float* s2[kMT];
  long long n[kMT];
  int blk_start[kMT + 1];
};

struct SgdH { float lr, momentum, grad_scale; };
struct AdamH { float lr, beta1, beta2, eps, bc1, bc2, decay_mul, step_size, grad_scale; int adamw; };

struct SgdOp {
  SgdH h;
  static constexpr bool kTwoStates = false;
  __device__ __forceinline__ void operator()(float& p, float g, float& m, float&) const {
    // optimizer.py:54-55: m = m*momentum + lr*g ; p = p - m
    const float mi = m * h.momentum + h.lr * (g * h.grad_scale);
    m = mi;
    p = p - mi;
  }
};

struct AdamOp {
  AdamH h;
  static constexpr bool kTwoStates = true;
  __device__ __forceinline__ void operator()(float& p, float g, float& m, float& v) const {
    const float gi = g * h.grad_scale;
    const float mi = h.beta1 * m + (1.f - h.beta1) * gi;
    const float vi = h.beta2 * v + (1.f - h.beta2) * (gi * gi);
    m = mi;
    v = vi;
    if (h.adamw) {
      // optimizer.py:125-128: p*(1 - lr*wd) - (lr/bc1) * (m / (sqrt(v/bc2) + eps))
      const float denom = sqrtf(vi / h.bc2) + h.eps;
      p = p * h.decay_mul - h.step_size * (mi / denom);
    } else {
      // optimizer.py:86-89: p - lr * (m/bc1) / (sqrt(v/bc2) + eps)
      p = p - h.lr * (mi / h.bc1) / (sqrtf(vi / h.bc2) + h.eps);
    }
  }
};

template <class Op>
__global__ void __launch_bounds__(256) multi_update_kernel(const __grid_constant__ MTArgs a, const Op op) {
  int t = 0;
  while ((int)blockIdx.x >= a.blk_start[t + 1]) ++t;   // <= 48 uniform steps over the argument bank
  float* __restrict__ p = a.p[t];
  const float* __restrict__ g = a.g[t];
  float* __restrict__ m = a.s1[t];
  float* __restrict__ v = Op::kTwoStates ? a.s2[t] : nullptr;
  const long long n = a.n[t];
  const long long base = (long long)((int)blockIdx.x - a.blk_start[t]) * kChunk;
  const long long end = base + kChunk < n ? base + kChunk : n;
  const bool aligned = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)(Op::kTwoStates ? v : p)) & 15) == 0);
  if (aligned && end - base == kChunk) {
    float4 P[4], G[4], M[4], V[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = base + 4 * ((long long)u * 256 + threadIdx.x);
      P[u] = *reinterpret_cast<const float4*>(p + i);
      G[u] = *reinterpret_cast<const float4*>(g + i);
      M[u] = *reinterpret_cast<const float4*>(m + i);
      if (Op::kTwoStates) V[u] = *reinterpret_cast<const float4*>(v + i);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = base + 4 * ((long long)u * 256 + threadIdx.x);
      op(P[u].x, G[u].x, M[u].x, V[u].x);
      op(P[u].y, G[u].y, M[u].y, V[u].y);
      op(P[u].z, G[u].z, M[u].z, V[u].z);
      op(P[u].w, G[u].w, M[u].w, V[u].w);
      *reinterpret_cast<float4*>(m + i) = M[u];
      if (Op::kTwoStates) *reinterpret_cast<float4*>(v + i) = V[u];
      *reinterpret_cast<float4*>(p + i) = P[u];
    }
    return;
  }
  for (long long i = base + threadIdx.x; i < end; i += 256) {
    float pi = p[i], mi = m[i], vi = Op::kTwoStates ? v[i] : 0.f;
    op(pi, g[i], mi, vi);
    m[i] = mi;
    if (Op::kTwoStates) v[i] = vi;
    p[i] = pi;
  }
}

// Fills blk_start from the element counts; returns the grid size.
static inline unsigned mt_plan(MTArgs& a, int c) {
  int blocks = 0;
  for (int i = 0; i < c; ++i) {
    a.blk_start[i] = blocks;
    blocks += (int)ceil_div(a.n[i], (long long)kChunk);
  }
  for (int i = c; i <= kMT; ++i) a.blk_start[i] = blocks;
  return (unsigned)blocks;
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
  double prof_elems_ = 0.0;
  for (int i = 0; i < count; ++i) prof_elems_ += (double)numels[i];
  ProfScope prof_(NG_PROF_OPTIM, 0.0, (momentum != 0.f ? 20.0 : 12.0) * prof_elems_);
  SgdH h{lr, momentum, grad_scale};
  for (int base = 0; base < count; base += kMT) {
    const int c = (count - base < kMT) ? count - base : kMT;
    MTArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < c; ++i) {
      a.p[i] = dptr<float>(params[base + i]);
      a.g[i] = dptr<const float>(grads[base + i]);
      a.s1[i] = dptr<float>(moms[base + i]);
      a.n[i] = numels[base + i];
    }
    const unsigned grid = mt_plan(a, c);
    if (grid == 0) continue;
    NG_LAUNCH((multi_update_kernel<SgdOp>), grid, 256, 0, a, SgdOp{h});
    NG_CHECK_LAUNCH();
  }
  return 0;
}

int ng_multi_adam(int count, const uint64_t* params, const uint64_t* grads, const uint64_t* exp_avg,
                  const uint64_t* exp_avg_sq, const int64_t* numels, float lr, float beta1, float beta2, float eps,
                  float bc1, float bc2, float weight_decay, int adamw, float grad_scale) {
  NG_ENSURE_INIT();
  double prof_elems_ = 0.0;
  for (int i = 0; i < count; ++i) prof_elems_ += (double)numels[i];
  ProfScope prof_(NG_PROF_OPTIM, 0.0, 28.0 * prof_elems_);
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
    for (int i = 0; i < c; ++i) {
      a.p[i] = dptr<float>(params[base + i]);
      a.g[i] = dptr<const float>(grads[base + i]);
      a.s1[i] = dptr<float>(exp_avg[base + i]);
      a.s2[i] = dptr<float>(exp_avg_sq[base + i]);
      a.n[i] = numels[base + i];
    }
    const unsigned grid = mt_plan(a, c);
    if (grid == 0) continue;
    NG_LAUNCH((multi_update_kernel<AdamOp>), grid, 256, 0, a, AdamOp{h});
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
