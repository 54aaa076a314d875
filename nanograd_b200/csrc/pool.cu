// Fused max / average pooling, forward and backward (window = stride, remainder cropped).
//
// The reference has no pooling Function: Tensor._pool2d composes Slice (crop) -> Reshape to
// (N,C,H/ph,ph,W/pw,pw) -> Max or Sum over axes (3,5) -> Mul (nanograd/tensor.py:370-406), and
// the max backward is Max.backward's equality mask with the gradient split equally among
// ties (nanograd/nn/ops_cpu.py:121-127: mask * dy / count).  These kernels do each direction
// in one pass over the activation; HBM-bound, one thread per output window (2x2 windows are
// read as two 64-bit loads).
#include "ng_common.h"

#include <math.h>

namespace ng {

struct PoolP {
  int64_t NC;
  int H, W, ph, pw, OH, OW;
  int fast2;  // ph == pw == 2, W even and base pointers 8-byte aligned
  FastDiv f_ow, f_oh;  // set when the output has fewer than 2^31 elements: index math without 64-bit divisions
};

// BN: the pooled tensor is the output of a training-mode BatchNorm2d (+ ReLU) that was never written
// (nn/buffer.py LazyBuffer): x is that layer's INPUT and every value read is normalised on the fly with the
// expression of the normalise pass (bn_affine), so the window maxima and the equality mask of the backward
// are bit-identical to pooling the materialised tensor.
struct PoolBN {
  const float* gamma;
  const float* beta;
  const float* mean;
  const float* invstd;
  int C, act;
};
__device__ __forceinline__ float pool_bn(float v, float mu, float a, float b, int act) {
  const float z = bn_affine(v, mu, a, b);
  return act ? fmaxf(z, 0.f) : z;
}

template <bool IS_MAX, bool BN = false>
__global__ void __launch_bounds__(256) pool_fwd_kernel(float* __restrict__ y, const float* __restrict__ x, PoolP p,
                                                       PoolBN bn = PoolBN{nullptr, nullptr, nullptr, nullptr, 1, 0}) {
  const int64_t total = p.NC * p.OH * p.OW;
  const float scale = 1.f / (float)(p.ph * p.pw);
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    int ow, oh;
    int64_t nc;
    if (p.f_ow.d) {
      const uint32_t t = fastdiv((uint32_t)o, p.f_ow);
      ow = (int)((uint32_t)o - t * p.f_ow.d);
      const uint32_t u = fastdiv(t, p.f_oh);
      oh = (int)(t - u * p.f_oh.d);
      nc = u;
    } else {
      ow = (int)(o % p.OW);
      const int64_t t = o / p.OW;
      oh = (int)(t % p.OH);
      nc = t / p.OH;
    }
    const float* src = x + (nc * p.H + (int64_t)oh * p.ph) * p.W + (int64_t)ow * p.pw;
    // (window loads first, BatchNorm constants after: all in flight together)
    float2 r0 = make_float2(0.f, 0.f), r1 = make_float2(0.f, 0.f);
    if (p.fast2) {
      r0 = *reinterpret_cast<const float2*>(src);
      r1 = *reinterpret_cast<const float2*>(src + p.W);
    }
    float mu = 0.f, ka = 1.f, kb = 0.f;
    if (BN) {
      const int c = (int)(nc % bn.C);
      mu = __ldg(bn.mean + c); ka = __ldg(bn.gamma + c) * __ldg(bn.invstd + c); kb = __ldg(bn.beta + c);
    }
    float acc;
    if (p.fast2) {
      if (BN) {
        r0.x = pool_bn(r0.x, mu, ka, kb, bn.act); r0.y = pool_bn(r0.y, mu, ka, kb, bn.act);
        r1.x = pool_bn(r1.x, mu, ka, kb, bn.act); r1.y = pool_bn(r1.y, mu, ka, kb, bn.act);
      }
      acc = IS_MAX ? fmaxf(fmaxf(r0.x, r0.y), fmaxf(r1.x, r1.y)) : ((r0.x + r0.y) + (r1.x + r1.y));
    } else {
      acc = IS_MAX ? -INFINITY : 0.f;
      for (int i = 0; i < p.ph; ++i)
        for (int j = 0; j < p.pw; ++j) {
          float v = src[i * p.W + j];
          if (BN) v = pool_bn(v, mu, ka, kb, bn.act);
          acc = IS_MAX ? fmaxf(acc, v) : acc + v;
        }
    }
    y[o] = IS_MAX ? acc : acc * scale;
  }
}

// One thread per output window writes the whole input window of dx.  Rows/columns cropped by
// the forward pass receive zero from a preceding memset (only when H % ph or W % pw != 0).
// 2x2 windows, two horizontally adjacent ones per thread: two 128-bit loads, one 64-bit store (a thread of the generic
// kernel keeps 16 bytes in flight: 3.7 TB/s).  W % 4 == 0, x 16-byte and y 8-byte aligned.
template <bool IS_MAX, bool BN>
__global__ void __launch_bounds__(256) pool2x2_fwd_pair_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                               int64_t NC, int H, int W, FastDiv f_op, FastDiv f_oh,
                                                               PoolBN bn) {
  const int OH = H >> 1, OW = W >> 1, OP = OW >> 1;
  const int64_t total = NC * OH * OP;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    int op, oh;
    int64_t nc;
    if (f_op.d && total < (1LL << 31)) {
      const uint32_t t = fastdiv((uint32_t)o, f_op);
      op = (int)((uint32_t)o - t * f_op.d);
      const uint32_t u = fastdiv(t, f_oh);
      oh = (int)(t - u * f_oh.d);
      nc = u;
    } else {
      op = (int)(o % OP);
      const int64_t t = o / OP;
      oh = (int)(t % OH);
      nc = t / OH;
    }
    const float* src = x + (nc * H + 2 * (int64_t)oh) * W + 4 * op;
    float4 a = *reinterpret_cast<const float4*>(src);
    float4 b = *reinterpret_cast<const float4*>(src + W);
    if (BN) {
      const int c = (int)(nc % bn.C);
      const float mu = __ldg(bn.mean + c), ka = __ldg(bn.gamma + c) * __ldg(bn.invstd + c), kb = __ldg(bn.beta + c);
      a.x = pool_bn(a.x, mu, ka, kb, bn.act); a.y = pool_bn(a.y, mu, ka, kb, bn.act);
      a.z = pool_bn(a.z, mu, ka, kb, bn.act); a.w = pool_bn(a.w, mu, ka, kb, bn.act);
      b.x = pool_bn(b.x, mu, ka, kb, bn.act); b.y = pool_bn(b.y, mu, ka, kb, bn.act);
      b.z = pool_bn(b.z, mu, ka, kb, bn.act); b.w = pool_bn(b.w, mu, ka, kb, bn.act);
    }
    float2 r;
    if (IS_MAX) {
      r.x = fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y));
      r.y = fmaxf(fmaxf(a.z, a.w), fmaxf(b.z, b.w));
    } else {
      r.x = ((a.x + a.y) + (b.x + b.y)) * 0.25f;
      r.y = ((a.z + a.w) + (b.z + b.w)) * 0.25f;
    }
    *reinterpret_cast<float2*>(y + (nc * OH + oh) * OW + 2 * op) = r;
  }
}

// true when the pair kernel applies (and launches it)
template <bool IS_MAX, bool BN>
static bool launch_pool2x2_pair(uint64_t y, uint64_t x, int64_t NC, int H, int W, int ph, int pw, const PoolBN& bn) {
  if (ph != 2 || pw != 2 || H % 2 || W % 4 || (x & 15u) || (y & 7u) || NC <= 0) return false;
  const int64_t n = NC * (H / 2) * (W / 4);
  int64_t g = ceil_div(n, 256);
  const int64_t cap = (int64_t)(g_sm_count > 0 ? g_sm_count : 148) * 16;
  if (g > cap) g = cap;
  NG_LAUNCH((pool2x2_fwd_pair_kernel<IS_MAX, BN>), (int)g, 256, 0, dptr<float>(y), dptr<const float>(x), NC, H, W,
            make_fastdiv(W / 4), make_fastdiv(H / 2), bn);
  return true;
}

template <bool IS_MAX, bool BN = false>
__global__ void __launch_bounds__(256) pool_bwd_kernel(float* __restrict__ dx, const float* __restrict__ dy,
                                                       const float* __restrict__ x, const float* __restrict__ y,
                                                       PoolP p, PoolBN bn = PoolBN{nullptr, nullptr, nullptr, nullptr, 1, 0}) {
  const int64_t total = p.NC * p.OH * p.OW;
  const float scale = 1.f / (float)(p.ph * p.pw);
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    int ow, oh;
    int64_t nc;
    if (p.f_ow.d) {
      const uint32_t t = fastdiv((uint32_t)o, p.f_ow);
      ow = (int)((uint32_t)o - t * p.f_ow.d);
      const uint32_t u = fastdiv(t, p.f_oh);
      oh = (int)(t - u * p.f_oh.d);
      nc = u;
    } else {
      ow = (int)(o % p.OW);
      const int64_t t = o / p.OW;
      oh = (int)(t % p.OH);
      nc = t / p.OH;
    }
    const int64_t base = (nc * p.H + (int64_t)oh * p.ph) * p.W + (int64_t)ow * p.pw;
    const float g = dy[o];
    // (window loads first, BatchNorm constants after: all in flight together)
    float2 r0 = make_float2(0.f, 0.f), r1 = make_float2(0.f, 0.f);
    float m = 0.f;
    if (IS_MAX) {
      m = y[o];
      if (p.fast2) {
        r0 = *reinterpret_cast<const float2*>(x + base);
        r1 = *reinterpret_cast<const float2*>(x + base + p.W);
      }
    }
    float mu = 0.f, ka = 1.f, kb = 0.f;
    if (BN) {
      const int c = (int)(nc % bn.C);
      mu = __ldg(bn.mean + c); ka = __ldg(bn.gamma + c) * __ldg(bn.invstd + c); kb = __ldg(bn.beta + c);
    }
    if (IS_MAX) {
      if (p.fast2) {
        if (BN) {
          r0.x = pool_bn(r0.x, mu, ka, kb, bn.act); r0.y = pool_bn(r0.y, mu, ka, kb, bn.act);
          r1.x = pool_bn(r1.x, mu, ka, kb, bn.act); r1.y = pool_bn(r1.y, mu, ka, kb, bn.act);
        }
        const int e00 = r0.x == m, e01 = r0.y == m, e10 = r1.x == m, e11 = r1.y == m;
        const float share = g / (float)(e00 + e01 + e10 + e11);
        float2 o0, o1;
        o0.x = e00 ? share : 0.f; o0.y = e01 ? share : 0.f;
        o1.x = e10 ? share : 0.f; o1.y = e11 ? share : 0.f;
        *reinterpret_cast<float2*>(dx + base) = o0;
        *reinterpret_cast<float2*>(dx + base + p.W) = o1;
      } else {
        int cnt = 0;
        for (int i = 0; i < p.ph; ++i)
          for (int j = 0; j < p.pw; ++j) {
            float v = x[base + i * p.W + j];
            if (BN) v = pool_bn(v, mu, ka, kb, bn.act);
            cnt += (v == m);
          }
        const float share = g / (float)cnt;
        for (int i = 0; i < p.ph; ++i)
          for (int j = 0; j < p.pw; ++j) {
            const int64_t a = base + i * p.W + j;
            float v = x[a];
            if (BN) v = pool_bn(v, mu, ka, kb, bn.act);
            dx[a] = (v == m) ? share : 0.f;
          }
      }
    } else {
      const float share = g * scale;
      if (p.fast2) {
        float2 v; v.x = share; v.y = share;
        *reinterpret_cast<float2*>(dx + base) = v;
        *reinterpret_cast<float2*>(dx + base + p.W) = v;
      } else {
        for (int i = 0; i < p.ph; ++i)
          for (int j = 0; j < p.pw; ++j) dx[base + i * p.W + j] = share;
      }
    }
  }
}

static int make_pool(PoolP& p, const char* who, int64_t NC, int H, int W, int ph, int pw, uint64_t a, uint64_t b) {
  NG_REQUIRE(NC >= 0 && H > 0 && W > 0 && ph > 0 && pw > 0, "%s: invalid dimensions", who);
  p.NC = NC; p.H = H; p.W = W; p.ph = ph; p.pw = pw;
  p.OH = H / ph; p.OW = W / pw;
  p.fast2 = (ph == 2 && pw == 2 && (W % 2 == 0) && (a % 8 == 0) && (b % 8 == 0)) ? 1 : 0;
  p.f_ow = p.f_oh = FastDiv{0, 0, 0};
  if (p.OH > 0 && p.OW > 0 && NC * p.OH * p.OW < (1LL << 31)) { p.f_ow = make_fastdiv(p.OW); p.f_oh = make_fastdiv(p.OH); }
  return 0;
}

static inline int pgrid(int64_t n) {
  int64_t g = ceil_div(n, 256);
  const int64_t cap = (int64_t)(g_sm_count > 0 ? g_sm_count : 148) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace ng

using namespace ng;

extern "C" {

int ng_maxpool2d_fwd(uint64_t y, uint64_t x, int64_t NC, int H, int W, int ph, int pw) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_POOL, 0.0, (ph > 0 && pw > 0) ? 4.0 * ((double)NC * H * W + (double)NC * (H / ph) * (W / pw)) : 0.0);
  PoolP p;
  if (make_pool(p, "ng_maxpool2d_fwd", NC, H, W, ph, pw, x, x)) return 1;
  const int64_t n = NC * p.OH * p.OW;
  if (n <= 0) return 0;
  if (launch_pool2x2_pair<true, false>(y, x, NC, H, W, ph, pw, PoolBN{nullptr, nullptr, nullptr, nullptr, 1, 0})) {
    NG_CHECK_LAUNCH();
    return 0;
  }
  NG_LAUNCH((pool_fwd_kernel<true>), pgrid(n), 256, 0, dptr<float>(y), dptr<const float>(x), p);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_avgpool2d_fwd(uint64_t y, uint64_t x, int64_t NC, int H, int W, int ph, int pw) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_POOL, 0.0, (ph > 0 && pw > 0) ? 4.0 * ((double)NC * H * W + (double)NC * (H / ph) * (W / pw)) : 0.0);
  PoolP p;
  if (make_pool(p, "ng_avgpool2d_fwd", NC, H, W, ph, pw, x, x)) return 1;
  const int64_t n = NC * p.OH * p.OW;
  if (n <= 0) return 0;
  if (launch_pool2x2_pair<false, false>(y, x, NC, H, W, ph, pw, PoolBN{nullptr, nullptr, nullptr, nullptr, 1, 0})) {
    NG_CHECK_LAUNCH();
    return 0;
  }
  NG_LAUNCH((pool_fwd_kernel<false>), pgrid(n), 256, 0, dptr<float>(y), dptr<const float>(x), p);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_maxpool2d_bwd(uint64_t dx, uint64_t dy, uint64_t x, uint64_t y, int64_t NC, int H, int W, int ph, int pw) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_POOL, 0.0, (ph > 0 && pw > 0) ? 4.0 * (2.0 * (double)NC * H * W + 2.0 * (double)NC * (H / ph) * (W / pw)) : 0.0);
  PoolP p;
  if (make_pool(p, "ng_maxpool2d_bwd", NC, H, W, ph, pw, x, dx)) return 1;
  if (NC * H * W <= 0) return 0;
  if (H % ph || W % pw) NG_CHECK_CUDA(cudaMemsetAsync(dptr<void>(dx), 0, (size_t)(NC * H * W) * sizeof(float), g_stream));
  const int64_t n = NC * p.OH * p.OW;
  if (n <= 0) return 0;
  NG_LAUNCH((pool_bwd_kernel<true>), pgrid(n), 256, 0, dptr<float>(dx), dptr<const float>(dy),
            dptr<const float>(x), dptr<const float>(y), p);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_avgpool2d_bwd(uint64_t dx, uint64_t dy, int64_t NC, int H, int W, int ph, int pw) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_POOL, 0.0, (ph > 0 && pw > 0) ? 4.0 * ((double)NC * H * W + (double)NC * (H / ph) * (W / pw)) : 0.0);
  PoolP p;
  if (make_pool(p, "ng_avgpool2d_bwd", NC, H, W, ph, pw, dx, dx)) return 1;
  if (NC * H * W <= 0) return 0;
  if (H % ph || W % pw) NG_CHECK_CUDA(cudaMemsetAsync(dptr<void>(dx), 0, (size_t)(NC * H * W) * sizeof(float), g_stream));
  const int64_t n = NC * p.OH * p.OW;
  if (n <= 0) return 0;
  NG_LAUNCH((pool_bwd_kernel<false>), pgrid(n), 256, 0, dptr<float>(dx), dptr<const float>(dy),
            (const float*)nullptr, (const float*)nullptr, p);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_maxpool2d_fwd_bn(uint64_t y, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t mean, uint64_t invstd, int64_t NC,
                        int C, int H, int W, int ph, int pw, int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_POOL, 0.0, (ph > 0 && pw > 0) ? 4.0 * ((double)NC * H * W + (double)NC * (H / ph) * (W / pw)) : 0.0);
  PoolP p;
  if (make_pool(p, "ng_maxpool2d_fwd_bn", NC, H, W, ph, pw, x, x)) return 1;
  NG_REQUIRE(C > 0 && NC % C == 0 && (act == 0 || act == 1), "ng_maxpool2d_fwd_bn: bad channel count or act");
  const int64_t n = NC * p.OH * p.OW;
  if (n <= 0) return 0;
  const PoolBN bn{dptr<const float>(gamma), dptr<const float>(beta), dptr<const float>(mean), dptr<const float>(invstd), C, act};
  if (launch_pool2x2_pair<true, true>(y, x, NC, H, W, ph, pw, bn)) {
    NG_CHECK_LAUNCH();
    return 0;
  }
  NG_LAUNCH((pool_fwd_kernel<true, true>), pgrid(n), 256, 0, dptr<float>(y), dptr<const float>(x), p, bn);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_maxpool2d_bwd_bn(uint64_t dx, uint64_t dy, uint64_t x, uint64_t y, uint64_t gamma, uint64_t beta, uint64_t mean,
                        uint64_t invstd, int64_t NC, int C, int H, int W, int ph, int pw, int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_POOL, 0.0,
                  (ph > 0 && pw > 0) ? 4.0 * (2.0 * (double)NC * H * W + 2.0 * (double)NC * (H / ph) * (W / pw)) : 0.0);
  PoolP p;
  if (make_pool(p, "ng_maxpool2d_bwd_bn", NC, H, W, ph, pw, x, dx)) return 1;
  NG_REQUIRE(C > 0 && NC % C == 0 && (act == 0 || act == 1), "ng_maxpool2d_bwd_bn: bad channel count or act");
  if (NC * H * W <= 0) return 0;
  if (H % ph || W % pw) NG_CHECK_CUDA(cudaMemsetAsync(dptr<void>(dx), 0, (size_t)(NC * H * W) * sizeof(float), g_stream));
  const int64_t n = NC * p.OH * p.OW;
  if (n <= 0) return 0;
  const PoolBN bn{dptr<const float>(gamma), dptr<const float>(beta), dptr<const float>(mean), dptr<const float>(invstd), C, act};
  NG_LAUNCH((pool_bwd_kernel<true, true>), pgrid(n), 256, 0, dptr<float>(dx), dptr<const float>(dy), dptr<const float>(x),
            dptr<const float>(y), p, bn);
  NG_CHECK_LAUNCH();
  return 0;
}


}  // extern "C"
