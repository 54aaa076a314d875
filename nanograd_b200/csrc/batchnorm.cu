// Fused BatchNorm2d: training forward (one statistics pass + one normalise pass, running-stat
// update in place), evaluation forward, and backward (one reduction pass + one dx pass).
//
// The reference composes BatchNorm2d.forward / _normalize from ~30 primitive Tensor ops
// (nanograd/nn/module.py:356-391): mean over (0,2,3); biased variance normalises; the unbiased
// variance sum/(NHW-1) feeds running_var; running = (1-m)*running + m*batch; eps inside sqrt.
// Algorithmic traffic: forward 12 B/element (x read twice, y written), backward 20 B/element.
// Statistics are accumulated around a per-channel shift (the channel's first element) so the
// single pass does not suffer E[x^2]-E[x]^2 cancellation on offset data.
#include "ng_common.h"

#include <math.h>

namespace ng {

// bn_affine (ng_common.h): z = (x - mean) * (gamma * invstd) + beta, written ONE way so the forward kernel, the
// backward kernels that recompute z for the fused-ReLU mask and the fused staging kernels of the FP16x3
// convolution engine (umma_conv_f16.cu) produce bit-identical values.

// Block-wide maximum for blockDim.x a multiple of 32 (<= 1024); thread 0 holds the result.
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = (lane < nwarp) ? scratch[lane] : -3.4e38f;
  return warp_max(t);
}

// Per-channel fold of the split partials, shared by the stand-alone finalize kernels (synchronised BatchNorm keeps
// a collective between the reduction and the fold) and by the LAST block of a channel inside bn_reduce_kernel
// (single process: the fold rides on the reduction launch; 12 launches of 5-8 us each per training step were
// nothing but this fold).  Partials are read with ld.global.cg: they were written by other blocks.
struct BnFin {
  unsigned int* counter;   // [C] arrival counters, zero between launches; null = no fused fold
  int S, PS, act;
  float count, eps, momentum;
  const float* gamma;
  const float* beta;
  unsigned int* bits;      // forward: bits of max|y|; backward: bits of a bound of max|dx| (null: not wanted)
  float* save_mean;        // forward
  float* save_invstd;
  float* snap_gamma;       // forward, optional: copies of gamma / beta as they are now (a recipe evaluated after an
  float* snap_beta;        // in-place optimizer step must not see the updated parameters)
  float* running_mean;
  float* running_var;
  const float* invstd;     // backward
  float* dgamma;
  float* dbeta;
  float* sums;
};

__device__ __forceinline__ void bn_fwd_fold(const float* partial, const float* x, int64_t c, int64_t C, int64_t P,
                                            const BnFin& f) {
  // PS == 4 (bits given): the partials carry max (x - shift) and max (shift - x); the normalised output is
  // monotone in x, so max |y| of the channel is attained at one of the two extremes and *bits becomes the
  // bits of max |y| over the tensor (y = the output of the apply pass, ReLU included, evaluated with bn_affine)
  float sa = 0.f, sb = 0.f, hi = 0.f, lo = 0.f;
  for (int s = 0; s < f.S; ++s) {
    sa += __ldcg(partial + (s * C + c) * f.PS + 0);
    sb += __ldcg(partial + (s * C + c) * f.PS + 1);
    if (f.PS == 4) {
      hi = fmaxf(hi, __ldcg(partial + (s * C + c) * f.PS + 2));
      lo = fmaxf(lo, __ldcg(partial + (s * C + c) * f.PS + 3));
    }
  }
  const float shift = x[c * P];
  const float d = sa / f.count;             // mean - shift
  const float mean = shift + d;
  float ssd = sb - sa * d;                  // sum (x-mean)^2 = sum a^2 - (sum a)^2 / n
  if (ssd < 0.f) ssd = 0.f;
  const float var_b = ssd / f.count;        // module.py:366
  const float invstd = 1.f / sqrtf(var_b + f.eps);
  f.save_mean[c] = mean;
  f.save_invstd[c] = invstd;
  if (f.snap_gamma) {
    f.snap_gamma[c] = f.gamma[c];
    f.snap_beta[c] = f.beta[c];
  }
  if (f.running_mean) {
    const float var_u = ssd / (f.count - 1.f);  // module.py:367-368
    f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mean;
    f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * var_u;
  }
  if (f.bits) {
    const float a = f.gamma[c] * invstd, b = f.beta[c];
    const float z0 = bn_affine(shift + hi, mean, a, b), z1 = bn_affine(shift - lo, mean, a, b);
    // (shift + hi and shift - lo reproduce the extreme elements up to rounding: a hair of slack is added)
    float m = f.act ? fmaxf(fmaxf(z0, z1), 0.f) : fmaxf(fabsf(z0), fabsf(z1));
    m *= 1.0001f;
    if (m > 0.f && m < 3.0e38f) atomicMax(f.bits, __float_as_uint(m));
  }
}

__device__ __forceinline__ void bn_bwd_fold(const float* partial, int64_t c, int64_t C, const BnFin& f) {
  // PS == 4 (bits given, single process): *bits becomes the bits of an upper bound of max |dx| over the tensor,
  // |dx| <= |k0| (max |dy| + |k1| + max |x - mean| |k2|) per channel with the k's of bn_bwd_dx_kernel
  if (c == 0) f.sums[2 * C] = f.count;  // summed across ranks with the two per-channel sums (synchronised BatchNorm)
  float sa = 0.f, sb = 0.f, mg = 0.f, mx = 0.f;
  for (int s = 0; s < f.S; ++s) {
    sa += __ldcg(partial + (s * C + c) * f.PS + 0);
    sb += __ldcg(partial + (s * C + c) * f.PS + 1);
    if (f.PS == 4) {
      mg = fmaxf(mg, __ldcg(partial + (s * C + c) * f.PS + 2));
      mx = fmaxf(mx, __ldcg(partial + (s * C + c) * f.PS + 3));
    }
  }
  f.dbeta[c] = sa;
  f.dgamma[c] = sb * f.invstd[c];
  f.sums[c * 2 + 0] = sa;  // sum dy
  f.sums[c * 2 + 1] = sb;  // sum dy * (x - mean)
  if (f.bits) {
    const float inv = f.invstd[c], inv_count = 1.f / f.count;
    const float k0 = f.gamma[c] * inv, k1 = sa * inv_count, k2 = inv * inv * sb * inv_count;
    const float m = fabsf(k0) * (mg + fabsf(k1) + mx * fabsf(k2)) * 1.0001f;
    if (m > 0.f && m < 3.0e38f) atomicMax(f.bits, __float_as_uint(m));
  }
}

// MODE 0: a = x - shift[c], b = a*a          (forward statistics)
// MODE 1: a = dy,           b = dy*(x-mean)  (backward reductions)
// MODE 2: like MODE 1 with dy masked by the fused ReLU: dy * (z >= 0), z recomputed from x
// partial layout: [S][C][2]; grid = (C, S)
// EXTRA: two more partials per (split, channel), folded with max by the *_extra finalize kernels, from which the
// magnitude of the pass's OUTPUT is known before it is written (the FP16x3 convolution engine stages its operands
// with one power-of-two scale per tensor):  MODE 0: max (x - shift), max (shift - x);  MODE 1/2: max |dy (masked)|,
// max |x - mean|.  partial layout [S][C][4].
template <int MODE, bool EXTRA = false>
__global__ void __launch_bounds__(256) bn_reduce_kernel(float* __restrict__ partial, const float* __restrict__ x,
                                                        const float* __restrict__ dy, const float* __restrict__ mean,
                                                        int64_t N, int64_t C, int64_t P, int64_t n_per_split,
                                                        int vec_ok, const float* __restrict__ gamma = nullptr,
                                                        const float* __restrict__ beta = nullptr,
                                                        const float* __restrict__ invstd = nullptr, FastDiv fp4 = FastDiv{0, 0, 0},
                                                        const BnFin fin = BnFin{}) {
  __shared__ float scratch[32];
  const int64_t c = blockIdx.x, s = blockIdx.y;
  float za = 0.f, zb = 0.f;
  if (MODE == 2) { za = gamma[c] * invstd[c]; zb = beta[c]; }
  const int64_t n0 = s * n_per_split;
  int64_t n1 = n0 + n_per_split;
  if (n1 > N) n1 = N;
  const float ref = (MODE == 0) ? x[c * P] : mean[c];
  float sa = 0.f, sb = 0.f;
  float ea = 0.f, eb = 0.f;   // EXTRA maxima (all four candidates are >= 0 at the reference point)
  if (vec_ok) {
    const int64_t P4 = P >> 2;
    const int64_t items = (n1 - n0) * P4;
    const bool fast = fp4.d != 0 && items < (1LL << 31);
    // Several iterations' loads are issued before any is consumed: with one 16-byte load in flight per thread
    // the pass was latency-bound (3.7 TB/s measured).  Out-of-range slots read nothing and add zero.
    constexpr int U = (MODE == 0) ? 4 : 2;   // the backward passes load two tensors: same bytes in flight
    for (int64_t i0 = threadIdx.x; i0 < items; i0 += U * (int64_t)blockDim.x) {
      float4 xv[U], gv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * (int64_t)blockDim.x;
        xv[u].x = xv[u].y = xv[u].z = xv[u].w = ref;
        gv[u].x = gv[u].y = gv[u].z = gv[u].w = 0.f;
        if (i < items) {
          int64_t n, p4;
          if (fast) {
            const uint32_t q = fastdiv((uint32_t)i, fp4);
            n = n0 + q;
            p4 = (uint32_t)i - q * fp4.d;
          } else {
            n = n0 + i / P4;
            p4 = i % P4;
          }
          const int64_t off = (n * C + c) * P + (p4 << 2);
          xv[u] = *reinterpret_cast<const float4*>(x + off);
          if (MODE != 0) gv[u] = *reinterpret_cast<const float4*>(dy + off);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (MODE == 0) {
          const float a0 = xv[u].x - ref, a1 = xv[u].y - ref, a2 = xv[u].z - ref, a3 = xv[u].w - ref;
          sa += (a0 + a1) + (a2 + a3);
          sb += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
          if (EXTRA) {
            ea = fmaxf(ea, fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)));
            eb = fmaxf(eb, -fminf(fminf(a0, a1), fminf(a2, a3)));
          }
        } else {
          float4 g = gv[u];
          if (MODE == 2) {
            if (!(bn_affine(xv[u].x, ref, za, zb) >= 0.f)) g.x = 0.f;
            if (!(bn_affine(xv[u].y, ref, za, zb) >= 0.f)) g.y = 0.f;
            if (!(bn_affine(xv[u].z, ref, za, zb) >= 0.f)) g.z = 0.f;
            if (!(bn_affine(xv[u].w, ref, za, zb) >= 0.f)) g.w = 0.f;
          }
          sa += (g.x + g.y) + (g.z + g.w);
          sb += (g.x * (xv[u].x - ref) + g.y * (xv[u].y - ref)) + (g.z * (xv[u].z - ref) + g.w * (xv[u].w - ref));
          if (EXTRA) {
            ea = fmaxf(ea, fmaxf(fmaxf(fabsf(g.x), fabsf(g.y)), fmaxf(fabsf(g.z), fabsf(g.w))));
            eb = fmaxf(eb, fmaxf(fmaxf(fabsf(xv[u].x - ref), fabsf(xv[u].y - ref)),
                                 fmaxf(fabsf(xv[u].z - ref), fabsf(xv[u].w - ref))));
          }
        }
      }
    }
  } else {
    const int64_t items = (n1 - n0) * P;
    for (int64_t i = threadIdx.x; i < items; i += blockDim.x) {
      const int64_t n = n0 + i / P, p = i % P;
      const int64_t off = (n * C + c) * P + p;
      if (MODE == 0) {
        const float a = x[off] - ref;
        sa += a;
        sb += a * a;
        if (EXTRA) { ea = fmaxf(ea, a); eb = fmaxf(eb, -a); }
      } else {
        float g = dy[off];
        if (MODE == 2 && !(bn_affine(x[off], ref, za, zb) >= 0.f)) g = 0.f;
        sa += g;
        sb += g * (x[off] - ref);
        if (EXTRA) { ea = fmaxf(ea, fabsf(g)); eb = fmaxf(eb, fabsf(x[off] - ref)); }
      }
    }
  }
  constexpr int PS = EXTRA ? 4 : 2;
  const float ta = block_sum(sa, scratch);
  const float tb = block_sum(sb, scratch);
  if (threadIdx.x == 0) {
    partial[(s * C + c) * PS + 0] = ta;
    partial[(s * C + c) * PS + 1] = tb;
  }
  if (EXTRA) {
    const float ma = block_max(ea, scratch);
    const float mb = block_max(eb, scratch);
    if (threadIdx.x == 0) {
      partial[(s * C + c) * PS + 2] = ma;
      partial[(s * C + c) * PS + 3] = mb;
    }
  }
  if (fin.counter && threadIdx.x == 0) {
    // the last block of the channel to arrive folds the channel's partials (and re-arms the counter)
    __threadfence();
    const unsigned int prev = atomicAdd(fin.counter + c, 1u);
    if (prev == gridDim.y - 1) {
      __threadfence();
      if (MODE == 0) bn_fwd_fold(partial, x, c, C, P, fin);
      else bn_bwd_fold(partial, c, C, fin);
      fin.counter[c] = 0u;
    }
  }
}

// Backward reduction of a BatchNorm2d (+ ReLU) layer whose output went through a 2x2 max pooling, with the pooling
// backward fused in: the incoming gradient is never written.  Per window the thread recomputes the four activations
// y = act(bn_affine(x)) from the layer input, splits the pooled gradient among the elements equal to the pooled
// maximum (Max.backward's tie rule, ops_cpu.py:121-127 — the same float32 values as ng_maxpool2d_bwd_bn sees),
// applies the ReLU mask (z >= 0) and accumulates sum dy, sum dy (x - mean), max |dy|, max |x - mean|.
// Replaces pool_bwd (2.5 passes over the tensor) + the dy read of bn_reduce_kernel<2> by two quarter-size reads.
__global__ void __launch_bounds__(256) bn_reduce_pool_kernel(float* __restrict__ partial, const float* __restrict__ x,
                                                             const float* __restrict__ dpooled,
                                                             const float* __restrict__ pooled,
                                                             const float* __restrict__ mean, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ invstd,
                                                             int64_t N, int64_t C, int H, int W, int64_t n_per_split, int act,
                                                             FastDiv f_win, FastDiv f_ow, const BnFin fin) {
  __shared__ float scratch[32];
  const int64_t c = blockIdx.x, s = blockIdx.y;
  const float mu = mean[c], za = gamma[c] * invstd[c], zb = beta[c];
  const int64_t n0 = s * n_per_split;
  int64_t n1 = n0 + n_per_split;
  if (n1 > N) n1 = N;
  const int OW = W >> 1, OH = H >> 1;
  const int64_t wins = (int64_t)OH * OW;
  // one item = two horizontally adjacent windows: two 128-bit loads of x, one 64-bit load each of the pooled maxima
  // and their gradient (W % 4 == 0); two items' loads are in flight before the first is consumed
  const int64_t pairs = wins >> 1;
  const int64_t items = (n1 - n0) * pairs;
  float sa = 0.f, sb = 0.f, ea = 0.f, eb = 0.f;
  constexpr int U = 2;
  for (int64_t i0 = threadIdx.x; i0 < items; i0 += U * (int64_t)blockDim.x) {
    float4 r0[U], r1[U];
    float2 pm[U], pg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * (int64_t)blockDim.x;
      r0[u] = r1[u] = make_float4(mu, mu, mu, mu);
      pm[u] = make_float2(3.0e38f, 3.0e38f);   // out-of-range slot: no element equals the "maximum", nothing is accumulated
      pg[u] = make_float2(0.f, 0.f);
      if (i < items) {
        const uint32_t q = fastdiv((uint32_t)i, f_win);          // image within the split
        const uint32_t wi = (uint32_t)i - q * f_win.d;           // window pair within the image
        const uint32_t oh = fastdiv(wi, f_ow), op = wi - oh * f_ow.d;
        const int64_t plane = (n0 + q) * C + c;
        const float* src = x + (plane * H + 2 * (int64_t)oh) * W + 4 * op;
        r0[u] = *reinterpret_cast<const float4*>(src);
        r1[u] = *reinterpret_cast<const float4*>(src + W);
        pm[u] = *reinterpret_cast<const float2*>(pooled + plane * wins + 2 * (int64_t)wi);
        pg[u] = *reinterpret_cast<const float2*>(dpooled + plane * wins + 2 * (int64_t)wi);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float xs[4] = {half ? r0[u].z : r0[u].x, half ? r0[u].w : r0[u].y, half ? r1[u].z : r1[u].x,
                             half ? r1[u].w : r1[u].y};
        const float pmax = half ? pm[u].y : pm[u].x, pgrad = half ? pg[u].y : pg[u].x;
        float z[4], y[4];
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          z[k] = bn_affine(xs[k], mu, za, zb);
          y[k] = act ? fmaxf(z[k], 0.f) : z[k];
          cnt += (y[k] == pmax);
        }
        const float share = cnt ? pgrad / (float)cnt : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float g = (y[k] == pmax) ? share : 0.f;
          if (act && !(z[k] >= 0.f)) g = 0.f;
          const float d = xs[k] - mu;
          sa += g;
          sb += g * d;
          ea = fmaxf(ea, fabsf(g));
          eb = fmaxf(eb, fabsf(d));
        }
      }
    }
  }
  const float ta = block_sum(sa, scratch);
  const float tb = block_sum(sb, scratch);
  if (threadIdx.x == 0) {
    partial[(s * C + c) * 4 + 0] = ta;
    partial[(s * C + c) * 4 + 1] = tb;
  }
  const float ma = block_max(ea, scratch);
  const float mb = block_max(eb, scratch);
  if (threadIdx.x == 0) {
    partial[(s * C + c) * 4 + 2] = ma;
    partial[(s * C + c) * 4 + 3] = mb;
    __threadfence();
    const unsigned int prev = atomicAdd(fin.counter + c, 1u);
    if (prev == gridDim.y - 1) {
      __threadfence();
      bn_bwd_fold(partial, c, C, fin);
      fin.counter[c] = 0u;
    }
  }
}

// one thread per channel: fold the split partials, produce mean / invstd, update running stats
__global__ void __launch_bounds__(256) bn_fwd_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ x,
                                                              int64_t C, int64_t P, const BnFin fin) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) bn_fwd_fold(partial, x, c, C, P, fin);
}

// ---- synchronised BatchNorm (data parallel, ng_nccl_sync_bn) ----------------------------------------
// The reference computes the statistics over the whole batch (module.py:356-391); under data parallelism
// the batch is sharded, so every rank folds its partials into (mean_r, M2_r = sum (x - mean_r)^2, n_r),
// the triples are all-gathered ([2C + 4] floats per rank) and combined with Chan's pairwise update in
// rank order on every rank — the same arithmetic everywhere, hence bit-identical replicas, and no
// E[x^2] - E[x]^2 cancellation across ranks.
__global__ void __launch_bounds__(256) bn_fwd_local_kernel(const float* __restrict__ partial, const float* __restrict__ x,
                                                           float* __restrict__ send, int64_t C, int64_t P, int S,
                                                           float count) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0) send[2 * C] = count;
  if (c >= C) return;
  float sa = 0.f, sb = 0.f;
  for (int s = 0; s < S; ++s) {
    sa += partial[(s * C + c) * 2 + 0];
    sb += partial[(s * C + c) * 2 + 1];
  }
  const float d = sa / count;
  float m2 = sb - sa * d;
  if (m2 < 0.f) m2 = 0.f;
  send[2 * c + 0] = x[c * P] + d;
  send[2 * c + 1] = m2;
}

__global__ void __launch_bounds__(256) bn_fwd_finalize_sync_kernel(const float* __restrict__ gathered, int world,
                                                                   int64_t stride, float* __restrict__ save_mean,
                                                                   float* __restrict__ save_invstd, float* running_mean,
                                                                   float* running_var, int64_t C, float eps,
                                                                   float momentum) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int r = 0; r < world; ++r) {
    const float* g = gathered + (int64_t)r * stride;
    const float nr = g[2 * C], mr = g[2 * c], m2r = g[2 * c + 1];
    const float tot = n + nr;
    const float delta = mr - mean;
    mean += delta * (nr / tot);
    m2 += m2r + delta * delta * (n * nr / tot);
    n = tot;
  }
  save_mean[c] = mean;
  save_invstd[c] = 1.f / sqrtf(m2 / n + eps);
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (m2 / (n - 1.f));
  }
}

// y = (x - mean[c]) * gamma[c] * invstd[c] + beta[c]
// use_var: invstd is computed on the fly from a variance array (evaluation mode)
__global__ void __launch_bounds__(256) bn_apply_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ mean, const float* __restrict__ inv_or_var,
                                                       int64_t total, int64_t C, int64_t P, float eps, int use_var,
                                                       int vec_ok, int relu, FastDiv fp4 = FastDiv{0, 0, 0},
                                                       FastDiv fc = FastDiv{0, 0, 0}) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok) {
    const int64_t n4 = total >> 2;
    const bool fast = fp4.d != 0 && fc.d != 0 && n4 < (1LL << 31);
    for (int64_t i = tid; i < n4; i += nth) {
      int64_t c;
      if (fast) {
        const uint32_t plane = fastdiv((uint32_t)i, fp4);
        c = plane - fastdiv(plane, fc) * fc.d;
      } else {
        c = ((i << 2) / P) % C;
      }
      const float inv = use_var ? 1.f / sqrtf(inv_or_var[c] + eps) : inv_or_var[c];
      const float a = gamma[c] * inv, mu = mean[c], b = beta[c];
      float4 v = reinterpret_cast<const float4*>(x)[i];
      v.x = bn_affine(v.x, mu, a, b); v.y = bn_affine(v.y, mu, a, b);
      v.z = bn_affine(v.z, mu, a, b); v.w = bn_affine(v.w, mu, a, b);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      reinterpret_cast<float4*>(y)[i] = v;
    }
  } else {
    for (int64_t i = tid; i < total; i += nth) {
      const int64_t c = (i / P) % C;
      const float inv = use_var ? 1.f / sqrtf(inv_or_var[c] + eps) : inv_or_var[c];
      const float z = bn_affine(x[i], mean[c], gamma[c] * inv, beta[c]);
      y[i] = relu ? fmaxf(z, 0.f) : z;
    }
  }
}

__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ partial, int64_t C,
                                                              const BnFin fin) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) bn_bwd_fold(partial, c, C, fin);
}

// dx = gamma*invstd * (dy - sum_dy/n - (x-mean) * invstd^2 * sum_dy_xmu / n)
__global__ void __launch_bounds__(256) bn_bwd_dx_kernel(float* __restrict__ dx, const float* __restrict__ dy,
                                                        const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ mean, const float* __restrict__ invstd,
                                                        const float* __restrict__ sums, int64_t total, int64_t C, int64_t P,
                                                        float inv_count, int vec_ok, const float* __restrict__ beta,
                                                        FastDiv fp4 = FastDiv{0, 0, 0}, FastDiv fc = FastDiv{0, 0, 0},
                                                        int global_count = 0) {
  // beta != nullptr: fused ReLU, dy is masked by (z >= 0) with z recomputed from x
  // global_count: the element count behind `sums` is sums[2C] (allreduced with them), not the local one
  if (global_count) inv_count = 1.f / sums[2 * C];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok) {
    const int64_t n4 = total >> 2;
    const bool fast = fp4.d != 0 && fc.d != 0 && n4 < (1LL << 31);
    for (int64_t i = tid; i < n4; i += nth) {
      int64_t c;
      if (fast) {
        const uint32_t plane = fastdiv((uint32_t)i, fp4);
        c = plane - fastdiv(plane, fc) * fc.d;
      } else {
        c = ((i << 2) / P) % C;
      }
      const float inv = invstd[c], mu = mean[c];
      const float k0 = gamma[c] * inv, k1 = sums[c * 2] * inv_count, k2 = inv * inv * sums[c * 2 + 1] * inv_count;
      float4 g = reinterpret_cast<const float4*>(dy)[i];
      const float4 xv = reinterpret_cast<const float4*>(x)[i];
      if (beta) {
        const float zb = beta[c];
        if (!(bn_affine(xv.x, mu, k0, zb) >= 0.f)) g.x = 0.f;
        if (!(bn_affine(xv.y, mu, k0, zb) >= 0.f)) g.y = 0.f;
        if (!(bn_affine(xv.z, mu, k0, zb) >= 0.f)) g.z = 0.f;
        if (!(bn_affine(xv.w, mu, k0, zb) >= 0.f)) g.w = 0.f;
      }
      float4 o;
      o.x = k0 * (g.x - k1 - (xv.x - mu) * k2); o.y = k0 * (g.y - k1 - (xv.y - mu) * k2);
      o.z = k0 * (g.z - k1 - (xv.z - mu) * k2); o.w = k0 * (g.w - k1 - (xv.w - mu) * k2);
      reinterpret_cast<float4*>(dx)[i] = o;
    }
  } else {
    for (int64_t i = tid; i < total; i += nth) {
      const int64_t c = (i / P) % C;
      const float inv = invstd[c];
      const float k0 = gamma[c] * inv, k1 = sums[c * 2] * inv_count, k2 = inv * inv * sums[c * 2 + 1] * inv_count;
      float g = dy[i];
      if (beta && !(bn_affine(x[i], mean[c], k0, beta[c]) >= 0.f)) g = 0.f;
      dx[i] = k0 * (g - k1 - (x[i] - mean[c]) * k2);
    }
  }
}

// dx pass that also returns the per-channel sums of dx (the bias gradient of the convolution that produced this
// layer's input, Add.backward's unbroadcast, ops_cpu.py:12-18): grid (C, S) like the reduction passes, a block streams
// its channel's planes of n_per images, writes dx and keeps a running sum; bn_plane_fold_kernel adds the S partials
// in a fixed order.  Used when the convolution below does not stage its operands (first layer): saves the separate
// pass over dx that the channel sum would take.
__global__ void __launch_bounds__(256) bn_bwd_dx_sum_kernel(float* __restrict__ dx, float* __restrict__ partial,
                                                            const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, const float* __restrict__ sums,
                                                            int64_t N, int64_t C, int64_t P, int64_t n_per_split,
                                                            float inv_count, int vec_ok, const float* __restrict__ beta,
                                                            FastDiv fp4) {
  __shared__ float scratch[32];
  const int64_t c = blockIdx.x, s = blockIdx.y;
  const float inv = invstd[c], mu = mean[c];
  const float k0 = gamma[c] * inv, k1 = sums[c * 2] * inv_count, k2 = inv * inv * sums[c * 2 + 1] * inv_count;
  const float zb = beta ? beta[c] : 0.f;
  const int64_t n0 = s * n_per_split;
  int64_t n1 = n0 + n_per_split;
  if (n1 > N) n1 = N;
  float acc = 0.f;
  if (vec_ok) {
    const int64_t P4 = P >> 2;
    const int64_t items = (n1 - n0) * P4;
    const bool fast = fp4.d != 0 && items < (1LL << 31);
    constexpr int U = 2;
    for (int64_t i0 = threadIdx.x; i0 < items; i0 += U * (int64_t)blockDim.x) {
      float4 xv[U], gv[U];
      int64_t off[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * (int64_t)blockDim.x;
        off[u] = -1;
        if (i < items) {
          int64_t n, p4;
          if (fast) {
            const uint32_t q = fastdiv((uint32_t)i, fp4);
            n = n0 + q;
            p4 = (uint32_t)i - q * fp4.d;
          } else {
            n = n0 + i / P4;
            p4 = i % P4;
          }
          off[u] = (n * C + c) * P + (p4 << 2);
          xv[u] = *reinterpret_cast<const float4*>(x + off[u]);
          gv[u] = *reinterpret_cast<const float4*>(dy + off[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (off[u] < 0) continue;
        float4 g = gv[u];
        if (beta) {
          if (!(bn_affine(xv[u].x, mu, k0, zb) >= 0.f)) g.x = 0.f;
          if (!(bn_affine(xv[u].y, mu, k0, zb) >= 0.f)) g.y = 0.f;
          if (!(bn_affine(xv[u].z, mu, k0, zb) >= 0.f)) g.z = 0.f;
          if (!(bn_affine(xv[u].w, mu, k0, zb) >= 0.f)) g.w = 0.f;
        }
        float4 o;
        o.x = k0 * (g.x - k1 - (xv[u].x - mu) * k2); o.y = k0 * (g.y - k1 - (xv[u].y - mu) * k2);
        o.z = k0 * (g.z - k1 - (xv[u].z - mu) * k2); o.w = k0 * (g.w - k1 - (xv[u].w - mu) * k2);
        *reinterpret_cast<float4*>(dx + off[u]) = o;
        acc += (o.x + o.y) + (o.z + o.w);
      }
    }
  } else {
    const int64_t items = (n1 - n0) * P;
    for (int64_t i = threadIdx.x; i < items; i += blockDim.x) {
      const int64_t n = n0 + i / P, pp = i % P;
      const int64_t off = (n * C + c) * P + pp;
      float g = dy[off];
      if (beta && !(bn_affine(x[off], mu, k0, zb) >= 0.f)) g = 0.f;
      const float o = k0 * (g - k1 - (x[off] - mu) * k2);
      dx[off] = o;
      acc += o;
    }
  }
  const float t = block_sum(acc, scratch);
  if (threadIdx.x == 0) partial[s * C + c] = t;
}

// out[c] = sum_s partial[s][c]  (fixed order; one warp per channel)
__global__ void __launch_bounds__(256) bn_plane_fold_kernel(float* __restrict__ out, const float* __restrict__ partial,
                                                            int64_t S, int64_t C) {
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int64_t n = lane; n < S; n += 32) acc += partial[n * C + c];
  acc = warp_sum(acc);
  if (lane == 0) out[c] = acc;
}

// Splits of the batch axis for the reduction passes (grid = C x S blocks, about four blocks per SM).  A search for
// the split count with the smallest last-wave loss (e.g. 64 channels x 16 splits instead of x 10) was measured and is
// no better forward and 20 % slower backward (53 vs 44 us at 64 channels x 512 x 32^2): fewer, longer blocks win.
static int bn_split(int64_t N, int64_t C, int64_t P, int64_t* n_per) {
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  int64_t S = ceil_div((int64_t)sms * 4, C);
  if (S > N) S = N;
  while (S > 1 && (N / S) * P < 8192) S--;
  *n_per = ceil_div(N, S);
  return (int)ceil_div(N, *n_per);
}

static inline int bgrid(int64_t n) {
  int64_t g = ceil_div(n, 256);
  const int64_t cap = (int64_t)(g_sm_count > 0 ? g_sm_count : 148) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace ng

using namespace ng;

extern "C" {

// Statistics half of the training forward: mean / invstd (and the running statistics) of x.  With absmax_bits
// (single process only) the statistics pass also tracks the channel extremes and *absmax_bits receives the bits of
// max |y| of the normalised (and activated) output before any of it is written.
// [C] arrival counters of the fused fold: zero-initialised once, re-armed by the folding block, used by one
// launch at a time (single in-order compute stream)
static unsigned int* fold_counters() {
  static unsigned int* buf = nullptr;
  if (!buf) {
    if (cudaMalloc((void**)&buf, 65536 * sizeof(unsigned int)) != cudaSuccess) return nullptr;
    cudaMemsetAsync(buf, 0, 65536 * sizeof(unsigned int), g_stream);
  }
  return buf;
}

static int bn_fwd_stats(uint64_t x, uint64_t gamma, uint64_t beta, uint64_t running_mean, uint64_t running_var,
                        uint64_t save_mean, uint64_t save_invstd, uint64_t absmax_bits, int64_t N, int64_t C, int64_t HW,
                        float eps, float momentum, int act, int vec, uint64_t snapshot = 0) {
  int64_t n_per = 0;
  const int S = bn_split(N, C, HW, &n_per);
  const bool extra = absmax_bits != 0;
  const bool sync = dp_sync_bn();
  NG_REQUIRE(!(extra && sync), "ng_bn2d_act_fwd_stats: not available under synchronised BatchNorm");
  float* partial = nullptr;
  if (pool_alloc((void**)&partial, (uint64_t)S * C * (extra ? 4 : 2) * sizeof(float))) return 1;
  BnFin fin{};
  fin.S = S; fin.PS = extra ? 4 : 2; fin.act = act;
  fin.count = (float)(N * HW); fin.eps = eps; fin.momentum = momentum;
  fin.gamma = dptr<const float>(gamma); fin.beta = dptr<const float>(beta);
  fin.bits = extra ? dptr<unsigned int>(absmax_bits) : nullptr;
  fin.save_mean = dptr<float>(save_mean); fin.save_invstd = dptr<float>(save_invstd);
  fin.snap_gamma = snapshot ? dptr<float>(snapshot) : nullptr;
  fin.snap_beta = snapshot ? dptr<float>(snapshot) + C : nullptr;
  fin.running_mean = running_mean ? dptr<float>(running_mean) : nullptr;
  fin.running_var = running_var ? dptr<float>(running_var) : nullptr;
  fin.counter = sync ? nullptr : fold_counters();   // single process: the last block of a channel folds it
  if (!sync && !fin.counter) return set_error("bn_fwd_stats: cudaMalloc of the fold counters failed");
  const dim3 grid((unsigned)C, (unsigned)S);
  const FastDiv fd = make_fastdiv(vec ? HW / 4 : 0);
  if (extra) {
    NG_CHECK_CUDA(cudaMemsetAsync(dptr<void>(absmax_bits), 0, 4, g_stream));
    NG_LAUNCH((bn_reduce_kernel<0, true>), grid, 256, 0, partial, dptr<const float>(x), (const float*)nullptr,
              (const float*)nullptr, N, C, HW, n_per, vec, (const float*)nullptr, (const float*)nullptr,
              (const float*)nullptr, fd, fin);
  } else {
    NG_LAUNCH((bn_reduce_kernel<0>), grid, 256, 0, partial, dptr<const float>(x), (const float*)nullptr,
              (const float*)nullptr, N, C, HW, n_per, vec, (const float*)nullptr, (const float*)nullptr,
              (const float*)nullptr, fd, fin);
  }
  NG_CHECK_LAUNCH();
  if (sync) {
    const int world = dp_world();
    const int64_t stride = 2 * C + 4;
    float* xchg = nullptr;  // [stride] to send, then [world][stride] gathered
    if (pool_alloc((void**)&xchg, (uint64_t)(world + 1) * stride * sizeof(float))) return 1;
    NG_LAUNCH(bn_fwd_local_kernel, (int)ceil_div(C, 256), 256, 0, (const float*)partial, dptr<const float>(x), xchg, C,
              HW, S, (float)(N * HW));
    NG_CHECK_LAUNCH();
    if (dp_allgather_inline(xchg, xchg + stride, stride)) return 1;
    NG_LAUNCH(bn_fwd_finalize_sync_kernel, (int)ceil_div(C, 256), 256, 0, (const float*)(xchg + stride), world, stride,
              dptr<float>(save_mean), dptr<float>(save_invstd), running_mean ? dptr<float>(running_mean) : nullptr,
              running_var ? dptr<float>(running_var) : nullptr, C, eps, momentum);
    NG_CHECK_LAUNCH();
    pool_free(xchg);
  }
  pool_free(partial);
  return 0;
}

static int bn_apply(uint64_t y, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t mean, uint64_t invstd, int64_t N,
                    int64_t C, int64_t HW, int act, int vec) {
  const int64_t total = N * C * HW;
  NG_LAUNCH(bn_apply_kernel, bgrid(vec ? total / 4 : total), 256, 0, dptr<float>(y), dptr<const float>(x),
            dptr<const float>(gamma), dptr<const float>(beta), dptr<const float>(mean), dptr<const float>(invstd), total, C,
            HW, 0.f, 0, vec, act, make_fastdiv(vec ? HW / 4 : 0), make_fastdiv(C));
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_bn2d_act_fwd_train(uint64_t y, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t running_mean,
                          uint64_t running_var, uint64_t save_mean, uint64_t save_invstd, int64_t N, int64_t C,
                          int64_t HW, float eps, float momentum, int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_BN, 0.0, 12.0 * (double)N * C * HW);
  NG_REQUIRE(N > 0 && C > 0 && HW > 0, "ng_bn2d_fwd_train: empty input");
  NG_REQUIRE(C <= 65535, "ng_bn2d_fwd_train: too many channels");
  NG_REQUIRE(act == 0 || act == 1, "ng_bn2d_act_fwd_train: act must be 0 (none) or 1 (ReLU)");
  const int vec = (HW % 4 == 0) && ((x & 15u) == 0) && ((y & 15u) == 0);
  if (bn_fwd_stats(x, gamma, beta, running_mean, running_var, save_mean, save_invstd, 0, N, C, HW, eps, momentum, act, vec))
    return 1;
  return bn_apply(y, x, gamma, beta, save_mean, save_invstd, N, C, HW, act, vec);
}

int ng_bn2d_act_fwd_stats(uint64_t x, uint64_t gamma, uint64_t beta, uint64_t running_mean, uint64_t running_var,
                          uint64_t save_mean, uint64_t save_invstd, uint64_t absmax_bits, uint64_t snapshot, int64_t N,
                          int64_t C, int64_t HW, float eps, float momentum, int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_BN, 0.0, 4.0 * (double)N * C * HW);
  NG_REQUIRE(N > 0 && C > 0 && HW > 0, "ng_bn2d_act_fwd_stats: empty input");
  NG_REQUIRE(C <= 65535, "ng_bn2d_act_fwd_stats: too many channels");
  NG_REQUIRE(act == 0 || act == 1, "ng_bn2d_act_fwd_stats: act must be 0 (none) or 1 (ReLU)");
  NG_REQUIRE(absmax_bits != 0, "ng_bn2d_act_fwd_stats: absmax_bits is required");
  const int vec = (HW % 4 == 0) && ((x & 15u) == 0);
  return bn_fwd_stats(x, gamma, beta, running_mean, running_var, save_mean, save_invstd, absmax_bits, N, C, HW, eps,
                      momentum, act, vec, snapshot);
}

int ng_bn2d_act_apply(uint64_t y, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t save_mean, uint64_t save_invstd,
                      int64_t N, int64_t C, int64_t HW, int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_BN, 0.0, 8.0 * (double)N * C * HW);
  NG_REQUIRE(N > 0 && C > 0 && HW > 0, "ng_bn2d_act_apply: empty input");
  NG_REQUIRE(act == 0 || act == 1, "ng_bn2d_act_apply: act must be 0 (none) or 1 (ReLU)");
  const int vec = (HW % 4 == 0) && ((x & 15u) == 0) && ((y & 15u) == 0);
  return bn_apply(y, x, gamma, beta, save_mean, save_invstd, N, C, HW, act, vec);
}

int ng_bn2d_fwd_train(uint64_t y, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t running_mean,
                      uint64_t running_var, uint64_t save_mean, uint64_t save_invstd, int64_t N, int64_t C,
                      int64_t HW, float eps, float momentum) {
  return ng_bn2d_act_fwd_train(y, x, gamma, beta, running_mean, running_var, save_mean, save_invstd, N, C, HW, eps,
                               momentum, 0);
}

int ng_bn2d_act_fwd_eval(uint64_t y, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t running_mean,
                         uint64_t running_var, int64_t N, int64_t C, int64_t HW, float eps, int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_BN, 0.0, 8.0 * (double)N * C * HW);
  NG_REQUIRE(act == 0 || act == 1, "ng_bn2d_act_fwd_eval: act must be 0 (none) or 1 (ReLU)");
  const int64_t total = N * C * HW;
  if (total <= 0) return 0;
  const int vec = (HW % 4 == 0) && ((x & 15u) == 0) && ((y & 15u) == 0);
  NG_LAUNCH(bn_apply_kernel, bgrid(vec ? total / 4 : total), 256, 0, dptr<float>(y), dptr<const float>(x),
            dptr<const float>(gamma), dptr<const float>(beta), dptr<const float>(running_mean),
            dptr<const float>(running_var), total, C, HW, eps, 1, vec, act, make_fastdiv(vec ? HW / 4 : 0),
            make_fastdiv(C));
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_bn2d_fwd_eval(uint64_t y, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t running_mean,
                     uint64_t running_var, int64_t N, int64_t C, int64_t HW, float eps) {
  return ng_bn2d_act_fwd_eval(y, x, gamma, beta, running_mean, running_var, N, C, HW, eps, 0);
}

// Reduction half of the backward: dgamma, dbeta and sums[2C + 1] = {sum dy, sum dy (x - mean)} per channel and
// the element count behind them (all-reduced under synchronised BatchNorm).  With bound_bits (single process
// only) *bound_bits receives the bits of an upper bound of max |dx|.
static int bn_bwd_sums(float* sums, uint64_t dgamma, uint64_t dbeta, uint64_t bound_bits, uint64_t dy, uint64_t x,
                       uint64_t gamma, uint64_t beta, uint64_t save_mean, uint64_t save_invstd, int64_t N, int64_t C,
                       int64_t HW, int act, int vec) {
  int64_t n_per = 0;
  const int S = bn_split(N, C, HW, &n_per);
  const bool extra = bound_bits != 0;
  const bool sync = dp_sync_bn();
  NG_REQUIRE(!(extra && sync), "ng_bn2d_act_bwd_sums: not available under synchronised BatchNorm");
  float* partial = nullptr;
  if (pool_alloc((void**)&partial, (uint64_t)S * C * (extra ? 4 : 2) * sizeof(float))) return 1;
  const float* g = act ? dptr<const float>(gamma) : nullptr;
  const float* b = act ? dptr<const float>(beta) : nullptr;
  const float* iv = act ? dptr<const float>(save_invstd) : nullptr;
  BnFin fin{};
  fin.S = S; fin.PS = extra ? 4 : 2; fin.act = act;
  fin.count = (float)(N * HW);
  fin.gamma = dptr<const float>(gamma);
  fin.bits = extra ? dptr<unsigned int>(bound_bits) : nullptr;
  fin.invstd = dptr<const float>(save_invstd);
  fin.dgamma = dptr<float>(dgamma); fin.dbeta = dptr<float>(dbeta); fin.sums = sums;
  // the fold rides on the reduction launch (last block of a channel); synchronised BatchNorm folds separately so
  // that the sums exist as one block before the all-reduce
  fin.counter = sync ? nullptr : fold_counters();
  if (!sync && !fin.counter) return set_error("bn_bwd_sums: cudaMalloc of the fold counters failed");
  const dim3 grid((unsigned)C, (unsigned)S);
  const FastDiv fd = make_fastdiv(vec ? HW / 4 : 0);
  if (extra) NG_CHECK_CUDA(cudaMemsetAsync(dptr<void>(bound_bits), 0, 4, g_stream));
  if (act && extra)
    NG_LAUNCH((bn_reduce_kernel<2, true>), grid, 256, 0, partial, dptr<const float>(x), dptr<const float>(dy),
              dptr<const float>(save_mean), N, C, HW, n_per, vec, g, b, iv, fd, fin);
  else if (act)
    NG_LAUNCH((bn_reduce_kernel<2>), grid, 256, 0, partial, dptr<const float>(x), dptr<const float>(dy),
              dptr<const float>(save_mean), N, C, HW, n_per, vec, g, b, iv, fd, fin);
  else if (extra)
    NG_LAUNCH((bn_reduce_kernel<1, true>), grid, 256, 0, partial, dptr<const float>(x), dptr<const float>(dy),
              dptr<const float>(save_mean), N, C, HW, n_per, vec, g, b, iv, fd, fin);
  else
    NG_LAUNCH((bn_reduce_kernel<1>), grid, 256, 0, partial, dptr<const float>(x), dptr<const float>(dy),
              dptr<const float>(save_mean), N, C, HW, n_per, vec, g, b, iv, fd, fin);
  NG_CHECK_LAUNCH();
  if (sync) {
    NG_LAUNCH(bn_bwd_finalize_kernel, (int)ceil_div(C, 256), 256, 0, (const float*)partial, C, fin);
    NG_CHECK_LAUNCH();
  }
  pool_free(partial);
  return 0;
}

static int bn_bwd_dx(uint64_t dx, uint64_t dy, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t save_mean,
                     uint64_t save_invstd, const float* sums, int64_t N, int64_t C, int64_t HW, int act, int vec, int sync) {
  const int64_t total = N * C * HW;
  NG_LAUNCH(bn_bwd_dx_kernel, bgrid(vec ? total / 4 : total), 256, 0, dptr<float>(dx), dptr<const float>(dy),
            dptr<const float>(x), dptr<const float>(gamma), dptr<const float>(save_mean),
            dptr<const float>(save_invstd), sums, total, C, HW, 1.f / (float)(N * HW), vec,
            act ? dptr<const float>(beta) : (const float*)nullptr, make_fastdiv(vec ? HW / 4 : 0), make_fastdiv(C),
            sync);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_bn2d_act_bwd(uint64_t dx, uint64_t dgamma, uint64_t dbeta, uint64_t dy, uint64_t x, uint64_t gamma,
                    uint64_t beta, uint64_t save_mean, uint64_t save_invstd, int64_t N, int64_t C, int64_t HW,
                    int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_BN, 0.0, 20.0 * (double)N * C * HW);
  NG_REQUIRE(N > 0 && C > 0 && HW > 0, "ng_bn2d_bwd: empty input");
  NG_REQUIRE(C <= 65535, "ng_bn2d_bwd: too many channels");
  NG_REQUIRE(act == 0 || (act == 1 && beta != 0), "ng_bn2d_act_bwd: act must be 0, or 1 (ReLU) with beta given");
  float* sums = nullptr;  // [C][2] sums, then the element count behind them
  if (pool_alloc((void**)&sums, ((uint64_t)C * 2 + 4) * sizeof(float))) return 1;
  const int vec = (HW % 4 == 0) && ((x & 15u) == 0) && ((dy & 15u) == 0) && ((dx & 15u) == 0);
  if (bn_bwd_sums(sums, dgamma, dbeta, 0, dy, x, gamma, beta, save_mean, save_invstd, N, C, HW, act, vec)) return 1;
  // synchronised BatchNorm: dgamma / dbeta stay this rank's sums (the gradient bucket adds them across
  // ranks); dx needs the two sums, and the count, over the global batch
  const int sync = dp_sync_bn() ? 1 : 0;
  if (sync && dp_allreduce_inline(sums, 2 * C + 1)) return 1;
  const int rc = bn_bwd_dx(dx, dy, x, gamma, beta, save_mean, save_invstd, sums, N, C, HW, act, vec, sync);
  pool_free(sums);
  return rc;
}

int ng_bn2d_act_bwd_sums(uint64_t dgamma, uint64_t dbeta, uint64_t sums, uint64_t bound_bits, uint64_t dy, uint64_t x,
                         uint64_t gamma, uint64_t beta, uint64_t save_mean, uint64_t save_invstd, int64_t N, int64_t C,
                         int64_t HW, int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_BN, 0.0, 8.0 * (double)N * C * HW);
  NG_REQUIRE(N > 0 && C > 0 && HW > 0, "ng_bn2d_act_bwd_sums: empty input");
  NG_REQUIRE(C <= 65535, "ng_bn2d_act_bwd_sums: too many channels");
  NG_REQUIRE(act == 0 || (act == 1 && beta != 0), "ng_bn2d_act_bwd_sums: act must be 0, or 1 (ReLU) with beta given");
  NG_REQUIRE(sums != 0 && bound_bits != 0, "ng_bn2d_act_bwd_sums: sums and bound_bits are required");
  const int vec = (HW % 4 == 0) && ((x & 15u) == 0) && ((dy & 15u) == 0);
  return bn_bwd_sums(dptr<float>(sums), dgamma, dbeta, bound_bits, dy, x, gamma, beta, save_mean, save_invstd, N, C, HW,
                     act, vec);
}

int ng_bn2d_act_bwd_sums_pool(uint64_t dgamma, uint64_t dbeta, uint64_t sums, uint64_t bound_bits, uint64_t dpooled,
                              uint64_t pooled, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t save_mean,
                              uint64_t save_invstd, int64_t N, int64_t C, int H, int W, int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_BN, 0.0, 6.0 * (double)N * C * H * W);   // x once, pooled and its gradient (a quarter each)
  NG_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 4 == 0 && (x & 15u) == 0 && (pooled & 7u) == 0 &&
                 (dpooled & 7u) == 0,
             "ng_bn2d_act_bwd_sums_pool: needs even H, W %% 4 == 0, a 16-byte aligned input and 8-byte aligned pooled tensors");
  NG_REQUIRE(C <= 65535 && (act == 0 || act == 1), "ng_bn2d_act_bwd_sums_pool: bad channel count or act");
  NG_REQUIRE(sums != 0 && bound_bits != 0, "ng_bn2d_act_bwd_sums_pool: sums and bound_bits are required");
  NG_REQUIRE(!dp_sync_bn(), "ng_bn2d_act_bwd_sums_pool: not available under synchronised BatchNorm");
  const int64_t HW = (int64_t)H * W;
  NG_REQUIRE(N * (HW / 4) < (1LL << 31), "ng_bn2d_act_bwd_sums_pool: tensor too large");
  int64_t n_per = 0;
  const int S = bn_split(N, C, HW, &n_per);
  float* partial = nullptr;
  if (pool_alloc((void**)&partial, (uint64_t)S * C * 4 * sizeof(float))) return 1;
  BnFin fin{};
  fin.S = S; fin.PS = 4; fin.act = act;
  fin.count = (float)(N * HW);
  fin.gamma = dptr<const float>(gamma);
  fin.bits = dptr<unsigned int>(bound_bits);
  fin.invstd = dptr<const float>(save_invstd);
  fin.dgamma = dptr<float>(dgamma); fin.dbeta = dptr<float>(dbeta); fin.sums = dptr<float>(sums);
  fin.counter = fold_counters();
  if (!fin.counter) return set_error("ng_bn2d_act_bwd_sums_pool: cudaMalloc of the fold counters failed");
  NG_CHECK_CUDA(cudaMemsetAsync(dptr<void>(bound_bits), 0, 4, g_stream));
  NG_LAUNCH(bn_reduce_pool_kernel, dim3((unsigned)C, (unsigned)S), 256, 0, partial, dptr<const float>(x),
            dptr<const float>(dpooled), dptr<const float>(pooled), dptr<const float>(save_mean), dptr<const float>(gamma),
            dptr<const float>(beta), dptr<const float>(save_invstd), N, C, H, W, n_per, act, make_fastdiv(HW / 8),
            make_fastdiv(W / 4), fin);
  NG_CHECK_LAUNCH();
  pool_free(partial);
  return 0;
}

int ng_bn2d_act_bwd_dx(uint64_t dx, uint64_t dy, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t save_mean,
                       uint64_t save_invstd, uint64_t sums, int64_t N, int64_t C, int64_t HW, int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_BN, 0.0, 12.0 * (double)N * C * HW);
  NG_REQUIRE(N > 0 && C > 0 && HW > 0, "ng_bn2d_act_bwd_dx: empty input");
  NG_REQUIRE(act == 0 || (act == 1 && beta != 0), "ng_bn2d_act_bwd_dx: act must be 0, or 1 (ReLU) with beta given");
  const int vec = (HW % 4 == 0) && ((x & 15u) == 0) && ((dy & 15u) == 0) && ((dx & 15u) == 0);
  return bn_bwd_dx(dx, dy, x, gamma, beta, save_mean, save_invstd, dptr<const float>(sums), N, C, HW, act, vec, 0);
}

int ng_bn2d_act_bwd_dx_sum(uint64_t dx, uint64_t chan_sums, uint64_t dy, uint64_t x, uint64_t gamma, uint64_t beta,
                           uint64_t save_mean, uint64_t save_invstd, uint64_t sums, int64_t N, int64_t C, int64_t HW, int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_BN, 0.0, 12.0 * (double)N * C * HW);
  NG_REQUIRE(N > 0 && C > 0 && HW > 0 && N * C < 2147483647LL, "ng_bn2d_act_bwd_dx_sum: bad dimensions");
  NG_REQUIRE(act == 0 || (act == 1 && beta != 0), "ng_bn2d_act_bwd_dx_sum: act must be 0, or 1 (ReLU) with beta given");
  NG_REQUIRE(chan_sums != 0 && sums != 0, "ng_bn2d_act_bwd_dx_sum: chan_sums and sums are required");
  NG_REQUIRE(C <= 65535, "ng_bn2d_act_bwd_dx_sum: too many channels");
  int64_t n_per = 0;
  const int S = bn_split(N, C, HW, &n_per);
  float* partial = nullptr;
  if (pool_alloc((void**)&partial, (uint64_t)S * C * sizeof(float))) return 1;
  const int vec = (HW % 4 == 0) && ((x & 15u) == 0) && ((dy & 15u) == 0) && ((dx & 15u) == 0);
  NG_LAUNCH(bn_bwd_dx_sum_kernel, dim3((unsigned)C, (unsigned)S), 256, 0, dptr<float>(dx), partial, dptr<const float>(dy),
            dptr<const float>(x), dptr<const float>(gamma), dptr<const float>(save_mean), dptr<const float>(save_invstd),
            dptr<const float>(sums), N, C, HW, n_per, 1.f / (float)(N * HW), vec,
            act ? dptr<const float>(beta) : (const float*)nullptr, make_fastdiv(vec ? HW / 4 : 0));
  NG_CHECK_LAUNCH();
  NG_LAUNCH(bn_plane_fold_kernel, (int)ceil_div(C, 8), 256, 0, dptr<float>(chan_sums), (const float*)partial, (int64_t)S, C);
  NG_CHECK_LAUNCH();
  pool_free(partial);
  return 0;
}

int ng_bn2d_bwd(uint64_t dx, uint64_t dgamma, uint64_t dbeta, uint64_t dy, uint64_t x, uint64_t gamma,
                uint64_t save_mean, uint64_t save_invstd, int64_t N, int64_t C, int64_t HW) {
  return ng_bn2d_act_bwd(dx, dgamma, dbeta, dy, x, gamma, 0, save_mean, save_invstd, N, C, HW, 0);
}

}  // extern "C"
