// FP16x3: FP32-accurate convolution on tcgen05 tensor cores with FP16 operand pairs (sm_100a).
//
//   v * 2^s = hi + lo,  hi = fp16(v * 2^s),  lo = fp16(v * 2^s - hi)          (11 + 11 significant bits)
//   a * b  ~= (a_hi*b_hi + a_lo*b_hi) + a_hi*b_lo                              (dropped: a_lo*b_lo ~ 2^-22 |ab|)
//
// FP16 products are exact in the tensor core's FP32 accumulator, so this is the same error-compensated
// scheme as 3xTF32 (umma_conv.cu) — TF32 carries 11 significant bits as well — on kind::f16 MMAs, which
// run at twice the TF32 rate, and with the split done ONCE per tensor by the staging pass the tensor-core
// path needs anyway (NCHW -> channels-last), not per tile per stage inside the MMA kernel: no split warps,
// no TMEM operand ring, no lo rings.  FP16 has a 5-bit exponent, so each tensor is scaled by a power of two
// that puts its largest magnitude in [2^14, 2^15) (exact; undone in the epilogue): an element down to 2^-18
// of the maximum keeps a normal lo part, smaller ones lose relative (never absolute: <= 2^-40 max) accuracy.
//
// Staged layout ("pair-interleaved channels-last"), C % 64 == 0:
//     staged[n][pixel][g = c / 64][plane (0 = hi, 1 = lo)][c % 64]   FP16   (= the bytes of the FP32 tensor)
//     then 64 bytes of metadata: float inv_scale = 2^-s, uint32 bits of max|v|.
// One pixel's 64-channel group is a 128-byte row, so
//   * fprop / dgrad (K-major A operand, rows = pixels): ONE 5-D TMA box {64, w, rows, imgs, 2 planes} lands the
//     hi tile and the lo tile back to back in the 128B swizzle;
//   * wgrad (MN-major operands, rows = pixels = the reduction): ONE 5-D box per filter tap fetches every
//     (group, plane) block of the channel tile; the UMMA leading-dimension byte offset skips the other plane.
//
// The two big MMAs of a K = 16 step share one instruction where the layout allows it: B_hi and B_lo sit side by
// side as one N = 2n operand, D[:, 0:2n] = A_hi * [B_hi | B_lo], then D[:, 0:n] += A_lo * B_hi; the epilogue
// adds the two halves.  A is read twice instead of three times: 20 KB of shared-memory reads per step of a
// 128 x 128 tile in 192 tensor cycles (104 B/clk, under the 128 B/clk/SM shared-memory limit).
#include "umma_common.cuh"

#ifndef NG_EMU
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdlib.h>

namespace ng {
namespace umma {

bool f16x3_supported(int C, int K, int RS) { return C % 64 == 0 && K % 64 == 0 && RS <= 64; }

// ================================================================== staging
__global__ void __launch_bounds__(256) absmax_kernel(unsigned int* __restrict__ out_bits, const float* __restrict__ in,
                                                     long long n) {
  float m = 0.f;
  const long long n4 = n >> 2;
  const bool vec = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
  if (vec) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    if (blockIdx.x == 0)
      for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(in[i]));
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      m = fmaxf(m, fabsf(in[i]));
  }
  m = warp_max(m);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    if (m > 0.f && m < 3.0e38f) atomicMax(out_bits, __float_as_uint(m));   // positive floats order like their bits
  }
}

// scale = 2^(15 - e) with max = f * 2^e, f in [0.5, 1): the scaled maximum lies in [2^14, 2^15)
__device__ __forceinline__ float f16_scale(unsigned int max_bits) {
  const float mx = __uint_as_float(max_bits);
  if (!(mx > 0.f)) return 1.f;
  int e;
  frexpf(mx, &e);
  e = 15 - e;
  if (e > 126) e = 126;      // keep scale and 1/scale normal
  if (e < -126) e = -126;
  return ldexpf(1.f, e);
}
__device__ __forceinline__ void f16_split(float v, float scale, __half& hi, __half& lo) {
  const float sv = v * scale;
  hi = __float2half_rn(sv);
  lo = __float2half_rn(sv - __half2float(hi));
}

// Where the staged tensor comes from.  SRC 0: an FP32 NCHW tensor.  SRC 1: the output of a training-mode
// BatchNorm2d (+ ReLU) that was never written, y = act(bn_affine(x)) recomputed from the layer input x and its
// statistics.  SRC 2: the input gradient of such a layer that was never written,
// dx = k0 (dy (masked) - k1 - (x - mean) k2), recomputed from dy, x and the backward sums exactly as bn_bwd_dx_kernel
// (batchnorm.cu) evaluates it.  SRC 3: the same with dy itself recomputed from a fused 2x2 max-pool backward.
// Fusing the normalise passes into the staging pass removes one write and one read of the activation (forward) and
// of the gradient (backward) per layer.
struct StageSrc {
  const float* in;       // SRC 0: the tensor; SRC 1, 2: the BatchNorm input x
  const float* dy;       // SRC 2
  const float* gamma;
  const float* beta;     // null: no fused ReLU
  const float* mean;
  const float* invstd;
  const float* sums;     // SRC 2, 3: [C][2] = sum dy, sum dy (x - mean)
  float inv_count;       // SRC 2, 3
  int act;
  // SRC 3: like SRC 2, but the layer's output went through a 2x2 max pooling whose backward is fused in as well:
  // dy is never written, it is recomputed per window from the pooled maxima and their gradient (see
  // bn_reduce_pool_kernel in batchnorm.cu).  Tiles are th rows x tw columns (whole windows), tw * th == 32.
  const float* pooled;
  const float* dpooled;
  int H, W, tw;
};

// staged[n][p][g][plane][64] = split(value[n][g*64 + .][p]); one tile of 64 channels x 32 pixels per block,
// grid (pixel tiles, groups, images): many small blocks keep enough loads in flight (5.7 TB/s; a block walking all
// the tiles of an (image, group) to keep running sums was latency-bound at 2.3 TB/s).
//   SUM: the block also writes the channel sums of its tile, partial[(n * tiles + tile)][c] (folded in a fixed
//        order by channel_partial_sum: the bias gradient when the staged tensor is dy).
// One element of the staged tensor from its source values (xv: the tensor / the BatchNorm input, gv: dy).
template <int SRC>
__device__ __forceinline__ float stage_value(float xv, float gv, int act, const float (*cst)[64], int c) {
  if (SRC == 0) return xv;
  if (SRC == 1) {
    const float v = bn_affine(xv, cst[0][c], cst[1][c], cst[2][c]);
    return act ? fmaxf(v, 0.f) : v;
  }
  const float mu = cst[0][c], k0 = cst[1][c];
  if (act && !(bn_affine(xv, mu, k0, cst[4][c]) >= 0.f)) gv = 0.f;
  return k0 * (gv - cst[2][c] - (xv - mu) * cst[3][c]);
}

// VEC: HW % 4 == 0 and 16-byte aligned sources: the tile is read with 128-bit loads (a warp covers four channel
// rows of 128 bytes), twice the bytes in flight per load instruction of the scalar form.
template <int SRC, bool SUM, bool VEC>
__global__ void __launch_bounds__(256) stage_f16_kernel(__half* __restrict__ out, const StageSrc src,
                                                        float* __restrict__ partial, int C, int HW,
                                                        const unsigned int* __restrict__ max_bits,
                                                        float* __restrict__ inv_scale_out) {
  static_assert(SRC != 3 || VEC, "the pooled source exists in the vector form only");
  __shared__ float tile[64][33];
  __shared__ float cst[5][64];   // per-channel constants of the fused sources
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int G = C >> 6;
  const int g = blockIdx.y, n = blockIdx.z;
  const int c0 = g * 64;
  // pixel of local index l (0..31) of this block's tile: 32 consecutive pixels, or (SRC 3) th rows x tw columns
  int p0 = blockIdx.x * 32, h0 = 0, w0 = 0, tw = 32, tw_shift = 5;
  if (SRC == 3) {
    tw = src.tw;
    tw_shift = tw == 16 ? 4 : 3;
    const int tiles_w = src.W / tw;
    const int tyi = (int)blockIdx.x / tiles_w, txi = (int)blockIdx.x - tyi * tiles_w;
    h0 = tyi * (32 >> tw_shift);
    w0 = txi * tw;
    p0 = 0;
  }
  auto pix = [&](int l) -> int { return SRC == 3 ? (h0 + (l >> tw_shift)) * src.W + w0 + (l & (tw - 1)) : p0 + l; };
  const long long base = ((long long)n * C + c0) * HW;
  // 1. the tile loads go out first: everything below overlaps their latency
  float4 xq[VEC ? 2 : 1], gq[VEC ? 2 : 1];
  float2 pm[VEC ? 2 : 1], pg[VEC ? 2 : 1];   // SRC 3: pooled maxima and their gradient of the thread's two windows
  float xv[VEC ? 1 : 8], gv[VEC ? 1 : 8];
  const int vc = threadIdx.x >> 3, vj = threadIdx.x & 7;   // VEC: channel row (and + 32), float4 column
  if (VEC) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int p = pix(4 * vj);
      const long long off = base + (long long)(vc + 32 * i) * HW + p;
      xq[i] = gq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      pm[i] = pg[i] = make_float2(0.f, 0.f);
      if (p < HW) {
        xq[i] = *reinterpret_cast<const float4*>(src.in + off);
        if (SRC == 2) gq[i] = *reinterpret_cast<const float4*>(src.dy + off);
        if (SRC == 3) {
          const int hh = h0 + ((4 * vj) >> tw_shift), ww = w0 + ((4 * vj) & (tw - 1));
          const long long po = (((long long)n * C + c0 + vc + 32 * i) * (src.H >> 1) + (hh >> 1)) * (src.W >> 1) + (ww >> 1);
          pm[i] = *reinterpret_cast<const float2*>(src.pooled + po);
          pg[i] = *reinterpret_cast<const float2*>(src.dpooled + po);
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = p0 + tx;
      const long long off = base + (long long)(ty + 8 * i) * HW + p;
      xv[i] = (p < HW) ? src.in[off] : 0.f;
      gv[i] = (SRC == 2 && p < HW) ? src.dy[off] : 0.f;
    }
  }
  // 2. scale, and the per-channel constants (64 threads, one channel each)
  const float scale = f16_scale(*max_bits);
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) *inv_scale_out = 1.f / scale;
  if (SRC != 0) {
    if (threadIdx.x < 64) {
      const int c = c0 + threadIdx.x;
      const float inv = src.invstd[c];
      const float k0 = src.gamma[c] * inv;
      cst[0][threadIdx.x] = src.mean[c];
      cst[1][threadIdx.x] = k0;
      if (SRC == 1) {
        cst[2][threadIdx.x] = src.beta[c];
      } else {
        cst[2][threadIdx.x] = src.sums[c * 2] * src.inv_count;
        cst[3][threadIdx.x] = inv * inv * src.sums[c * 2 + 1] * src.inv_count;
        cst[4][threadIdx.x] = (src.act || SRC == 3) ? src.beta[c] : 0.f;
      }
    }
    __syncthreads();
  }
  // 3. values -> transpose tile (+ channel sums of the tile)
  const long long prow = ((long long)n * gridDim.x + blockIdx.x) * C + c0;
  if (VEC) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = vc + 32 * i;
      const bool ok = pix(4 * vj) < HW;
      float v0, v1, v2, v3;
      if (SRC == 3) {
        // own four activations and those of the vertical neighbour row (the other half of the two windows)
        const float mu = cst[0][c], k0 = cst[1][c], zb = cst[4][c];
        const float z0 = bn_affine(xq[i].x, mu, k0, zb), z1 = bn_affine(xq[i].y, mu, k0, zb);
        const float z2 = bn_affine(xq[i].z, mu, k0, zb), z3 = bn_affine(xq[i].w, mu, k0, zb);
        const float y0 = src.act ? fmaxf(z0, 0.f) : z0, y1 = src.act ? fmaxf(z1, 0.f) : z1;
        const float y2 = src.act ? fmaxf(z2, 0.f) : z2, y3 = src.act ? fmaxf(z3, 0.f) : z3;
        const int partner = tw >> 2;   // lane distance of the thread holding the other row of the same windows
        const float n0 = __shfl_xor_sync(0xffffffffu, y0, partner), n1 = __shfl_xor_sync(0xffffffffu, y1, partner);
        const float n2 = __shfl_xor_sync(0xffffffffu, y2, partner), n3 = __shfl_xor_sync(0xffffffffu, y3, partner);
        const int ca = (y0 == pm[i].x) + (y1 == pm[i].x) + (n0 == pm[i].x) + (n1 == pm[i].x);
        const int cb = (y2 == pm[i].y) + (y3 == pm[i].y) + (n2 == pm[i].y) + (n3 == pm[i].y);
        const float sa = ca ? pg[i].x / (float)ca : 0.f, sb = cb ? pg[i].y / (float)cb : 0.f;
        float g0 = (y0 == pm[i].x) ? sa : 0.f, g1 = (y1 == pm[i].x) ? sa : 0.f;
        float g2 = (y2 == pm[i].y) ? sb : 0.f, g3 = (y3 == pm[i].y) ? sb : 0.f;
        if (src.act) {
          if (!(z0 >= 0.f)) g0 = 0.f;
          if (!(z1 >= 0.f)) g1 = 0.f;
          if (!(z2 >= 0.f)) g2 = 0.f;
          if (!(z3 >= 0.f)) g3 = 0.f;
        }
        const float k1 = cst[2][c], k2 = cst[3][c];
        v0 = k0 * (g0 - k1 - (xq[i].x - mu) * k2); v1 = k0 * (g1 - k1 - (xq[i].y - mu) * k2);
        v2 = k0 * (g2 - k1 - (xq[i].z - mu) * k2); v3 = k0 * (g3 - k1 - (xq[i].w - mu) * k2);
      } else {
        v0 = stage_value<SRC>(xq[i].x, gq[i].x, src.act, cst, c); v1 = stage_value<SRC>(xq[i].y, gq[i].y, src.act, cst, c);
        v2 = stage_value<SRC>(xq[i].z, gq[i].z, src.act, cst, c); v3 = stage_value<SRC>(xq[i].w, gq[i].w, src.act, cst, c);
      }
      if (!ok) v0 = v1 = v2 = v3 = 0.f;
      tile[c][4 * vj + 0] = v0; tile[c][4 * vj + 1] = v1; tile[c][4 * vj + 2] = v2; tile[c][4 * vj + 3] = v3;
      if (SUM) {
        // pixel order of the sum: 4 per thread, then the 8 threads of the channel row (fixed, like the scalar form)
        float t = (v0 + v1) + (v2 + v3);
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        if (vj == 0) partial[prow + c] = t;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = ty + 8 * i;
      float v = stage_value<(SRC == 3 ? 2 : SRC)>(xv[i], gv[i], src.act, cst, c);
      if (p0 + tx >= HW) v = 0.f;
      tile[c][tx] = v;
      if (SUM) {
        const float t = warp_sum(v);
        if (tx == 0) partial[prow + c] = t;
      }
    }
  }
  __syncthreads();
  __half* dst = out + ((long long)n * HW * G + g) * 128;   // + p * G * 128 + plane * 64 + c
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pl = ty + 8 * i, p = pix(pl);
    if (p < HW) {
      __half h0_, l0, h1, l1;
      f16_split(tile[2 * tx][pl], scale, h0_, l0);
      f16_split(tile[2 * tx + 1][pl], scale, h1, l1);
      __half2* row = reinterpret_cast<__half2*>(dst + (long long)p * G * 128);
      row[tx] = __halves2half2(h0_, h1);
      row[32 + tx] = __halves2half2(l0, l1);
    }
  }
}

// dst[plane][rs][j][kc] = split(w[j*sj + kc*sc + rs])
// (A single launch doing the magnitude pass, a grid barrier and the re-layout was built and measured: 11 us per filter
//  against 4 + 5 us for the two launches below — the barrier's global round trips cost more than a launch.)
__global__ void __launch_bounds__(256) filter_relayout_f16_kernel(__half* __restrict__ dst, const float* __restrict__ w,
                                                                  int J, int KC, int RS, long long sj, long long sc,
                                                                  const unsigned int* __restrict__ max_bits,
                                                                  float* __restrict__ inv_scale_out) {
  const float scale = f16_scale(*max_bits);
  if (blockIdx.x == 0 && threadIdx.x == 0) *inv_scale_out = 1.f / scale;
  const long long n = (long long)J * KC * RS;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int kc = (int)(e % KC);
    const long long t = e / KC;
    const int j = (int)(t % J);
    const int rs = (int)(t / J);
    __half h, l;
    f16_split(w[(long long)j * sj + (long long)kc * sc + rs], scale, h, l);
    dst[e] = h;
    dst[n + e] = l;
  }
}

static int launch_absmax(unsigned int* bits, const float* in, long long n) {
  if (cudaMemsetAsync(bits, 0, 4, g_stream) != cudaSuccess) return set_error("cudaMemsetAsync failed");
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  long long g = ceil_div(n, 256 * 16);
  if (g > sms * 8) g = sms * 8;
  if (g < 1) g = 1;
  NG_LAUNCH(absmax_kernel, (int)g, 256, 0, bits, in, n);
  NG_CHECK_LAUNCH();
  return 0;
}

// Both staged layouts of one KCRS filter tensor in one pass over it (the weights are the same for the forward and
// the backward of a step, so the layer stages them once, at the forward): fprop reads [plane][rs][k][c],
// dgrad [plane][rs][c][k].  One thread per (k, c, rs) source element, read in storage order.
__global__ void __launch_bounds__(256) filter_relayout2_f16_kernel(__half* __restrict__ dst_f, __half* __restrict__ dst_d,
                                                                   const float* __restrict__ w, int K, int C, int RS,
                                                                   const unsigned int* __restrict__ max_bits,
                                                                   float* __restrict__ inv_f, float* __restrict__ inv_d) {
  const float scale = f16_scale(*max_bits);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *inv_f = 1.f / scale;
    if (dst_d) *inv_d = 1.f / scale;
  }
  const long long n = (long long)K * C * RS;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int rs = (int)(e % RS);
    const long long t = e / RS;
    const int c = (int)(t % C);
    const int k = (int)(t / C);
    __half h, l;
    f16_split(w[e], scale, h, l);
    const long long ef = ((long long)rs * K + k) * C + c;
    dst_f[ef] = h;
    dst_f[n + ef] = l;
    if (dst_d) {
      const long long ed = ((long long)rs * C + c) * K + k;
      dst_d[ed] = h;
      dst_d[n + ed] = l;
    }
  }
}

static inline float* staged_meta(void* staged, long long n_elems) {
  return reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(staged) + n_elems * 4);
}

// Stage a tensor (C % 64 == 0) as the pair-interleaved FP16 channels-last copy.  `staged` holds N*C*HW*4 + 64
// bytes; sums (optional) receives the per-channel sums of the staged values; known_bits (optional) points at the
// bits of max |value| (or an upper bound) established by the producer, else an absmax pass over src.in runs first.
template <int SRC>
static int stage_f16(void* staged, const StageSrc& src, float* sums, const unsigned int* known_bits, int N, int C, int HW) {
  if (C % 64 != 0) return set_error("to_nhwc_f16: channel count %d is not a multiple of 64", C);
  const long long n = (long long)N * C * HW;
  float* meta = staged_meta(staged, n);
  const unsigned int* bits = known_bits;
  if (!bits) {
    if (SRC != 0) return set_error("to_nhwc_f16: a fused source needs the magnitude of its values");
    unsigned int* own = reinterpret_cast<unsigned int*>(meta + 1);
    if (launch_absmax(own, src.in, n)) return 1;
    bits = own;
  }
  __half* out = reinterpret_cast<__half*>(staged);
  dim3 grid((unsigned)ceil_div(HW, 32), (unsigned)(C / 64), (unsigned)N);
  const bool vec = HW % 4 == 0 && (reinterpret_cast<uintptr_t>(src.in) & 15u) == 0 &&
                   (SRC != 2 || (reinterpret_cast<uintptr_t>(src.dy) & 15u) == 0);
  if (SRC == 3 && (!vec || HW % 32 != 0)) return set_error("to_nhwc_f16: the pooled source needs whole 32-pixel tiles");
  if (sums) {
    const long long rows = (long long)N * grid.x;
    float* partial = nullptr;
    if (pool_alloc((void**)&partial, (uint64_t)rows * C * sizeof(float))) return 1;
    if (vec) NG_LAUNCH((stage_f16_kernel<SRC, true, true>), grid, 256, 0, out, src, partial, C, HW, bits, meta);
    else NG_LAUNCH((stage_f16_kernel<(SRC == 3 ? 2 : SRC), true, false>), grid, 256, 0, out, src, partial, C, HW, bits, meta);
    NG_CHECK_LAUNCH();
    if (channel_partial_sum(sums, partial, rows, C)) return 1;
    pool_free(partial);
  } else {
    if (vec) NG_LAUNCH((stage_f16_kernel<SRC, false, true>), grid, 256, 0, out, src, (float*)nullptr, C, HW, bits, meta);
    else NG_LAUNCH((stage_f16_kernel<(SRC == 3 ? 2 : SRC), false, false>), grid, 256, 0, out, src, (float*)nullptr, C, HW, bits, meta);
    NG_CHECK_LAUNCH();
  }
  return 0;
}

int to_nhwc_f16(void* staged, const float* in, float* sums, int N, int C, int HW, const unsigned int* known_bits) {
  StageSrc src;
  memset(&src, 0, sizeof(src));
  src.in = in;
  return stage_f16<0>(staged, src, sums, known_bits, N, C, HW);
}

int to_nhwc_f16_bn(void* staged, const float* x, const float* gamma, const float* beta, const float* mean,
                   const float* invstd, int act, const unsigned int* absmax_bits, int N, int C, int HW) {
  StageSrc src;
  memset(&src, 0, sizeof(src));
  src.in = x; src.gamma = gamma; src.beta = beta; src.mean = mean; src.invstd = invstd; src.act = act;
  return stage_f16<1>(staged, src, nullptr, absmax_bits, N, C, HW);
}

int to_nhwc_f16_bnbwd(void* staged, float* chan_sums, const float* dy, const float* x, const float* gamma,
                      const float* beta, const float* mean, const float* invstd, const float* bn_sums, int act,
                      const unsigned int* bound_bits, int N, int C, int HW) {
  StageSrc src;
  memset(&src, 0, sizeof(src));
  src.in = x; src.dy = dy; src.gamma = gamma; src.beta = beta; src.mean = mean; src.invstd = invstd;
  src.sums = bn_sums; src.inv_count = 1.f / ((float)N * (float)HW); src.act = act;
  return stage_f16<2>(staged, src, chan_sums, bound_bits, N, C, HW);
}

int to_nhwc_f16_bnbwd_pool(void* staged, float* chan_sums, const float* dpooled, const float* pooled, const float* x,
                           const float* gamma, const float* beta, const float* mean, const float* invstd,
                           const float* bn_sums, int act, const unsigned int* bound_bits, int N, int C, int H, int W) {
  if (H % 2 || W % 8 || ((H * W) % 32) != 0) return set_error("to_nhwc_f16_bnbwd_pool: needs even H, W %% 8 == 0, H*W %% 32 == 0");
  const int tw = (W % 16 == 0) ? 16 : 8;
  const int th = 32 / tw;
  if (H % th) return set_error("to_nhwc_f16_bnbwd_pool: H must be a multiple of %d", th);
  if ((reinterpret_cast<uintptr_t>(pooled) & 7u) || (reinterpret_cast<uintptr_t>(dpooled) & 7u))
    return set_error("to_nhwc_f16_bnbwd_pool: pooled tensors must be 8-byte aligned");
  StageSrc src;
  memset(&src, 0, sizeof(src));
  src.in = x; src.gamma = gamma; src.beta = beta; src.mean = mean; src.invstd = invstd;
  src.sums = bn_sums; src.inv_count = 1.f / ((float)N * (float)H * (float)W); src.act = act;
  src.pooled = pooled; src.dpooled = dpooled; src.H = H; src.W = W; src.tw = tw;
  return stage_f16<3>(staged, src, chan_sums, bound_bits, N, C, H * W);
}

// Stage the filters of a layer for fprop (and, with staged_d, for dgrad): each buffer holds K*C*RS*4 + 64 bytes.
int stage_filters_f16(void* staged_f, void* staged_d, const float* w, int K, int C, int RS) {
  const long long n = (long long)K * C * RS;
  float* mf = staged_meta(staged_f, n);
  float* md = staged_d ? staged_meta(staged_d, n) : nullptr;
  if (launch_absmax(reinterpret_cast<unsigned int*>(mf + 1), w, n)) return 1;
  int g = (int)ceil_div(n, 256);
  if (g > 148 * 8) g = 148 * 8;
  NG_LAUNCH(filter_relayout2_f16_kernel, g, 256, 0, reinterpret_cast<__half*>(staged_f),
            reinterpret_cast<__half*>(staged_d), w, K, C, RS, (const unsigned int*)(mf + 1), mf, md);
  NG_CHECK_LAUNCH();
  return 0;
}

// ================================================================== gather kernel (fprop / dgrad)
constexpr int GF_THREADS = 224;     // producer, MMA issuer, 4 epilogue warps, second producer
constexpr int GF_CCHUNK = 64;       // channels per pipeline stage = one 128-byte FP16 row = 4 MMAs steps of K = 16
constexpr int GF_MAX_STAGES = 4;
constexpr uint32_t GF_A_PLANE = 128u * 128u;   // one 128-pixel x 64-channel FP16 tile

struct GatherF16P {
  CUtensorMap tm_a;  // staged activations as (64, W, H, N, 2G): box {64, w_tile, rows_tile, imgs_tile, 2}
  CUtensorMap tm_b;  // filters [plane][tap][j][kc] as (KC, J, RS, 2): box {64, n_tile, 1, 2}
  float* out;
  const float* bias;
  const float* a_meta;   // inv_scale of the activations
  const float* b_meta;   // inv_scale of the filters
  int bias_vec;
  int n_img, J, KC, RS, OH, OW;
  int w_tile, rows_tile, imgs_tile;   // w_tile * rows_tile * imgs_tile == 128
  int tiles_w, tiles_h, tiles_img;
  int n_tile, n_tiles;
  int c_chunks;
  uint32_t idesc_all, idesc_hi;       // N = 2 * n_tile ([B_hi | B_lo]) and N = n_tile (B_hi)
  uint32_t stage_b_bytes;             // 2 * n_tile * 128
  int stages;
  uint32_t tmem_cols;
  int relu;
  uint32_t* dbg;
  short tap_dx[64], tap_dy[64];
};

__global__ void __launch_bounds__(GF_THREADS, 1) umma_conv_gather_f16_kernel(const __grid_constant__ GatherF16P p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = 2u * GF_A_PLANE + p.stage_b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + GF_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + GF_MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_a);
    tma_prefetch_desc(&p.tm_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_cols = 2u * (uint32_t)p.n_tile;   // columns of one accumulator: [0,n) main, [n,2n) A_hi*B_lo

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_img;
  const int total_tiles = m_tiles * p.n_tiles;
  const int k_iters = p.RS * p.c_chunks;

  if (warp == 0 || warp == 6) {
    if (lane == 0) {
      const uint32_t pid = warp == 0 ? 0u : 1u;   // two TMA-issuing warps alternate stages (one thread issues a box in ~200 cycles)
      uint32_t it = 0;
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_t = tile % p.n_tiles;
        int m_t = tile / p.n_tiles;
        const int tw = m_t % p.tiles_w; m_t /= p.tiles_w;
        const int th = m_t % p.tiles_h;
        const int ti = m_t / p.tiles_h;
        const int col0 = tw * p.w_tile, row0 = th * p.rows_tile, img0 = ti * p.imgs_tile;
        for (int rs = 0; rs < p.RS; ++rs) {
          const int cx = col0 + p.tap_dx[rs], cy = row0 + p.tap_dy[rs];
          for (int cc = 0; cc < p.c_chunks; ++cc, ++it) {
            if ((it & 1u) == pid) {
              mbar_wait(&empty_bar[s], ph ^ 1u, p.dbg, 0x100u + s);
              uint8_t* sa = smem + (size_t)s * stage_bytes;
              mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
              tma_load_5d(sa, &p.tm_a, &full_bar[s], 0, cx, cy, img0, 2 * cc);
              tma_load_4d(sa + 2u * GF_A_PLANE, &p.tm_b, &full_bar[s], cc * GF_CCHUNK, n_t * p.n_tile, rs, 0);
            }
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the whole warp walks the pipeline (uniform control flow), one elected lane issues.
    int s = 0;
    uint32_t ph = 0;
    int as = 0;
    uint32_t aph = 0;
    const uint32_t dhi = smem_desc_hi(1024u, SWZ_128B);
    const uint32_t smem0 = smem_u32(smem);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[as], aph ^ 1u, p.dbg, 0x200u + as);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)as * acc_cols;
      uint32_t accumulate = 0;
      for (int it = 0; it < k_iters; ++it) {
        mbar_wait(&full_bar[s], ph, p.dbg, 0x300u + s);
        tc_fence_after();
        const uint32_t sa = smem0 + (uint32_t)s * stage_bytes;
        const uint32_t a_hi = smem_desc_lo(sa, 16u), a_lo = smem_desc_lo(sa + GF_A_PLANE, 16u);
        const uint32_t b_all = smem_desc_lo(sa + 2u * GF_A_PLANE, 16u);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 32 bytes along K per step
            mma_f16(d_tmem, desc64(a_hi + 2u * k, dhi), desc64(b_all + 2u * k, dhi), p.idesc_all, k == 0 ? accumulate : 1u);
            mma_f16(d_tmem, desc64(a_lo + 2u * k, dhi), desc64(b_all + 2u * k, dhi), p.idesc_hi, 1u);
          }
          mma_commit(&empty_bar[s]);
        }
        accumulate = 1u;
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) mma_commit(&tmem_full[as]);
      __syncwarp();
      as ^= 1;
      if (as == 0) aph ^= 1u;
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;
    const int g = m / p.w_tile, wcol = m - g * p.w_tile;
    const int gi = g / p.rows_tile, gr = g - gi * p.rows_tile;
    const long long opix = (long long)p.OH * p.OW;
    const float inv = __ldg(p.a_meta) * __ldg(p.b_meta);
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_t = tile % p.n_tiles;
      int m_t = tile / p.n_tiles;
      const int tw = m_t % p.tiles_w; m_t /= p.tiles_w;
      const int th = m_t % p.tiles_h;
      const int ti = m_t / p.tiles_h;
      const int col = tw * p.w_tile + wcol, row = th * p.rows_tile + gr, img = ti * p.imgs_tile + gi;
      const bool valid = (col < p.OW) && (row < p.OH) && (img < p.n_img);
      float* out_pix = p.out + ((long long)img * p.J) * opix + (long long)row * p.OW + col;
      mbar_wait(&tmem_full[as], aph, p.dbg, 0x400u + as);
      tc_fence_after();
      for (int j0 = 0; j0 < p.n_tile; j0 += 32) {
        uint32_t r[32], r2[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * acc_cols + (uint32_t)j0;
        tmem_ld_32x32(taddr, r);
        tmem_ld_32x32(taddr + (uint32_t)p.n_tile, r2);
        tmem_ld_wait();
        const int jbase = n_t * p.n_tile + j0;
        if (valid) {
          float* dst = out_pix + (long long)jbase * opix;
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint((__uint_as_float(r[i]) + __uint_as_float(r2[i])) * inv);
          if (jbase + 32 <= p.J) {
            if (p.bias_vec) {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + jbase);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 bv = __ldg(b4 + i);
                r[4 * i + 0] = __float_as_uint(__uint_as_float(r[4 * i + 0]) + bv.x);
                r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + bv.y);
                r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + bv.z);
                r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + bv.w);
              }
            } else if (p.bias) {
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __ldg(p.bias + jbase + i));
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(fmaxf(__uint_as_float(r[i]), 0.f));
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              *dst = __uint_as_float(r[i]);
              dst += opix;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int j = jbase + i;
              if (j < p.J) {
                float v = __uint_as_float(r[i]);
                if (p.bias) v += __ldg(p.bias + j);
                if (p.relu) v = fmaxf(v, 0.f);
                dst[(long long)i * opix] = v;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      as ^= 1;
      if (as == 0) aph ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static inline int pow2_ceil_i(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

static uint32_t* debug_words(bool* debug) {
  static uint32_t* dbg_host = nullptr;
  *debug = getenv("NG_UMMA_DEBUG") != nullptr;
  if (!*debug) return nullptr;
  if (!dbg_host) cudaHostAlloc((void**)&dbg_host, 64 * sizeof(uint32_t), cudaHostAllocMapped);
  memset(dbg_host, 0, 64 * sizeof(uint32_t));
  return dbg_host;
}

// `act` is the gathered activation tensor [n_img, KC, AH, AW]: NCHW FP32, or (act_is_staged) its to_nhwc_f16 copy.
// `filt` are the KCRS filters addressed as filt[j*sj + kc*sc + rs]; the output is [n_img, J, OH, OW] NCHW FP32.
static int launch_gather_f16(float* out, const void* act, const float* filt, const float* bias, int n_img, int KC,
                             int AH, int AW, int J, int R, int S, long long sj, long long sc, int OH, int OW,
                             const int* tap_dy, const int* tap_dx, int relu, int act_is_staged,
                             const void* filt_staged) {
  const int RS = R * S;
  GatherF16P p;
  memset(&p, 0, sizeof(p));
  p.out = out;
  p.bias = bias;
  p.bias_vec = (bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & 15u) == 0) ? 1 : 0;
  p.n_img = n_img; p.J = J; p.KC = KC; p.RS = RS; p.OH = OH; p.OW = OW;
  int w_tile = pow2_ceil_i(OW);
  if (w_tile > 32) w_tile = 32;
  int rows = 128 / w_tile;
  const int oh2 = pow2_ceil_i(OH);
  if (rows > oh2) rows = oh2;
  p.w_tile = w_tile;
  p.rows_tile = rows;
  p.imgs_tile = 128 / (w_tile * rows);
  p.tiles_w = (int)ceil_div(OW, p.w_tile);
  p.tiles_h = (int)ceil_div(OH, rows);
  p.tiles_img = (int)ceil_div(n_img, p.imgs_tile);
  int n_tile = (int)ceil_div(J, 32) * 32;
  if (n_tile > 128) n_tile = 128;
  p.n_tile = n_tile;
  p.n_tiles = (int)ceil_div(J, n_tile);
  p.c_chunks = KC / GF_CCHUNK;
  p.idesc_all = make_idesc_f16(128, 2 * n_tile, 0, 0);
  p.idesc_hi = make_idesc_f16(128, n_tile, 0, 0);
  p.stage_b_bytes = 2u * (uint32_t)n_tile * 128u;
  const uint32_t stage_bytes = 2u * GF_A_PLANE + p.stage_b_bytes;
  int stages = (int)((227u * 1024u - 2048u) / stage_bytes);
  if (stages > GF_MAX_STAGES) stages = GF_MAX_STAGES;
  if (stages < 2) return -1;
  p.stages = stages;
  p.tmem_cols = (uint32_t)pow2_ceil_i(4 * n_tile < 32 ? 32 : 4 * n_tile);
  p.relu = relu;
  for (int t = 0; t < RS; ++t) { p.tap_dx[t] = (short)tap_dx[t]; p.tap_dy[t] = (short)tap_dy[t]; }

  const long long act_elems = (long long)n_img * KC * AH * AW;
  const long long filt_elems = (long long)RS * J * KC;
  void* act_t = nullptr;
  void* wt_own = nullptr;
  if (!act_is_staged && pool_alloc(&act_t, (uint64_t)act_elems * 4 + 64)) return 1;
  if (!filt_staged && pool_alloc(&wt_own, (uint64_t)filt_elems * 4 + 64)) { pool_free(act_t); return 1; }
  void* wt = filt_staged ? const_cast<void*>(filt_staged) : wt_own;
  int rc = 0;
  const void* act_s = act_is_staged ? act : act_t;
  if (!act_is_staged) rc = to_nhwc_f16(act_t, reinterpret_cast<const float*>(act), nullptr, n_img, KC, AH * AW, nullptr);
  float* wmeta = staged_meta(wt, filt_elems);   // [0] 1/scale, [1] bits of max|w|
  if (rc == 0 && !filt_staged) {
    // max|w| over the filter tensor: the KCRS block is contiguous whichever of (j, kc) is the outer axis
    rc = launch_absmax(reinterpret_cast<unsigned int*>(wmeta + 1), filt, filt_elems);
    if (rc == 0) {
      int g = (int)ceil_div(filt_elems, 256);
      if (g > 148 * 8) g = 148 * 8;
      NG_LAUNCH(filter_relayout_f16_kernel, g, 256, 0, reinterpret_cast<__half*>(wt), filt, J, KC, RS, sj, sc,
                (const unsigned int*)(wmeta + 1), wmeta);
    }
  }
  if (rc == 0) {
    const int G = KC / 64;
    const uint64_t pix = (uint64_t)KC * 4u;   // bytes per staged pixel
    const uint64_t dims[5] = {64u, (uint64_t)AW, (uint64_t)AH, (uint64_t)n_img, (uint64_t)(2 * G)};
    const uint64_t strides[4] = {pix, (uint64_t)AW * pix, (uint64_t)AH * AW * pix, 128u};
    const uint32_t box[5] = {64u, (uint32_t)p.w_tile, (uint32_t)p.rows_tile, (uint32_t)p.imgs_tile, 2u};
    if (encode_tensor_map(&p.tm_a, act_s, 5, dims, strides, box, SWZ_128B, /*fp16=*/1)) rc = -1;
  }
  if (rc == 0) {
    const uint64_t dims[4] = {(uint64_t)KC, (uint64_t)J, (uint64_t)RS, 2u};
    const uint64_t strides[3] = {(uint64_t)KC * 2u, (uint64_t)J * KC * 2u, (uint64_t)filt_elems * 2u};
    const uint32_t box[4] = {64u, (uint32_t)n_tile, 1u, 2u};
    if (encode_tensor_map(&p.tm_b, wt, 4, dims, strides, box, SWZ_128B, /*fp16=*/1)) rc = -1;
  }
  p.a_meta = staged_meta(const_cast<void*>(act_s), act_elems);
  p.b_meta = wmeta;
  if (rc == 0) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(umma_conv_gather_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(227 * 1024));
      if (e != cudaSuccess) rc = set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      configured = (e == cudaSuccess);
    }
  }
  if (rc == 0) {
    const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 + 256;
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_img * p.n_tiles;
    const int sms = g_sm_count > 0 ? g_sm_count : 148;
    const int grid = total_tiles < sms ? total_tiles : sms;
    bool debug;
    p.dbg = debug_words(&debug);
    if (debug)
      fprintf(stderr, "[umma f16] gather: tiles %d grid %d n_tile %d stages %d w_tile %d rows %d imgs %d tmem %u\n",
              total_tiles, grid, p.n_tile, p.stages, p.w_tile, p.rows_tile, p.imgs_tile, p.tmem_cols);
    NG_LAUNCH(umma_conv_gather_f16_kernel, grid, GF_THREADS, smem_bytes, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = set_error("umma_conv_gather_f16_kernel launch failed: %s", cudaGetErrorString(e));
    if (debug) {
      cudaError_t e2 = cudaStreamSynchronize(g_stream);
      fprintf(stderr, "[umma f16] sync: %s ; dbg: %08x %08x %08x\n", cudaGetErrorString(e2), p.dbg[0], p.dbg[1], p.dbg[2]);
    }
  }
  pool_free(wt_own);
  pool_free(act_t);
  return rc;
}

int conv_fprop_f16(float* y, const void* x, const float* w, const float* bias, int N, int C, int H, int W, int K, int R,
                   int S, int stride, int pad_t, int pad_l, int OH, int OW, int relu, int x_is_staged, int w_is_staged) {
  if (stride != 1 || !f16x3_supported(C, K, R * S)) return -1;
  int dy[64], dx[64];
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) { dy[r * S + s] = r - pad_t; dx[r * S + s] = s - pad_l; }
  return launch_gather_f16(y, x, w, bias, N, C, H, W, K, R, S, (long long)C * R * S, (long long)R * S, OH, OW, dy, dx,
                           relu, x_is_staged, w_is_staged ? w : nullptr);
}

int conv_dgrad_f16(float* dx_out, const void* dyv, const float* w, int N, int C, int H, int W, int K, int R, int S,
                   int stride, int pad_t, int pad_l, int OH, int OW, int dy_is_staged, int w_is_staged) {
  if (stride != 1 || !f16x3_supported(C, K, R * S)) return -1;
  int dy[64], dx[64];
  // dx[n,c,y,x] = sum_{k,r,s} dy[n,k,y+pt-r,x+pl-s] * w[k,c,r,s]
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) { dy[r * S + s] = pad_t - r; dx[r * S + s] = pad_l - s; }
  return launch_gather_f16(dx_out, dyv, w, nullptr, N, K, OH, OW, C, R, S, (long long)R * S, (long long)C * R * S, H, W,
                           dy, dx, 0, dy_is_staged, w_is_staged ? w : nullptr);
}

// ================================================================== wgrad kernel
//   per filter tap t = (r, s):   D_t[c][k] = sum_{n,oh,ow} x[n, oh+r-pt, ow+s-pl, c] * dy[n, oh, ow, k]
// Both operands are MN-major (channel-contiguous) 128-byte rows of one pixel's 64-channel group in the 128B
// swizzle; the reduction runs over pixels (16 per MMA, 32 per pipeline stage).  Shared memory per stage:
//   A: per tap, the blocks [g][plane] (32 pixels x 128 B = 4 KB each) of the channel tile, in staged order
//      (g0 hi, g0 lo, g1 hi, g1 lo): M = 128 rows = (128 / Cm) taps x Cm channels, Cm in {64, 128}; either way
//      the 64-row groups of A_hi start 8 KB apart (LBO) and A_lo = A_hi + 4 KB;
//   B: the blocks [g][plane] of the output-channel tile (ng = n_tile / 64 groups): [B_hi | B_lo] interleaved by
//      group IS one N = 2 n_tile MN-major operand with LBO = 4 KB, so D[:, g*128 + (0..63)] = A_hi*B_hi(g) and
//      D[:, g*128 + 64 + (0..63)] = A_hi*B_lo(g) come from one instruction; A_lo*B_hi(g) is one N = 64 MMA per
//      group accumulating into the first half.
// The tensor core's FP32 accumulation is not round-to-nearest (a 17k-pixel chain drifts by 1.3e-4), so the
// reduction is cut into groups of WF_PROMOTE chunks that ping-pong between two TMEM accumulators while the
// epilogue warps add each finished group into FP32 registers, as in the 3xTF32 kernel.
constexpr int WF_THREADS = 224;
constexpr int WF_MAX_STAGES = 6;
constexpr int WF_PROMOTE = 8;
constexpr uint32_t WF_BLOCK = 4096u;   // one (group, plane) block: 32 pixels x 128 bytes
constexpr uint32_t WF_A_BYTES = 4u * WF_BLOCK;

struct WgradF16P {
  CUtensorMap tm_x;   // staged x as (64, W, H, 2*C/64, N):   box {64, w_c, rows_c, 2*Cm/64, 1}
  CUtensorMap tm_dy;  // staged dy as (64, OW, OH, 2*K/64, N): box {64, w_c, rows_c, 2*ng, 1}
  float* ws;          // [splits][RS][C][K], unscaled
  int C, K, RS, S;
  int Cm, taps_per_cta, tap_groups, c_tiles, n_tile, n_tiles, ng;
  int w_c, rows_c, chunks_w, chunks_h, total_chunks, chunks_per_split, splits;
  int pad_t, pad_l;
  uint32_t idesc_all, idesc_g, stage_b_bytes, tmem_cols;
  int stages;
  int ct, cn;   // thread-block cluster = ct tap groups x cn output-channel tiles of the same pixels
  uint32_t* dbg;
};

// CL (thread-block clusters, experiment: see conv_wgrad_f16): the kernel streams both operands with no reuse inside
// a CTA, 32 KB per 32-pixel chunk for 384 tensor cycles, and every (tap group, k tile) CTA of the same pixels
// re-reads the same x / dy chunks from L2 (18x redundancy at 256 -> 256 channels).  The ct x cn CTAs of a cluster
// work on the same chunk at the same time: the x chunk of a tap group is multicast to the cn CTAs of its k tiles
// and the dy chunk of a k tile to the ct CTAs of its tap groups (the issuing CTA rotates with the chunk), which
// divides the L2 reads by up to ct*cn/(ct+cn).  A slot is refilled only after every CTA of the cluster released
// it (tcgen05.commit multicast to all their empty barriers).
template <bool CL>
__global__ void __launch_bounds__(WF_THREADS, 1) umma_conv_wgrad_f16_kernel(const __grid_constant__ WgradF16P p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = WF_A_BYTES + p.stage_b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + WF_MAX_STAGES;
  uint64_t* acc_full = empty_bar + WF_MAX_STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_x);
    tma_prefetch_desc(&p.tm_dy);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL ? (uint32_t)(p.ct * p.cn) : 1u);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CL) cluster_sync_all();   // every CTA's barriers exist before a neighbour signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_cols = 2u * (uint32_t)p.n_tile;

  // work item: cluster = (split, k-tile group, c tile, tap-group cluster); rank inside = (ci, cj)
  const int csize = CL ? p.ct * p.cn : 1;
  const int rank = CL ? (int)cluster_ctarank() : 0;
  const int ci = rank % p.ct, cj = rank / p.ct;
  int item = blockIdx.x / csize;
  const int split = item % p.splits; item /= p.splits;
  const int n_t = (item % (p.n_tiles / p.cn)) * p.cn + cj; item /= (p.n_tiles / p.cn);
  const int c_t = item % p.c_tiles;
  const int tg = (item / p.c_tiles) * p.ct + ci;
  const int chunk_begin = split * p.chunks_per_split;
  int chunk_end = chunk_begin + p.chunks_per_split;
  if (chunk_end > p.total_chunks) chunk_end = p.total_chunks;
  const int n_chunks = chunk_end > chunk_begin ? chunk_end - chunk_begin : 0;
  const int n_groups = (n_chunks + WF_PROMOTE - 1) / WF_PROMOTE;
  const int tap0 = tg * p.taps_per_cta;
  int n_taps = p.RS - tap0;
  if (n_taps > p.taps_per_cta) n_taps = p.taps_per_cta;
  const int blocks_per_tap = p.Cm >> 5;   // (Cm / 64) groups x 2 planes
  const uint32_t tap_bytes = (uint32_t)blocks_per_tap * WF_BLOCK;

  if (warp == 0 || warp == 6) {
    if (lane == 0) {
      // producer `pid` takes chunks pid, pid + 2, ...; the chunk position is stepped, not divided
      const int pid = warp == 0 ? 0 : 1;
      int s = pid % p.stages;
      uint32_t ph = (uint32_t)(pid / p.stages) & 1u;
      int t = chunk_begin + pid;
      int cw = t % p.chunks_w; t /= p.chunks_w;
      int chh = t % p.chunks_h;
      int img = t / p.chunks_h;
      int tdx[2], tdy[2];
#pragma unroll
      for (int tl = 0; tl < 2; ++tl) {
        const int tap = tap0 + tl;
        tdy[tl] = tap / p.S - p.pad_t;
        tdx[tl] = tap % p.S - p.pad_l;
      }
      const uint32_t tx_bytes = (uint32_t)n_taps * tap_bytes + p.stage_b_bytes;
      const int xg0 = c_t * blocks_per_tap, yg0 = n_t * 2 * p.ng;
      // cluster ranks that share this CTA's x chunk (same tap group ci) / dy chunk (same k tile cj)
      uint16_t mask_x = 0, mask_dy = 0;
      if (CL) {
        for (int j = 0; j < p.cn; ++j) mask_x |= (uint16_t)(1u << (ci + p.ct * j));
        mask_dy = (uint16_t)(((1u << p.ct) - 1u) << (p.ct * cj));
      }
      int xi = pid % p.cn, yi = pid % p.ct;   // whose turn it is to issue chunk i: i % cn, i % ct (stepped)
      const int xstep = 2 % p.cn, ystep = 2 % p.ct;
      for (int i = pid; i < n_chunks; i += 2) {
        const int ow0 = cw * p.w_c, oh0 = chh * p.rows_c;
        if (CL) mbar_wait_cluster(&empty_bar[s], ph ^ 1u, p.dbg, 0x100u + s);
        else mbar_wait(&empty_bar[s], ph ^ 1u, p.dbg, 0x100u + s);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
        if (CL) {
          if (xi == cj) {
#pragma unroll
            for (int tl = 0; tl < 2; ++tl)
              if (tl < n_taps)
                tma_load_5d_multicast(sa + (size_t)tl * tap_bytes, &p.tm_x, &full_bar[s], 0, ow0 + tdx[tl], oh0 + tdy[tl],
                                      xg0, img, mask_x);
          }
          if (yi == ci) tma_load_5d_multicast(sa + WF_A_BYTES, &p.tm_dy, &full_bar[s], 0, ow0, oh0, yg0, img, mask_dy);
          xi += xstep; if (xi >= p.cn) xi -= p.cn;
          yi += ystep; if (yi >= p.ct) yi -= p.ct;
        } else {
#pragma unroll
          for (int tl = 0; tl < 2; ++tl)
            if (tl < n_taps)
              tma_load_5d(sa + (size_t)tl * tap_bytes, &p.tm_x, &full_bar[s], 0, ow0 + tdx[tl], oh0 + tdy[tl], xg0, img);
          tma_load_5d(sa + WF_A_BYTES, &p.tm_dy, &full_bar[s], 0, ow0, oh0, yg0, img);
        }
        s += 2;
        if (s >= p.stages) { s -= p.stages; ph ^= 1u; }
        cw += 2;
        while (cw >= p.chunks_w) { cw -= p.chunks_w; ++chh; }
        while (chh >= p.chunks_h) { chh -= p.chunks_h; ++img; }
      }
    }
  } else if (warp == 1) {
    int s = 0;
    uint32_t ph = 0;
    int ch = 0;
    const uint32_t dhi = smem_desc_hi(1024u, SWZ_128B);   // SBO: 8 pixel rows
    const uint32_t smem0 = smem_u32(smem);
    for (int g = 0; g < n_groups; ++g) {
      const int b = g & 1;
      const uint32_t d_tmem = tmem_base + (uint32_t)b * acc_cols;
      mbar_wait(&acc_empty[b], (uint32_t)(((g >> 1) & 1) ^ 1), p.dbg, 0x200u + b);
      tc_fence_after();
      const int g_end = ch + WF_PROMOTE < n_chunks ? ch + WF_PROMOTE : n_chunks;
      uint32_t accumulate = 0;
      for (; ch < g_end; ++ch) {
        mbar_wait(&full_bar[s], ph, p.dbg, 0x300u + s);
        tc_fence_after();
        const uint32_t sa = smem0 + (uint32_t)s * stage_bytes;
        const uint32_t a_hi = smem_desc_lo(sa, 2u * WF_BLOCK), a_lo = smem_desc_lo(sa + WF_BLOCK, 2u * WF_BLOCK);
        const uint32_t sb = sa + WF_A_BYTES;
        const uint32_t b_all = smem_desc_lo(sb, WF_BLOCK);
        const uint32_t b_h0 = smem_desc_lo(sb, 2u * WF_BLOCK), b_h1 = smem_desc_lo(sb + 2u * WF_BLOCK, 2u * WF_BLOCK);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {   // 16 pixel rows = 2048 bytes per step
            mma_f16(d_tmem, desc64(a_hi + 128u * k, dhi), desc64(b_all + 128u * k, dhi), p.idesc_all, k == 0 ? accumulate : 1u);
            mma_f16(d_tmem, desc64(a_lo + 128u * k, dhi), desc64(b_h0 + 128u * k, dhi), p.idesc_g, 1u);
            if (p.ng == 2)
              mma_f16(d_tmem + 128u, desc64(a_lo + 128u * k, dhi), desc64(b_h1 + 128u * k, dhi), p.idesc_g, 1u);
          }
          if (CL) mma_commit_multicast(&empty_bar[s], (uint16_t)((1u << (p.ct * p.cn)) - 1u));
          else mma_commit(&empty_bar[s]);
        }
        accumulate = 1u;
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) mma_commit(&acc_full[b]);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int tl = m / p.Cm, cl = m - tl * p.Cm;
    const int tap = tap0 + tl;
    const int c = c_t * p.Cm + cl;
    float* dst_row = p.ws + (((long long)split * p.RS + tap) * p.C + c) * p.K + (long long)n_t * p.n_tile;
    const bool row_ok = (tl < n_taps);
    float acc[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) acc[i] = 0.f;
    for (int g = 0; g < n_groups; ++g) {
      const int b = g & 1;
      mbar_wait(&acc_full[b], (uint32_t)((g >> 1) & 1), p.dbg, 0x400u + b);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)b * acc_cols;
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
        if (gg < p.ng) {
#pragma unroll
          for (int j0 = 0; j0 < 64; j0 += 32) {
            uint32_t r[32], r2[32];
            tmem_ld_32x32(t0 + (uint32_t)(gg * 128 + j0), r);
            tmem_ld_32x32(t0 + (uint32_t)(gg * 128 + 64 + j0), r2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[gg * 64 + j0 + i] += __uint_as_float(r[i]) + __uint_as_float(r2[i]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
    }
#pragma unroll
    for (int j0 = 0; j0 < 128; j0 += 32) {
      if (j0 < p.n_tile && row_ok && n_t * p.n_tile + j0 < p.K) {
        float4* d4 = reinterpret_cast<float4*>(dst_row + j0);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          d4[i] = make_float4(acc[j0 + 4 * i], acc[j0 + 4 * i + 1], acc[j0 + 4 * i + 2], acc[j0 + 4 * i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL) cluster_sync_all();   // no CTA leaves while a neighbour may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// dw[k][c][rs] = inv_x * inv_dy * sum_split ws[split][rs][c][k]   (fixed summation order)
__global__ void __launch_bounds__(256) wgrad_reduce_scaled_kernel(float* __restrict__ dw, const float* __restrict__ ws,
                                                                  int splits, int RS, int C, int K,
                                                                  const float* __restrict__ x_meta,
                                                                  const float* __restrict__ dy_meta) {
  const float inv = __ldg(x_meta) * __ldg(dy_meta);
  const long long n = (long long)RS * C * K;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(long long)s * n + e];
    const int k = (int)(e % K);
    const long long t = e / K;
    const int c = (int)(t % C);
    const int rs = (int)(t / C);
    dw[((long long)k * C + c) * RS + rs] = acc * inv;
  }
}

int conv_wgrad_f16(float* dw, const void* dyv, const void* x, int N, int C, int H, int W, int K, int R, int S,
                   int stride, int pad_t, int pad_l, int OH, int OW, int x_is_staged, int dy_is_staged) {
  const int RS = R * S;
  if (stride != 1 || !f16x3_supported(C, K, RS)) return -1;
  WgradF16P p;
  memset(&p, 0, sizeof(p));
  p.C = C; p.K = K; p.RS = RS; p.S = S;
  p.pad_t = pad_t; p.pad_l = pad_l;
  p.Cm = (C % 128 == 0) ? 128 : 64;
  p.taps_per_cta = 128 / p.Cm;
  p.tap_groups = (int)ceil_div(RS, p.taps_per_cta);
  p.c_tiles = C / p.Cm;
  p.n_tile = (K % 128 == 0) ? 128 : 64;
  p.ng = p.n_tile / 64;
  p.n_tiles = K / p.n_tile;
  int w_c = pow2_ceil_i(OW);
  if (w_c > 32) w_c = 32;
  p.w_c = w_c;
  p.rows_c = 32 / w_c;
  p.chunks_w = (int)ceil_div(OW, w_c);
  p.chunks_h = (int)ceil_div(OH, p.rows_c);
  p.total_chunks = N * p.chunks_h * p.chunks_w;
  // cluster shape: ct = the largest divisor (<= 8) of the tap-group count, cn = 2 k tiles when that still fits.
  // MEASURED (B200, config-2 layers): correct, and 1.5-1.6x SLOWER than independent CTAs (64->64@32^2: 506 vs 335 us,
  // 256->256@8^2: 259 vs 164 us, staging included): the lock-step of 5-6 CTAs on every stage costs more than the
  // multicast saves, and the per-SM TMA ingest rate (~45 B/clk/SM, tools/tma_bw.cu) is the same either way.
  // Kept behind NG_WGRAD_CLUSTER=1 as the record of the experiment; off by default.
  p.ct = 1;
  p.cn = 1;
  {
    const char* e = getenv("NG_WGRAD_CLUSTER");
    if (e && atoi(e) != 0) {
      for (int d = 8; d >= 2; --d)
        if (p.tap_groups % d == 0) { p.ct = d; break; }
      if (p.n_tiles % 2 == 0 && p.ct * 2 <= 8) p.cn = 2;
    }
  }
  const int items = p.tap_groups * p.c_tiles * p.n_tiles;
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  int splits = sms / items;
  if (splits < 1) splits = 1;
  if (splits > p.total_chunks) splits = p.total_chunks;
  p.chunks_per_split = (int)ceil_div(p.total_chunks, splits);
  splits = (int)ceil_div(p.total_chunks, p.chunks_per_split);
  p.splits = splits;
  p.idesc_all = make_idesc_f16(128, 2 * p.n_tile, /*a MN-major*/ 1, /*b MN-major*/ 1);
  p.idesc_g = make_idesc_f16(128, 64, 1, 1);
  p.stage_b_bytes = (uint32_t)(2 * p.ng) * WF_BLOCK;
  const uint32_t stage_bytes = WF_A_BYTES + p.stage_b_bytes;
  int stages = (int)((227u * 1024u - 2048u) / stage_bytes);
  if (stages > WF_MAX_STAGES) stages = WF_MAX_STAGES;
  p.stages = stages;
  p.tmem_cols = (uint32_t)pow2_ceil_i(4 * p.n_tile);

  const long long x_elems = (long long)N * C * H * W, dy_elems = (long long)N * K * OH * OW;
  void *x_t = nullptr, *dy_t = nullptr;
  float* ws = nullptr;
  if (!x_is_staged && pool_alloc(&x_t, (uint64_t)x_elems * 4 + 64)) return 1;
  if (!dy_is_staged && pool_alloc(&dy_t, (uint64_t)dy_elems * 4 + 64)) { pool_free(x_t); return 1; }
  if (pool_alloc((void**)&ws, (uint64_t)splits * RS * C * K * sizeof(float))) { pool_free(x_t); pool_free(dy_t); return 1; }
  p.ws = ws;
  int rc = 0;
  const void* x_s = x_is_staged ? x : x_t;
  const void* dy_s = dy_is_staged ? dyv : dy_t;
  if (!x_is_staged) rc = to_nhwc_f16(x_t, reinterpret_cast<const float*>(x), nullptr, N, C, H * W, nullptr);
  if (rc == 0 && !dy_is_staged) rc = to_nhwc_f16(dy_t, reinterpret_cast<const float*>(dyv), nullptr, N, K, OH * OW, nullptr);
  if (rc == 0) {
    const uint64_t pix = (uint64_t)C * 4u;
    const uint64_t dims[5] = {64u, (uint64_t)W, (uint64_t)H, (uint64_t)(2 * (C / 64)), (uint64_t)N};
    const uint64_t strides[4] = {pix, (uint64_t)W * pix, 128u, (uint64_t)H * W * pix};
    const uint32_t box[5] = {64u, (uint32_t)w_c, (uint32_t)p.rows_c, (uint32_t)(p.Cm / 32), 1u};
    if (encode_tensor_map(&p.tm_x, x_s, 5, dims, strides, box, SWZ_128B, 1)) rc = -1;
  }
  if (rc == 0) {
    const uint64_t pix = (uint64_t)K * 4u;
    const uint64_t dims[5] = {64u, (uint64_t)OW, (uint64_t)OH, (uint64_t)(2 * (K / 64)), (uint64_t)N};
    const uint64_t strides[4] = {pix, (uint64_t)OW * pix, 128u, (uint64_t)OH * OW * pix};
    const uint32_t box[5] = {64u, (uint32_t)w_c, (uint32_t)p.rows_c, (uint32_t)(2 * p.ng), 1u};
    if (encode_tensor_map(&p.tm_dy, dy_s, 5, dims, strides, box, SWZ_128B, 1)) rc = -1;
  }
  if (rc == 0) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(umma_conv_wgrad_f16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(227 * 1024));
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(umma_conv_wgrad_f16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(227 * 1024));
      if (e != cudaSuccess) rc = set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      configured = (e == cudaSuccess);
    }
  }
  if (rc == 0) {
    const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 + 256;
    const int grid = items * splits;
    bool debug;
    p.dbg = debug_words(&debug);
    if (debug)
      fprintf(stderr, "[umma f16] wgrad: grid %d items %d splits %d chunks %d/split Cm %d taps/cta %d n_tile %d stages %d w_c %d rows_c %d cluster %dx%d\n",
              grid, items, splits, p.chunks_per_split, p.Cm, p.taps_per_cta, p.n_tile, stages, w_c, p.rows_c, p.ct, p.cn);
    cudaError_t e;
    if (p.ct * p.cn > 1) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)grid);
      cfg.blockDim = dim3(WF_THREADS);
      cfg.dynamicSmemBytes = smem_bytes;
      cfg.stream = g_stream;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = (unsigned)(p.ct * p.cn);
      attr.val.clusterDim.y = 1;
      attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, umma_conv_wgrad_f16_kernel<true>, p);
      g_launches++;
      if (e == cudaSuccess) e = cudaGetLastError();
    } else {
      NG_LAUNCH(umma_conv_wgrad_f16_kernel<false>, grid, WF_THREADS, smem_bytes, p);
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) rc = set_error("umma_conv_wgrad_f16_kernel launch failed: %s", cudaGetErrorString(e));
    if (debug) {
      cudaError_t e2 = cudaStreamSynchronize(g_stream);
      fprintf(stderr, "[umma f16] wgrad sync: %s ; dbg: %08x %08x %08x\n", cudaGetErrorString(e2), p.dbg[0], p.dbg[1], p.dbg[2]);
    }
  }
  if (rc == 0) {
    const long long n = (long long)RS * C * K;
    int g = (int)ceil_div(n, 256);
    if (g > sms * 8) g = sms * 8;
    NG_LAUNCH(wgrad_reduce_scaled_kernel, g, 256, 0, dw, (const float*)ws, splits, RS, C, K,
              (const float*)staged_meta(const_cast<void*>(x_s), x_elems),
              (const float*)staged_meta(const_cast<void*>(dy_s), dy_elems));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = set_error("wgrad_reduce_scaled_kernel launch failed: %s", cudaGetErrorString(e));
  }
  pool_free(ws);
  pool_free(dy_t);
  pool_free(x_t);
  return rc;
}

}  // namespace umma
}  // namespace ng
#endif  // NG_EMU
