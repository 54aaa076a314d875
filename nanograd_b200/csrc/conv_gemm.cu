// Implicit-GEMM convolution (fprop, dgrad, wgrad) and GEMM on the FP32 SIMT contraction core
// (contract.cuh).  NCHW activations, KCRS filters, exactly the reference's layouts.
//
// Replaces, in the reference:
//   conv2d / conv1d OpenCL programs            nanograd/nn/ops_gpu.py:308-374   (K1, K4)
//   conv_backward_x (undefined results there)  ops_gpu.py:685-709, 769-795      (K2, K5)
//   conv_backward_weight                       ops_gpu.py:711-731, 797-820      (K3, K6)
//   matmul + perm                              ops_gpu.py:162-181, 214-232      (K7, K8)
// and states the arithmetic of ops_cpu.py:193-204 (MatMul), 290-342 (Conv1d), 345-400 (Conv2d).
// Padding is applied by predicated loads (no padded copy: tensor.py:419-429 disappears) and the
// bias add of module.py:434,476 rides in the epilogue.
#include "contract.cuh"
#include "umma_common.cuh"

namespace ng {

// =====================================================================================
// fprop / dgrad: gathered activations x dense filter matrix
// =====================================================================================
struct ConvGatherP {
  const float* a;     // gathered tensor: x (fprop) or dy (dgrad)
  const float* b;     // filters
  const float* bias;  // per output channel or null
  float* out;
  int M;              // images * opix
  int J;              // output channels of this contraction (K for fprop, C for dgrad)
  int KC;             // reduction channels (C for fprop, K for dgrad)
  int RS;             // filter taps
  int opix, ow_out;   // output pixels per image, output row width
  int a_chan_stride;  // H*W (fprop) / OH*OW (dgrad)
  long long a_img_stride;
  int a_h, a_w;       // spatial bounds of the gathered tensor
  int y_mul, y_add, x_mul, x_add;  // y0 = orow*y_mul + y_add
  int dstride;        // dgrad with conv stride > 1: tap position must divide by it
  long long b_sj, b_sc;  // filter address = j*b_sj + kc*b_sc + rs
  int vec_store;      // opix % 4 == 0 and out 16B-aligned
  int relu;
  short tap_dy[64], tap_dx[64];
};

// Reduction order of the gather loaders:
//   GATHER_GENERIC / GATHER_STRIDED  k = kc*RS + rs (any channel count; STRIDED = dgrad of a
//                                    strided convolution, tap positions must divide by the stride)
//   GATHER_TAPMAJOR                  k = rs*KC + kc, KC % 8 == 0: a K-tile is 8 channels of ONE
//                                    filter tap, so bounds are checked once per pixel per K-tile
//                                    (a precomputed tap-validity bit) and addresses step by the
//                                    channel stride.
enum { GATHER_GENERIC = 0, GATHER_STRIDED = 1, GATHER_TAPMAJOR = 2 };

template <class T, int MODE>
struct ConvGatherLoader {
  using AM = CtAMapI<T>;
  const ConvGatherP& p;
  const float* a_pix[AM::NP];  // &a[img][0][y0][x0] (dereferenced only where the tap is in bounds)
  int y0[AM::NP], x0[AM::NP], il[AM::NP];
  unsigned long long tapmask[AM::NP];  // TAPMAJOR: bit rs = tap rs in bounds for this pixel (0 if the column is padding)
  bool i_valid[AM::NP];
  int krow0;
  float ra[T::EA];
  // B: elements (klane, j_local + 32q)
  const float* b_col[T::EB];
  bool b_valid[T::EB];
  int klane, j_local;
  CtKPos bpos;   // GENERIC: position of row klane
  CtKPos kbase;  // GENERIC: position of row 0 of the current K-tile
  int cur_rs, cur_c0;  // TAPMAJOR
  float rb[T::EB];

  __device__ __forceinline__ ConvGatherLoader(const ConvGatherP& pp, int tile_i0, int tile_j0, int tid) : p(pp) {
    krow0 = AM::row0(tid);
#pragma unroll
    for (int q = 0; q < AM::NP; ++q) {
      il[q] = AM::col(tid, q);
      const int i = tile_i0 + il[q];
      i_valid[q] = i < p.M;
      const int ii = i_valid[q] ? i : 0;
      const int n_img = ii / p.opix, pix = ii - n_img * p.opix;
      const int orow = pix / p.ow_out, ocol = pix - orow * p.ow_out;
      y0[q] = orow * p.y_mul + p.y_add;
      x0[q] = ocol * p.x_mul + p.x_add;
      a_pix[q] = p.a + (long long)n_img * p.a_img_stride + (long long)y0[q] * p.a_w + x0[q];
      tapmask[q] = 0ull;
      if (MODE == GATHER_TAPMAJOR && i_valid[q]) {
        for (int t = 0; t < p.RS; ++t) {
          const int y = y0[q] + p.tap_dy[t], x = x0[q] + p.tap_dx[t];
          if ((unsigned)y < (unsigned)p.a_h && (unsigned)x < (unsigned)p.a_w) tapmask[q] |= (1ull << t);
        }
      }
    }
    klane = tid & 7;
    j_local = tid >> 3;
    bpos.set(klane, p.RS);
    kbase.set(0, p.RS);
    cur_rs = 0;
    cur_c0 = 0;
#pragma unroll
    for (int q = 0; q < T::EB; ++q) {
      const int jl = j_local + 32 * q;
      const int j = tile_j0 + jl;
      b_valid[q] = (jl < T::BN) && (j < p.J);
      b_col[q] = p.b + (long long)(b_valid[q] ? j : 0) * p.b_sj;
    }
  }

  __device__ __forceinline__ void fetch() {
    if (MODE == GATHER_TAPMAJOR) {
      const int toff = p.tap_dy[cur_rs] * p.a_w + p.tap_dx[cur_rs];
      const long long coff = (long long)(cur_c0 + krow0) * p.a_chan_stride + toff;
#pragma unroll
      for (int q = 0; q < AM::NP; ++q) {
        const bool ok = (tapmask[q] >> cur_rs) & 1ull;
        const float* src = a_pix[q] + coff;
#pragma unroll
        for (int r = 0; r < AM::KP; ++r)
          ra[q * AM::KP + r] = ok ? src[(long long)(r * AM::KSTEP) * p.a_chan_stride] : 0.f;
      }
      const long long boff = (long long)(cur_c0 + klane) * p.b_sc + cur_rs;
#pragma unroll
      for (int q = 0; q < T::EB; ++q) rb[q] = b_valid[q] ? b_col[q][boff] : 0.f;
      cur_c0 += CT_BK;
      if (cur_c0 >= p.KC) { cur_c0 = 0; ++cur_rs; }
    } else {
      CtKPos kp = kbase;
      kp.advance(krow0, p.RS);
#pragma unroll
      for (int r = 0; r < AM::KP; ++r) {
        const bool k_ok = kp.slow < p.KC;
        const int dy = p.tap_dy[kp.fast], dx = p.tap_dx[kp.fast];
#pragma unroll
        for (int q = 0; q < AM::NP; ++q) {
          float v = 0.f;
          if (i_valid[q] && k_ok) {
            int y = y0[q] + dy, x = x0[q] + dx;
            if (MODE == GATHER_STRIDED) {
              const bool div_ok = (y % p.dstride == 0) && (x % p.dstride == 0);
              y /= p.dstride;
              x /= p.dstride;
              if (div_ok && (unsigned)y < (unsigned)p.a_h && (unsigned)x < (unsigned)p.a_w)
                v = (a_pix[q] - ((long long)y0[q] * p.a_w + x0[q]))[(long long)kp.slow * p.a_chan_stride + y * p.a_w + x];
            } else {
              if ((unsigned)y < (unsigned)p.a_h && (unsigned)x < (unsigned)p.a_w)
                v = a_pix[q][(long long)kp.slow * p.a_chan_stride + dy * p.a_w + dx];
            }
          }
          ra[q * AM::KP + r] = v;
        }
        kp.advance(AM::KSTEP, p.RS);
      }
      kbase.advance(CT_BK, p.RS);
#pragma unroll
      for (int q = 0; q < T::EB; ++q) {
        float v = 0.f;
        if (b_valid[q] && bpos.slow < p.KC) v = b_col[q][(long long)bpos.slow * p.b_sc + bpos.fast];
        rb[q] = v;
      }
      bpos.advance(CT_BK, p.RS);
    }
  }

  __device__ __forceinline__ void commit(float* As, float* Bs) {
#pragma unroll
    for (int q = 0; q < AM::NP; ++q)
#pragma unroll
      for (int r = 0; r < AM::KP; ++r) As[(krow0 + r * AM::KSTEP) * T::LDA + il[q]] = ra[q * AM::KP + r];
#pragma unroll
    for (int q = 0; q < T::EB; ++q) {
      const int jl = j_local + 32 * q;
      if (jl < T::BN) Bs[klane * T::LDB + jl] = rb[q];
    }
  }
};

template <class T, int MODE>
__global__ void __launch_bounds__(CT_THREADS, 2) conv_gather_kernel(const __grid_constant__ ConvGatherP p) {
  NG_SHARED16 float smem[T::TILE_FLOATS];
  const int tid = threadIdx.x;
  const int tile_i0 = blockIdx.x * T::BM, tile_j0 = blockIdx.y * T::BN;
  const CtCoord c = ct_coord<T>(tid);
  float acc[8][T::TN];
  ct_zero<T>(acc);
  {
    ConvGatherLoader<T, MODE> ld(p, tile_i0, tile_j0, tid);
    const int num_kt = (p.KC * p.RS + CT_BK - 1) / CT_BK;
    ct_mainloop<T>(ld, num_kt, c, acc, smem);
  }
  // epilogue: out[(img*J + j)*opix + pix] (+ bias[j]) ; I = img*opix + pix is contiguous per (img, j)
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    const int i = tile_i0 + c.i0 + 32 * f;
    if (i >= p.M) continue;
    if (p.vec_store) {
      const int n_img = i / p.opix, pix = i - n_img * p.opix;
#pragma unroll
      for (int n = 0; n < T::TN; ++n) {
        const int j = tile_j0 + ct_jcol<T>(c, n);
        if (j >= p.J) continue;
        const float bv = p.bias ? p.bias[j] : 0.f;
        float4 v;
        v.x = acc[4 * f + 0][n] + bv; v.y = acc[4 * f + 1][n] + bv;
        v.z = acc[4 * f + 2][n] + bv; v.w = acc[4 * f + 3][n] + bv;
        if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *reinterpret_cast<float4*>(p.out + ((long long)n_img * p.J + j) * p.opix + pix) = v;
      }
    } else {
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int im = i + m;
        if (im >= p.M) continue;
        const int n_img = im / p.opix, pix = im - n_img * p.opix;
#pragma unroll
        for (int n = 0; n < T::TN; ++n) {
          const int j = tile_j0 + ct_jcol<T>(c, n);
          if (j >= p.J) continue;
          float v = acc[4 * f + m][n] + (p.bias ? p.bias[j] : 0.f);
          if (p.relu) v = fmaxf(v, 0.f);
          p.out[((long long)n_img * p.J + j) * p.opix + pix] = v;
        }
      }
    }
  }
}

// =====================================================================================
// fprop for the image-facing layer (C <= 4, 3x3, stride 1): with C*9 <= 36 reduction terms the
// implicit GEMM spends most of its time staging.  Here a thread owns a strip of 4 output pixels,
// holds its C x 3 x 6 input window in registers and walks the output channels: per channel the
// filter row comes from shared memory as warp-uniform LDS.128, C*36 FFMA follow, and the warp
// stores 512 contiguous bytes of one (image, channel) plane.  blockIdx.y splits the channels so
// that small batches still fill the machine.
// =====================================================================================
struct FpropSmallP {
  const float* x;
  const float* w;
  const float* bias;
  float* out;
  int N, H, W, K, OH, OW, pad_t, pad_l;
  int strips, kper, relu;
};

template <int C>
__global__ void __launch_bounds__(256) conv_fprop_smallc_kernel(const __grid_constant__ FpropSmallP p) {
  constexpr int CRS = C * 9, WP = (CRS + 3) & ~3;
  NG_DYN_SMEM(float, wsm);  // [kn][WP]
  const int tid = threadIdx.x;
  const int k0 = blockIdx.y * p.kper;
  const int kn = p.K - k0 < p.kper ? p.K - k0 : p.kper;
  for (int e = tid; e < kn * WP; e += 256) {
    const int kk = e / WP, j = e - kk * WP;
    wsm[e] = j < CRS ? p.w[(long long)(k0 + kk) * CRS + j] : 0.f;
  }
  __syncthreads();
  const int strip = blockIdx.x * 256 + tid;
  if (strip >= p.strips) return;
  const int ow4 = p.OW >> 2, per_img = p.OH * ow4;
  const int n = strip / per_img, rem = strip - n * per_img;
  const int oh = rem / ow4, ow0 = (rem - oh * ow4) * 4;
  float xv[C][3][6];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = oh + r - p.pad_t;
      const float* row = p.x + ((long long)(n * C + c) * p.H + ih) * p.W;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int iw = ow0 + j - p.pad_l;
        xv[c][r][j] = (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) ? row[iw] : 0.f;
      }
    }
  const long long plane = (long long)p.OH * p.OW;
  float* o = p.out + ((long long)n * p.K + k0) * plane + (long long)oh * p.OW + ow0;
  for (int kk = 0; kk < kn; ++kk) {
    float wv[WP];
#pragma unroll
    for (int j = 0; j < WP; j += 4) {
      const float4 t = *reinterpret_cast<const float4*>(wsm + kk * WP + j);
      wv[j] = t.x; wv[j + 1] = t.y; wv[j + 2] = t.z; wv[j + 3] = t.w;
    }
    const float b = p.bias ? p.bias[k0 + kk] : 0.f;
    float4 a;
    a.x = a.y = a.z = a.w = b;
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const float wgt = wv[(c * 3 + r) * 3 + s];
          a.x = fmaf(wgt, xv[c][r][s], a.x);
          a.y = fmaf(wgt, xv[c][r][s + 1], a.y);
          a.z = fmaf(wgt, xv[c][r][s + 2], a.z);
          a.w = fmaf(wgt, xv[c][r][s + 3], a.w);
        }
    if (p.relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
    *reinterpret_cast<float4*>(o + kk * plane) = a;
  }
}

// =====================================================================================
// Direct fprop / dgrad for narrow layers (both channel counts <= 32, stride 1): BASELINE configs[1] at C = K = 16 / 32.
// The implicit GEMM has J <= 32 output columns there (8 x 1..4 register micro-tiles: one FFMA per shared-memory
// read) and the tensor-core kernels spend their time filling and draining a pipeline that is 9 short stages deep.
// Here a thread owns a strip of 4 output pixels x JT output channels (JT x 4 accumulators), walks the reduction
// channels and filter rows, keeps the S + 3 input values of a row segment in registers and reads the filter taps as
// warp-uniform LDS.128 from a [kc][tap][JT] copy in shared memory: S * JT * 4 FFMA per (channel, row) for S + 3
// loads and S * JT / 4 LDS.128.  dgrad is the same correlation with the filter flipped (FLIP) and transposed by the
// (b_sj, b_sc) strides.
// =====================================================================================
template <int JT, int S, bool FLIP>
__global__ void __launch_bounds__(128) conv_direct_kernel(const __grid_constant__ ConvGatherP p, int n_img, int OH, int OW,
                                                          int strips_per_row, long long total_strips, int vec_ok) {
  // square S x S filters; blockIdx.y = tile of JT output channels (a grid dimension rather than a loop: the strips
  // alone leave a 256 x 30 x 30 output at 13 warps per SM)
  NG_DYN_SMEM(float, wsm);   // [KC][S*S][JT]
  constexpr int RS = S * S;
  const int opix = OH * OW;
  const int j0 = blockIdx.y * JT;
  for (int e = threadIdx.x; e < p.KC * RS * JT; e += blockDim.x) {
    const int j = e % JT, t = e / JT;
    const int rs = t % RS, kc = t / RS;
    wsm[e] = (j0 + j < p.J) ? p.b[(long long)(j0 + j) * p.b_sj + (long long)kc * p.b_sc + rs] : 0.f;
  }
  __syncthreads();
  for (long long strip = (long long)blockIdx.x * blockDim.x + threadIdx.x; strip < total_strips;
       strip += (long long)gridDim.x * blockDim.x) {
    const int sx = (int)(strip % strips_per_row);
    const long long t = strip / strips_per_row;
    const int oy = (int)(t % OH), n = (int)(t / OH);
    const int ox0 = sx * 4;
    float acc[JT][4];
#pragma unroll
    for (int j = 0; j < JT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const float* img = p.a + (long long)n * p.a_img_stride;
    // input columns of a row's taps: ox + x_add + dx, dx = s (fprop) or -s (dgrad, FLIP): S + 3 values per row
    const int ix0 = ox0 + p.x_add + (FLIP ? -(S - 1) : 0);
    const int iy0 = oy + p.y_add + (FLIP ? -(S - 1) : 0);   // smallest input row; filter row r reads iy0 + (FLIP ? S-1-r : r)
    bool colok[S + 3];
#pragma unroll
    for (int i = 0; i < S + 3; ++i) colok[i] = (ix0 + i >= 0) && (ix0 + i < p.a_w);
    // widest aligned loads of a window row: every strip starts at 4 * sx + const, so the mode is uniform
    const int ix_const = p.x_add + (FLIP ? -(S - 1) : 0);
    const int vmode = !vec_ok ? 0 : ((ix_const & 3) == 0 && (p.a_w & 3) == 0 && vec_ok == 2) ? 1
                      : ((ix_const & 1) == 0 && (p.a_w & 1) == 0) ? 2 : 0;
    for (int kc = 0; kc < p.KC; ++kc) {
      const float* plane = img + (long long)kc * p.a_chan_stride;
      // the S rows of this channel's window: all loads issued before the first use
      float in[S][S + 3];
#pragma unroll
      for (int rr = 0; rr < S; ++rr) {
        const int iy = iy0 + rr;
        const bool rowok = iy >= 0 && iy < p.a_h;
        const float* rowp = plane + (long long)iy * p.a_w + ix0;
        if (vmode == 1) {
          // ix0 % 4 == 0 and a_w % 4 == 0: 128-bit pieces (+ a 64-bit tail), each wholly inside or outside the row
#pragma unroll
          for (int i = 0; i + 4 <= S + 3; i += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rowok && colok[i]) v = *reinterpret_cast<const float4*>(rowp + i);
            in[rr][i] = v.x; in[rr][i + 1] = v.y; in[rr][i + 2] = v.z; in[rr][i + 3] = v.w;
          }
          if ((S + 3) % 4 != 0) {
            float2 v = make_float2(0.f, 0.f);
            if (rowok && colok[S + 1]) v = *reinterpret_cast<const float2*>(rowp + S + 1);
            in[rr][S + 1] = v.x; in[rr][S + 2] = v.y;
          }
        } else if (vmode == 2) {
          // ix0 and a_w even: 64-bit pieces
#pragma unroll
          for (int i = 0; i < S + 3; i += 2) {
            float2 v = make_float2(0.f, 0.f);
            if (rowok && colok[i]) v = *reinterpret_cast<const float2*>(rowp + i);
            in[rr][i] = v.x; in[rr][i + 1] = v.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < S + 3; ++i) in[rr][i] = (rowok && colok[i]) ? rowp[i] : 0.f;
        }
      }
#pragma unroll
      for (int r = 0; r < S; ++r) {
        const int rr = FLIP ? (S - 1 - r) : r;
        const float4* wrow = reinterpret_cast<const float4*>(wsm + (kc * RS + r * S) * JT);
#pragma unroll
        for (int s2 = 0; s2 < S; ++s2) {
          const int off = FLIP ? (S - 1 - s2) : s2;
#pragma unroll
          for (int j4 = 0; j4 < JT / 4; ++j4) {
            const float4 w = wrow[s2 * (JT / 4) + j4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              acc[4 * j4 + 0][q] = fmaf(in[rr][off + q], w.x, acc[4 * j4 + 0][q]);
              acc[4 * j4 + 1][q] = fmaf(in[rr][off + q], w.y, acc[4 * j4 + 1][q]);
              acc[4 * j4 + 2][q] = fmaf(in[rr][off + q], w.z, acc[4 * j4 + 2][q]);
              acc[4 * j4 + 3][q] = fmaf(in[rr][off + q], w.w, acc[4 * j4 + 3][q]);
            }
          }
        }
      }
    }
    float* o = p.out + ((long long)n * p.J + j0) * opix + (long long)oy * OW + ox0;
#pragma unroll
    for (int j = 0; j < JT; ++j) {
      if (j0 + j >= p.J) break;
      const float bv = p.bias ? p.bias[j0 + j] : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (ox0 + q < OW) {
          float v = acc[j][q] + bv;
          if (p.relu) v = fmaxf(v, 0.f);
          o[(long long)j * opix + q] = v;
        }
      }
    }
  }
}

// Returns 0 when launched, -1 when the shape is not one the direct kernel takes.
static int launch_conv_direct(const ConvGatherP& p, int R, int S, bool flip, int n_img, int OH, int OW) {
  if (p.y_mul != 1 || p.x_mul != 1 || p.dstride != 1 || R != S) return -1;
  if (p.J > 32 || p.KC > 32 || p.KC < 8 || (S != 3 && S != 5 && S != 1)) return -1;
#ifdef NG_EMU
  const long long min_pixels = 1024;    // the CPU suite's fixture sizes exercise the kernel under the host emulation
#else
  const long long min_pixels = 65536;   // tiny problems keep their own kernels (README CNN)
#endif
  if ((long long)n_img * OH * OW < min_pixels) return -1;
  // the taps must be in storage order (fprop: (dy, dx) = (r, s) + const, dgrad: (-r, -s) + const)
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s)
      if (p.tap_dy[r * S + s] != (flip ? -r : r) || p.tap_dx[r * S + s] != (flip ? -s : s)) return -1;
  const int strips_per_row = (int)ceil_div(OW, 4);
  const long long total = (long long)n_img * OH * strips_per_row;
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  // output-channel tile: the widest that still gives ~24 warps per SM (5 x 5 filters keep the window small)
  int JT = p.J <= 8 ? 8 : (p.J <= 16 ? 16 : 32);
  if (S == 5 && JT > 16) JT = 16;
  while (JT > 8 && total * ceil_div(p.J, JT) < (long long)sms * 24 * 32) JT >>= 1;
  const size_t smem = sizeof(float) * (size_t)p.KC * R * S * JT;
  if (smem > 48 * 1024) return -1;
  long long g = ceil_div(total, 128);
  if (g > (long long)sms * 16) g = (long long)sms * 16;
  const dim3 grid((unsigned)g, (unsigned)ceil_div(p.J, JT));
  // vector loads of the window rows need the planes' starts aligned: base pointer and channel / image strides
  const int vec_ok = (((uintptr_t)p.a & 15) == 0 && p.a_chan_stride % 4 == 0 && p.a_img_stride % 4 == 0) ? 2
                     : (((uintptr_t)p.a & 7) == 0 && p.a_chan_stride % 2 == 0 && p.a_img_stride % 2 == 0) ? 1 : 0;
#define NG_DIRECT(JT_, S_)                                                                                        \
  do {                                                                                                            \
    if (flip) NG_LAUNCH((conv_direct_kernel<JT_, S_, true>), grid, 128, smem, p, n_img, OH, OW, strips_per_row, total, vec_ok); \
    else NG_LAUNCH((conv_direct_kernel<JT_, S_, false>), grid, 128, smem, p, n_img, OH, OW, strips_per_row, total, vec_ok);     \
  } while (0)
#define NG_DIRECT_S(JT_)                                                                                          \
  do {                                                                                                            \
    if (S == 1) NG_DIRECT(JT_, 1);                                                                                \
    else if (S == 3) NG_DIRECT(JT_, 3);                                                                           \
    else NG_DIRECT(JT_, 5);                                                                                       \
  } while (0)
  if (JT == 8) NG_DIRECT_S(8);
  else if (JT == 16) NG_DIRECT_S(16);
  else NG_DIRECT_S(32);
#undef NG_DIRECT_S
#undef NG_DIRECT
  NG_CHECK_LAUNCH();
  return 0;
}

static int launch_fprop_smallc(const ConvGatherP& g, int N, int C, int H, int W, int K, int OH, int OW, int pad_t,
                               int pad_l) {
  FpropSmallP p;
  memset(&p, 0, sizeof(p));
  p.x = g.a; p.w = g.b; p.bias = g.bias; p.out = g.out;
  p.N = N; p.H = H; p.W = W; p.K = K; p.OH = OH; p.OW = OW; p.pad_t = pad_t; p.pad_l = pad_l;
  p.strips = N * OH * (OW / 4);
  p.relu = g.relu;
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  const int gx = (int)ceil_div(p.strips, 256);
  int ksplit = 1;
  while (gx * ksplit < 6 * sms && K / (ksplit * 2) >= 16) ksplit *= 2;
  p.kper = (int)ceil_div(K, ksplit);
  const int gy = (int)ceil_div(K, p.kper);
  const int wp = (C * 9 + 3) & ~3;
  const size_t smem = sizeof(float) * (size_t)p.kper * wp;
  if (smem > 48 * 1024) return -1;
  dim3 grid((unsigned)gx, (unsigned)gy);
  switch (C) {
    case 1: NG_LAUNCH((conv_fprop_smallc_kernel<1>), grid, 256, smem, p); break;
    case 2: NG_LAUNCH((conv_fprop_smallc_kernel<2>), grid, 256, smem, p); break;
    case 3: NG_LAUNCH((conv_fprop_smallc_kernel<3>), grid, 256, smem, p); break;
    default: NG_LAUNCH((conv_fprop_smallc_kernel<4>), grid, 256, smem, p); break;
  }
  NG_CHECK_LAUNCH();
  return 0;
}

// =====================================================================================
// dense epilogue shared by wgrad and GEMM: out[z][j*ldc + i]
// =====================================================================================
template <class T>
__device__ __forceinline__ void dense_store(float* out, long long ldc, int I, int J, int tile_i0, int tile_j0,
                                            const CtCoord& c, const float (&acc)[8][T::TN], const float* bias_i,
                                            int relu, int vec_store) {
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    const int i = tile_i0 + c.i0 + 32 * f;
    if (i >= I) continue;
#pragma unroll
    for (int n = 0; n < T::TN; ++n) {
      const int j = tile_j0 + ct_jcol<T>(c, n);
      if (j >= J) continue;
      float* dst = out + (long long)j * ldc + i;
      if (vec_store && i + 3 < I) {
        float4 v;
        v.x = acc[4 * f + 0][n]; v.y = acc[4 * f + 1][n]; v.z = acc[4 * f + 2][n]; v.w = acc[4 * f + 3][n];
        if (bias_i) { v.x += bias_i[i]; v.y += bias_i[i + 1]; v.z += bias_i[i + 2]; v.w += bias_i[i + 3]; }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *reinterpret_cast<float4*>(dst) = v;
      } else {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          if (i + m < I) {
            float v = acc[4 * f + m][n] + (bias_i ? bias_i[i + m] : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            dst[m] = v;
          }
        }
      }
    }
  }
}

// out[e] = sum_s ws[s][e] (+ bias_i[e % I]) — second pass of the deterministic split-K
__global__ void __launch_bounds__(256) splitk_reduce_kernel(float* __restrict__ out, const float* __restrict__ ws,
                                                            int S, long long n, const float* __restrict__ bias_i,
                                                            int I, int relu) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += ws[(long long)s * n + e];
    if (bias_i) acc += bias_i[e % I];
    if (relu) acc = fmaxf(acc, 0.f);
    out[e] = acc;
  }
}

// Many partials, few outputs (the direct first-layer wgrad writes one partial per CTA): 8 threads share
// an output, each sums every 8th partial, shared memory adds the 8 in a fixed order.
__global__ void __launch_bounds__(512) splitk_reduce_wide_kernel(float* __restrict__ out, const float* __restrict__ ws, int S,
                                                                 long long n) {
  __shared__ float red[8][64];
  const int el = threadIdx.x & 63, g = threadIdx.x >> 6;
  const long long e = (long long)blockIdx.x * 64 + el;
  float acc = 0.f;
  if (e < n)
    for (int s = g; s < S; s += 8) acc += ws[(long long)s * n + e];
  red[g][el] = acc;
  __syncthreads();
  if (g == 0 && e < n) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][el];
    out[e] = t;
  }
}

// ---- direct wgrad for tiny layers (the README CNN: C*R*S*K <= 2048 filter taps, one image's x and dy fit
// shared memory).  The implicit-GEMM kernel runs these as hundreds of 8-tile splits whose every step is
// an exposed load latency (32 us for conv1 at batch 128).  Here a CTA stages whole images, thread
// (group g, tap o) accumulates dW[o] over the output rows oh = g (mod G) in registers — dy is a warp
// broadcast, x a conflict-free shifted read — and one partial per CTA goes to the fixed-order reduce.
struct WgradTinyP {
  const float* x;
  const float* dy;
  float* ws;      // [grid][K*C*R*S]
  int N, C, H, W, K, R, S, OH, OW, stride, pad_t, pad_l;
  int O, G;       // O = K*C*R*S taps, G = row groups sharing a tap
};

__global__ void __launch_bounds__(256) conv_wgrad_tiny_kernel(const __grid_constant__ WgradTinyP p) {
  NG_DYN_SMEM(float, sm);
  float* xs = sm;                                // [C][H][W]
  float* ds = sm + p.C * p.H * p.W;              // [K][OH][OW]
  float* red = ds + p.K * p.OH * p.OW;           // [G][O] (G > 1 only)
  const int tid = threadIdx.x, T = blockDim.x;
  const int x_elems = p.C * p.H * p.W, d_elems = p.K * p.OH * p.OW;
  constexpr int kMaxPer = 8;                     // taps per thread when O > blockDim.x
  float acc[kMaxPer];
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) acc[j] = 0.f;
  const int g = p.G > 1 ? tid / p.O : 0;
  const bool active = p.G > 1 ? (tid < p.G * p.O) : true;
  for (int n = blockIdx.x; n < p.N; n += gridDim.x) {
    __syncthreads();
    const float* xg = p.x + (long long)n * x_elems;
    const float* dg = p.dy + (long long)n * d_elems;
    for (int i = tid; i < x_elems; i += T) xs[i] = xg[i];
    for (int i = tid; i < d_elems; i += T) ds[i] = dg[i];
    __syncthreads();
    if (!active) continue;
#pragma unroll
    for (int j = 0; j < kMaxPer; ++j) {
      const int o = p.G > 1 ? tid - g * p.O : tid + j * T;
      if (o >= p.O || (p.G > 1 && j > 0)) break;
      const int s = o % p.S, r = (o / p.S) % p.R, c = (o / (p.S * p.R)) % p.C, k = o / (p.S * p.R * p.C);
      const float* dk = ds + k * p.OH * p.OW;
      const float* xc = xs + c * p.H * p.W;
      // output columns whose input column ow*stride + s - pad_l lies inside the image
      int ow_lo = p.pad_l - s > 0 ? (p.pad_l - s + p.stride - 1) / p.stride : 0;
      int ow_hi = (p.W - 1 - s + p.pad_l) >= 0 ? (p.W - 1 - s + p.pad_l) / p.stride + 1 : 0;
      if (ow_hi > p.OW) ow_hi = p.OW;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int oh = g; oh < p.OH; oh += p.G) {
        const int ih = oh * p.stride + r - p.pad_t;
        if (ih < 0 || ih >= p.H) continue;
        const float* drow = dk + oh * p.OW;
        const float* xrow = xc + ih * p.W + (s - p.pad_l);
        int ow = ow_lo;
        for (; ow + 4 <= ow_hi; ow += 4) {
          a0 = fmaf(drow[ow + 0], xrow[(ow + 0) * p.stride], a0);
          a1 = fmaf(drow[ow + 1], xrow[(ow + 1) * p.stride], a1);
          a2 = fmaf(drow[ow + 2], xrow[(ow + 2) * p.stride], a2);
          a3 = fmaf(drow[ow + 3], xrow[(ow + 3) * p.stride], a3);
        }
        for (; ow < ow_hi; ++ow) a0 = fmaf(drow[ow], xrow[ow * p.stride], a0);
      }
      acc[j] += (a0 + a1) + (a2 + a3);
    }
  }
  float* out = p.ws + (long long)blockIdx.x * p.O;
  if (p.G > 1) {
    __syncthreads();
    if (active) red[g * p.O + (tid - g * p.O)] = acc[0];
    __syncthreads();
    if (tid < p.O) {
      float t = 0.f;
      for (int q = 0; q < p.G; ++q) t += red[q * p.O + tid];
      out[tid] = t;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kMaxPer; ++j) {
      const int o = tid + j * T;
      if (o < p.O) out[o] = acc[j];
    }
  }
}

// =====================================================================================
// wgrad: dW[k][c][r][s] = sum_{n,oh,ow} dy[n,k,oh,ow] * x[n,c,oh*st+r-pt,ow*st+s-pl]
//   I = flattened (c,r,s), J = k, reduction over (n, oh, ow) split across grid.z
// =====================================================================================
struct WgradP {
  const float* x;
  const float* dy;
  float* out;  // dW or the split-K workspace ([z][K][CRS])
  int N, C, H, W, K, R, S, OH, OW, stride, pad_t, pad_l;
  int CRS;
  long long total_kk;  // N*OH*OW
  int kt_per_split;
  int vec_store;
};

template <class T>
struct WgradLoader {
  const WgradP& p;
  int klane, il;  // this thread stages (klane, il + 32q) of both tiles
  // A (x gather): per q the filter position of column i
  int a_off[T::EA], a_dy[T::EA], a_dx[T::EA];
  bool a_valid[T::EA];
  // B (dy): per q the output channel
  long long b_off[T::EB];
  bool b_valid[T::EB];
  // reduction position (n, oh, ow) of row klane
  int n, oh, ow;
  long long kk, kk_end;
  float ra[T::EA], rb[T::EB];

  __device__ __forceinline__ WgradLoader(const WgradP& pp, int tile_i0, int tile_j0, int tid, long long kk_begin,
                                         long long kk_stop)
      : p(pp) {
    klane = tid & 7;
    il = tid >> 3;
    const int RS = p.R * p.S;
#pragma unroll
    for (int q = 0; q < T::EA; ++q) {
      const int i = tile_i0 + il + 32 * q;
      a_valid[q] = i < p.CRS;
      const int ii = a_valid[q] ? i : 0;
      const int cc = ii / RS, rs = ii - cc * RS;
      const int r = rs / p.S, s = rs - r * p.S;
      a_dy[q] = r - p.pad_t;
      a_dx[q] = s - p.pad_l;
      a_off[q] = cc * p.H * p.W;
    }
#pragma unroll
    for (int q = 0; q < T::EB; ++q) {
      const int jl = il + 32 * q;
      const int j = tile_j0 + jl;
      b_valid[q] = (jl < T::BN) && (j < p.K);
      b_off[q] = (long long)(b_valid[q] ? j : 0) * p.OH * p.OW;
    }
    kk = kk_begin + klane;
    kk_end = kk_stop;
    const int opix = p.OH * p.OW;
    n = (int)(kk / opix);
    const int pix = (int)(kk - (long long)n * opix);
    oh = pix / p.OW;
    ow = pix - oh * p.OW;
  }

  __device__ __forceinline__ void fetch() {
    const bool k_ok = kk < kk_end;
    const int yb = oh * p.stride, xb = ow * p.stride;
    const float* x_img = p.x + (long long)n * p.C * p.H * p.W;
#pragma unroll
    for (int q = 0; q < T::EA; ++q) {
      float v = 0.f;
      const int y = yb + a_dy[q], x = xb + a_dx[q];
      if (k_ok && a_valid[q] && (unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W)
        v = x_img[a_off[q] + y * p.W + x];
      ra[q] = v;
    }
    const float* dy_img = p.dy + (long long)n * p.K * p.OH * p.OW + oh * p.OW + ow;
#pragma unroll
    for (int q = 0; q < T::EB; ++q) {
      float v = 0.f;
      if (k_ok && b_valid[q]) v = dy_img[b_off[q]];
      rb[q] = v;
    }
    kk += CT_BK;
    ow += CT_BK;
    while (ow >= p.OW) { ow -= p.OW; ++oh; }
    while (oh >= p.OH) { oh -= p.OH; ++n; }
  }

  __device__ __forceinline__ void commit(float* As, float* Bs) {
#pragma unroll
    for (int q = 0; q < T::EA; ++q) As[klane * T::LDA + il + 32 * q] = ra[q];
#pragma unroll
    for (int q = 0; q < T::EB; ++q) {
      const int jl = il + 32 * q;
      if (jl < T::BN) Bs[klane * T::LDB + jl] = rb[q];
    }
  }
};

template <class T>
__global__ void __launch_bounds__(CT_THREADS, 2) conv_wgrad_kernel(const __grid_constant__ WgradP p) {
  NG_SHARED16 float smem[T::TILE_FLOATS];
  const int tid = threadIdx.x;
  const int tile_i0 = blockIdx.x * T::BM, tile_j0 = blockIdx.y * T::BN;
  const CtCoord c = ct_coord<T>(tid);
  const long long kk_begin = (long long)blockIdx.z * p.kt_per_split * CT_BK;
  long long kk_stop = kk_begin + (long long)p.kt_per_split * CT_BK;
  if (kk_stop > p.total_kk) kk_stop = p.total_kk;
  const int num_kt = (int)((kk_stop - kk_begin + CT_BK - 1) / CT_BK);
  float acc[8][T::TN];
  ct_zero<T>(acc);
  {
    WgradLoader<T> ld(p, tile_i0, tile_j0, tid, kk_begin, kk_stop);
    ct_mainloop<T>(ld, num_kt, c, acc, smem);
  }
  float* out = p.out + (long long)blockIdx.z * p.K * p.CRS;
  dense_store<T>(out, p.CRS, p.CRS, p.K, tile_i0, tile_j0, c, acc, nullptr, 0, p.vec_store);
}

// =====================================================================================
// wgrad for the image-facing layer (C*R <= 12 filter rows, S == 3, stride 1): the implicit GEMM
// above has only C*R*S = 27 useful columns in a 64-wide tile, so this case gets a direct kernel.
// A CTA stages TR output rows of one image — dy as [K][TR*OW (+4 pad)] and x with its halo as
// [C][TR+R-1][OW+4], shifted by pad_l so that column ow+s of a strip starts 16-byte aligned — and
// thread (k, g) keeps dW[k][row][0..2] for its RPG filter rows (row = c*R + r) in registers:
// per 4-pixel strip one LDS.128 of dy and two warp-uniform LDS.128 of x per row feed 12 FFMA.
// CTAs walk tiles grid-strided and write one partial per CTA; splitk_reduce_kernel sums them
// in a fixed order.
// =====================================================================================
struct WgradSmallP {
  const float* x;
  const float* dy;
  float* ws;  // [gridDim.x][K][C*R*3]
  int N, C, H, W, K, R, OH, OW, pad_t, pad_l;
  int TR, row_blocks, tiles, NR, xrows, xpitch, dpitch;
};

template <int RPG>
__global__ void __launch_bounds__(1024) conv_wgrad_smallc_kernel(const __grid_constant__ WgradSmallP p) {
  NG_DYN_SMEM(float, sm);
  float* dys = sm;
  float* xs = sm + p.K * p.dpitch;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int k = tid % p.K, g = tid / p.K;
  const int PX = p.TR * p.OW, PX4 = PX >> 2;
  int xoff[RPG];
  bool live[RPG];
  float acc[RPG][3];
#pragma unroll
  for (int q = 0; q < RPG; ++q) {
    const int row = g * RPG + q;
    live[q] = row < p.NR;
    const int c = live[q] ? row / p.R : 0, r = live[q] ? row % p.R : 0;
    xoff[q] = (c * p.xrows + r) * p.xpitch;
    acc[q][0] = acc[q][1] = acc[q][2] = 0.f;
  }
  const int xtile = p.C * p.xrows * p.xpitch;
  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const int n = tile / p.row_blocks, oh0 = (tile - n * p.row_blocks) * p.TR;
    const int nvalid = (p.OH - oh0 < p.TR ? p.OH - oh0 : p.TR) * p.OW;
    __syncthreads();  // the previous tile has been consumed
    for (int e = tid; e < p.K * PX4; e += nthr) {
      const int kk = e / PX4, q4 = (e - kk * PX4) * 4;
      float4 v;
      v.x = v.y = v.z = v.w = 0.f;
      if (q4 < nvalid)
        v = *reinterpret_cast<const float4*>(p.dy + ((long long)(n * p.K + kk) * p.OH + oh0) * p.OW + q4);
      *reinterpret_cast<float4*>(dys + kk * p.dpitch + q4) = v;
    }
    for (int e = tid; e < xtile; e += nthr) {
      const int j = e % p.xpitch, t = e / p.xpitch;
      const int rr = t % p.xrows, c = t / p.xrows;
      const int ih = oh0 + rr - p.pad_t, iw = j - p.pad_l;
      float v = 0.f;
      if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) v = p.x[((long long)(n * p.C + c) * p.H + ih) * p.W + iw];
      xs[e] = v;
    }
    __syncthreads();
    const float* drow = dys + k * p.dpitch;
    for (int tr = 0; tr < p.TR; ++tr) {
      for (int ow0 = 0; ow0 < p.OW; ow0 += 4) {
        const float4 d = *reinterpret_cast<const float4*>(drow + tr * p.OW + ow0);
#pragma unroll
        for (int q = 0; q < RPG; ++q) {
          if (!live[q]) continue;
          const float* xr = xs + xoff[q] + tr * p.xpitch + ow0;
          const float4 a = *reinterpret_cast<const float4*>(xr);
          const float4 b = *reinterpret_cast<const float4*>(xr + 4);
          acc[q][0] = fmaf(d.x, a.x, fmaf(d.y, a.y, fmaf(d.z, a.z, fmaf(d.w, a.w, acc[q][0]))));
          acc[q][1] = fmaf(d.x, a.y, fmaf(d.y, a.z, fmaf(d.z, a.w, fmaf(d.w, b.x, acc[q][1]))));
          acc[q][2] = fmaf(d.x, a.z, fmaf(d.y, a.w, fmaf(d.z, b.x, fmaf(d.w, b.y, acc[q][2]))));
        }
      }
    }
  }
  float* out = p.ws + ((long long)blockIdx.x * p.K + k) * (p.NR * 3);
#pragma unroll
  for (int q = 0; q < RPG; ++q) {
    if (!live[q]) continue;
    const int row = g * RPG + q;
    out[row * 3 + 0] = acc[q][0];
    out[row * 3 + 1] = acc[q][1];
    out[row * 3 + 2] = acc[q][2];
  }
}

// =====================================================================================
// GEMM: out[j*ldc + i] = sum_k a(i,k) * b(k,j), a(i,k) = A[i*sai + k*sak], b(k,j) = B[j*sbj + k*sbk]
// (a row-major C[M,N] = A@B maps to I = N, J = M with the operands swapped, see ng_gemm)
// =====================================================================================
struct GemmP {
  const float* a;
  const float* b;
  float* out;
  const float* bias_i;
  int I, J, K;
  long long sai, sak, sbj, sbk, ldc;
  long long batch_a, batch_b, batch_c;  // element strides between batches
  int splits;        // grid.z = batch * splits
  int kt_per_split;
  long long ws_stride;  // elements between split slices of the workspace (I*J*batch) when splits > 1
  int vec_store, relu;
};

template <class T>
struct GemmLoader {
  using AM = CtAMapI<T>;
  const GemmP& p;
  const float* a;
  const float* b;
  int tile_i0, tile_j0, tid;
  int k0, k_end;
  bool a_icontig, b_jcontig;
  float ra[T::EA], rb[T::EB];

  __device__ __forceinline__ GemmLoader(const GemmP& pp, const float* aa, const float* bb, int ti0, int tj0, int t,
                                        int kb, int ke)
      : p(pp), a(aa), b(bb), tile_i0(ti0), tile_j0(tj0), tid(t), k0(kb), k_end(ke) {
    a_icontig = (p.sai == 1);
    b_jcontig = (p.sbj == 1);
  }
  __device__ __forceinline__ void a_coord(int e, int& kl, int& il) const {
    if (a_icontig) { il = AM::col(tid, e / AM::KP); kl = AM::row0(tid) + (e % AM::KP) * AM::KSTEP; }
    else ct_map_kcontig(tid, e, kl, il);
  }
  __device__ __forceinline__ void b_coord(int e, int& kl, int& jl) const {
    if (b_jcontig) { const int x = tid + CT_THREADS * e; jl = x % T::BN; kl = x / T::BN; }
    else ct_map_kcontig(tid, e, kl, jl);
  }
  __device__ __forceinline__ void fetch() {
#pragma unroll
    for (int e = 0; e < T::EA; ++e) {
      int kl, il;
      a_coord(e, kl, il);
      const int i = tile_i0 + il, k = k0 + kl;
      ra[e] = (i < p.I && k < k_end) ? a[(long long)i * p.sai + (long long)k * p.sak] : 0.f;
    }
#pragma unroll
    for (int e = 0; e < T::EB; ++e) {
      int kl, jl;
      b_coord(e, kl, jl);
      const int j = tile_j0 + jl, k = k0 + kl;
      rb[e] = (kl < CT_BK && jl < T::BN && j < p.J && k < k_end)
                  ? b[(long long)j * p.sbj + (long long)k * p.sbk] : 0.f;
    }
    k0 += CT_BK;
  }
  __device__ __forceinline__ void commit(float* As, float* Bs) {
#pragma unroll
    for (int e = 0; e < T::EA; ++e) {
      int kl, il;
      a_coord(e, kl, il);
      As[kl * T::LDA + il] = ra[e];
    }
#pragma unroll
    for (int e = 0; e < T::EB; ++e) {
      int kl, jl;
      b_coord(e, kl, jl);
      if (kl < CT_BK && jl < T::BN) Bs[kl * T::LDB + jl] = rb[e];
    }
  }
};

template <class T>
__global__ void __launch_bounds__(CT_THREADS, 2) gemm_kernel(const __grid_constant__ GemmP p) {
  NG_SHARED16 float smem[T::TILE_FLOATS];
  const int tid = threadIdx.x;
  const int tile_i0 = blockIdx.x * T::BM, tile_j0 = blockIdx.y * T::BN;
  const int batch = blockIdx.z / p.splits, split = blockIdx.z - batch * p.splits;
  const CtCoord c = ct_coord<T>(tid);
  const int kb = split * p.kt_per_split * CT_BK;
  int ke = kb + p.kt_per_split * CT_BK;
  if (ke > p.K) ke = p.K;
  const int num_kt = (ke - kb + CT_BK - 1) / CT_BK;
  float acc[8][T::TN];
  ct_zero<T>(acc);
  {
    GemmLoader<T> ld(p, p.a + batch * p.batch_a, p.b + batch * p.batch_b, tile_i0, tile_j0, tid, kb, ke);
    ct_mainloop<T>(ld, num_kt, c, acc, smem);
  }
  float* out = p.out + (long long)split * p.ws_stride + batch * p.batch_c;
  const bool final_pass = (p.splits == 1);
  dense_store<T>(out, p.ldc, p.I, p.J, tile_i0, tile_j0, c, acc, final_pass ? p.bias_i : nullptr,
                 final_pass ? p.relu : 0, p.vec_store);
}

// ---- tile selection ------------------------------------------------------------------------------
// Tiles are BM (I) x BN (J); the 8 x 8 register micro-tile is kept whenever J > 16.
using TileA = CtTile<2, 4, 8>;  // 128 x 128
using TileB = CtTile<4, 2, 8>;  // 256 x  64
using TileC = CtTile<8, 1, 8>;  // 512 x  32
using TileD = CtTile<2, 4, 4>;  // 128 x  64
using TileE = CtTile<2, 4, 2>;  // 128 x  32
using TileF = CtTile<2, 4, 1>;  // 128 x  16
enum { TILE_A = 0, TILE_B, TILE_C, TILE_D, TILE_E, TILE_F };

struct TileShape { int bm, bn; };
static inline TileShape tile_shape(int t) {
  switch (t) {
    case TILE_A: return {128, 128};
    case TILE_B: return {256, 64};
    case TILE_C: return {512, 32};
    case TILE_D: return {128, 64};
    case TILE_E: return {128, 32};
    default: return {128, 16};
  }
}
// `wide_ok`: the loader supports BM > 128 (gather and GEMM loaders do, wgrad's does not)
// Tiny GEMMs (classifier heads, the README CNN's Linear(400, 10) and its backward products): the tiled
// kernel runs them as one or a few CTAs whose K loop is a chain of exposed global-load latencies (70 us
// for 128x10x400).  Here a thread owns one output element and walks K with four independent partial
// sums; operands come through L1/L2 by their strides, so x @ W.T, dy @ W and dy.T @ x need no transpose.
struct GemmSmallP {
  const float* a;      // op(A)[m, k] = a[m * sam + k * sak]
  const float* b;      // op(B)[k, n] = b[k * sbk + n * sbn]
  float* c;
  const float* bias;   // per output column n, or nullptr
  int M, N, K, relu;
  long long sam, sak, sbk, sbn;
};

// LANES = 1: a thread owns an output and walks K (short reductions).  LANES = 32: a warp owns an output,
// its lanes take every 32nd k (coalesced when the operand is k-contiguous) and a shuffle tree adds them in
// a fixed order (long reductions: 13 steps instead of 400 for the README CNN's Linear(400, 10)).
template <int LANES>
__global__ void __launch_bounds__(128) gemm_small_kernel(const GemmSmallP p) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long e = LANES == 1 ? t : t >> 5;
  if (e >= (long long)p.M * p.N) return;   // LANES = 32: whole warps leave together
  const int lane = LANES == 1 ? 0 : (int)(threadIdx.x & 31);
  const int m = (int)(e / p.N), n = (int)(e - (long long)m * p.N);
  const float* __restrict__ a = p.a + m * p.sam;
  const float* __restrict__ b = p.b + n * p.sbn;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int k = lane;
  for (; k + 3 * LANES < p.K; k += 4 * LANES) {
    s0 = fmaf(a[(long long)(k + 0 * LANES) * p.sak], b[(long long)(k + 0 * LANES) * p.sbk], s0);
    s1 = fmaf(a[(long long)(k + 1 * LANES) * p.sak], b[(long long)(k + 1 * LANES) * p.sbk], s1);
    s2 = fmaf(a[(long long)(k + 2 * LANES) * p.sak], b[(long long)(k + 2 * LANES) * p.sbk], s2);
    s3 = fmaf(a[(long long)(k + 3 * LANES) * p.sak], b[(long long)(k + 3 * LANES) * p.sbk], s3);
  }
  for (; k < p.K; k += LANES) s0 = fmaf(a[(long long)k * p.sak], b[(long long)k * p.sbk], s0);
  float v = (s0 + s1) + (s2 + s3);
  if (LANES == 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane != 0) return;
  }
  if (p.bias) v += p.bias[n];
  if (p.relu) v = fmaxf(v, 0.f);
  p.c[e] = v;
}

// out[e] = sum_s ws[s][e], one warp per output: lanes take every 32nd partial, fixed-order shuffle tree
// (hundreds of short wgrad splits folded into a few dozen filter taps)
__global__ void __launch_bounds__(256) splitk_reduce_warp_kernel(float* __restrict__ out, const float* __restrict__ ws,
                                                                 int S, long long n) {
  const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= n) return;
  const int lane = threadIdx.x & 31;
  float a0 = 0.f, a1 = 0.f;
  int s = lane;
  for (; s + 32 < S; s += 64) {
    a0 += ws[(long long)s * n + e];
    a1 += ws[(long long)(s + 32) * n + e];
  }
  if (s < S) a0 += ws[(long long)s * n + e];
  float v = a0 + a1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) out[e] = v;
}

static inline int pick_tile(long long I, int J, bool wide_ok) {
  if (J > 64) return TILE_A;
  if (J > 32) return (wide_ok && I >= 256) ? TILE_B : TILE_D;
  if (J > 16) return (wide_ok && I >= 512) ? TILE_C : TILE_E;
  return TILE_F;
}

#define NG_TILE_SWITCH(tile, EXPR)                      \
  switch (tile) {                                       \
    case TILE_A: { using TT = TileA; EXPR; } break;     \
    case TILE_B: { using TT = TileB; EXPR; } break;     \
    case TILE_C: { using TT = TileC; EXPR; } break;     \
    case TILE_D: { using TT = TileD; EXPR; } break;     \
    case TILE_E: { using TT = TileE; EXPR; } break;     \
    default:     { using TT = TileF; EXPR; } break;     \
  }
#define NG_TILE_SWITCH_NARROW(tile, EXPR)               \
  switch (tile) {                                       \
    case TILE_A: { using TT = TileA; EXPR; } break;     \
    case TILE_D: { using TT = TileD; EXPR; } break;     \
    case TILE_E: { using TT = TileE; EXPR; } break;     \
    default:     { using TT = TileF; EXPR; } break;     \
  }

// Number of reduction splits for `tiles` output tiles on `slots` concurrently resident CTAs:
// the candidate floor(w*slots/tiles), w = 1..8 waves, with the best wave occupancy (ties: fewer splits).
static inline long long pick_splits(long long tiles, long long slots, long long max_splits) {
  if (max_splits < 1) max_splits = 1;
  long long best = 1;
  double best_eff = 0.0;
  for (int w = 1; w <= 8; ++w) {
    long long s = (w * slots) / tiles;
    if (s < 1) s = 1;
    if (s > max_splits) s = max_splits;
    const long long ctas = tiles * s;
    const double eff = (double)ctas / (double)(ceil_div(ctas, slots) * slots);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    if (s == max_splits) break;
  }
  return best;
}

static int launch_conv_gather(const ConvGatherP& p, bool strided) {
  const int tile = pick_tile(p.M, p.J, true);
  const TileShape ts = tile_shape(tile);
  dim3 grid((unsigned)ceil_div(p.M, ts.bm), (unsigned)ceil_div(p.J, ts.bn), 1);
  if (strided) {
    NG_TILE_SWITCH(tile, NG_LAUNCH((conv_gather_kernel<TT, GATHER_STRIDED>), grid, CT_THREADS, 0, p));
  } else if (p.KC % CT_BK == 0) {
    NG_TILE_SWITCH(tile, NG_LAUNCH((conv_gather_kernel<TT, GATHER_TAPMAJOR>), grid, CT_THREADS, 0, p));
  } else {
    NG_TILE_SWITCH(tile, NG_LAUNCH((conv_gather_kernel<TT, GATHER_GENERIC>), grid, CT_THREADS, 0, p));
  }
  NG_CHECK_LAUNCH();
  return 0;
}

// Engine selection for one convolution layer: 0 = FP32 SIMT, 1 = the TF32-family tensor-core kernels
// (umma_conv.cu: TF32 or 3xTF32), 2 = the FP16x3 tensor-core kernels (umma_conv_f16.cu).  The staged
// channels-last copies of the two families differ, so all three passes of a layer use the same family.
static inline int conv_engine(int flags, double flops, int C, int K, int RS, int stride) {
  if (flags & NG_F_TF32) return 1;
  if (!(flags & (NG_F_TF32X3 | NG_F_F16X3))) return 0;
  if (!((flags & NG_F_TC_FORCE) || flops >= 2.0e8)) return 0;
#ifndef NG_EMU
  if ((flags & NG_F_F16X3) && stride == 1 && umma::f16x3_supported(C, K, RS)) return 2;
#else
  (void)C; (void)K; (void)RS; (void)stride;
#endif
  return (flags & NG_F_TF32X3) ? 1 : 0;
}

static int check_conv_dims(const char* who, int N, int C, int H, int W, int K, int R, int S, int stride, int pad_t,
                           int pad_l, int OH, int OW) {
  NG_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && K > 0 && R > 0 && S > 0, "%s: non-positive dimension", who);
  NG_REQUIRE(stride >= 1, "%s: stride must be >= 1 (got %d)", who, stride);
  NG_REQUIRE(R * S <= 64, "%s: filters with more than 64 taps are not supported (R*S = %d)", who, R * S);
  NG_REQUIRE(pad_t >= 0 && pad_l >= 0, "%s: negative padding", who);
  NG_REQUIRE(OH > 0 && OW > 0, "%s: empty output (%d x %d)", who, OH, OW);
  NG_REQUIRE((long long)N * K * OH * OW < 2147483647LL && (long long)N * C * H * W < 2147483647LL,
             "%s: tensor too large for 32-bit indexing", who);
  return 0;
}

// =====================================================================================
// Direct wgrad for narrow layers (C, K <= 32, S == 3, stride 1): conv_wgrad_smallc_kernel generalised to any output
// width (rows are zero-padded to a multiple of 4 in shared memory) and to a KT x RPG register tile: thread (kg, g)
// keeps dW[kg*KT + 0..KT-1][row][0..2] for its RPG filter rows (row = c*R + r); per 4-pixel strip KT LDS.128 of dy
// and 2*RPG LDS.128 of x feed KT*RPG*12 FFMA.  One partial per CTA, fixed-order reduce (splitk_reduce_wide_kernel).
// =====================================================================================
struct WgradDirectP {
  const float* x;
  const float* dy;
  float* ws;  // [gridDim.x][K][C*R*3]
  int N, C, H, W, K, R, OH, OW, OWP, pad_t, pad_l;
  int TR, row_blocks, tiles, NR, KG, G, PSN, xrows, xpitch, dpitch;
};

template <int KT, int RPG>
__global__ void __launch_bounds__(512) conv_wgrad_direct_kernel(const __grid_constant__ WgradDirectP p) {
  NG_DYN_SMEM(float, sm);
  float* dys = sm;                       // [K][dpitch]: TR rows of OWP pixels
  float* xs = sm + p.K * p.dpitch;       // [C][xrows][xpitch]: column j holds input column j - pad_l
  const int tid = threadIdx.x, nthr = blockDim.x;
  // thread = (kg, g, ps): KT output channels x RPG filter rows, and one of PSN interleaved subsets of the tile's
  // 4-pixel strips (with K/KT x NR/RPG = 128 threads alone a CTA keeps 4 warps busy: the strips are what is left to
  // split); the PSN partial tiles are added in a fixed order through shared memory at the end
  const int base = p.KG * p.G;
  const int ps = tid / base, tb = tid - ps * base;
  const int kg = tb % p.KG, g = tb / p.KG;
  int xoff[RPG];
  bool live[RPG];
  float acc[KT][RPG][3];
#pragma unroll
  for (int q = 0; q < RPG; ++q) {
    const int row = g * RPG + q;
    live[q] = row < p.NR;
    const int c = live[q] ? row / p.R : 0, r = live[q] ? row % p.R : 0;
    xoff[q] = (c * p.xrows + r) * p.xpitch;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) acc[kk][q][0] = acc[kk][q][1] = acc[kk][q][2] = 0.f;
  }
  const int dtile = p.K * p.TR * p.OWP, xtile = p.C * p.xrows * p.xpitch;
  const int spr = p.OWP >> 2, nstrips = p.TR * spr;   // strips per row / per tile
  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const int n = tile / p.row_blocks, oh0 = (tile - n * p.row_blocks) * p.TR;
    __syncthreads();  // the previous tile has been consumed
    for (int e = tid; e < dtile; e += nthr) {
      const int ow = e % p.OWP, t = e / p.OWP;
      const int tr = t % p.TR, kk = t / p.TR;
      float v = 0.f;
      if (ow < p.OW && oh0 + tr < p.OH) v = p.dy[((long long)(n * p.K + kk) * p.OH + oh0 + tr) * p.OW + ow];
      dys[kk * p.dpitch + tr * p.OWP + ow] = v;
    }
    for (int e = tid; e < xtile; e += nthr) {
      const int j = e % p.xpitch, t = e / p.xpitch;
      const int rr = t % p.xrows, c = t / p.xrows;
      const int ih = oh0 + rr - p.pad_t, iw = j - p.pad_l;
      float v = 0.f;
      if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) v = p.x[((long long)(n * p.C + c) * p.H + ih) * p.W + iw];
      xs[e] = v;
    }
    __syncthreads();
    const float* drow = dys + (kg * KT) * p.dpitch;
    for (int si = ps; si < nstrips; si += p.PSN) {
      const int tr = si / spr, ow0 = (si - tr * spr) << 2;
      float4 d[KT];
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) d[kk] = *reinterpret_cast<const float4*>(drow + kk * p.dpitch + tr * p.OWP + ow0);
#pragma unroll
      for (int q = 0; q < RPG; ++q) {
        if (!live[q]) continue;
        const float* xr = xs + xoff[q] + tr * p.xpitch + ow0;
        const float4 a = *reinterpret_cast<const float4*>(xr);
        const float4 b = *reinterpret_cast<const float4*>(xr + 4);
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
          acc[kk][q][0] = fmaf(d[kk].x, a.x, fmaf(d[kk].y, a.y, fmaf(d[kk].z, a.z, fmaf(d[kk].w, a.w, acc[kk][q][0]))));
          acc[kk][q][1] = fmaf(d[kk].x, a.y, fmaf(d[kk].y, a.z, fmaf(d[kk].z, a.w, fmaf(d[kk].w, b.x, acc[kk][q][1]))));
          acc[kk][q][2] = fmaf(d[kk].x, a.z, fmaf(d[kk].y, a.w, fmaf(d[kk].z, b.x, fmaf(d[kk].w, b.y, acc[kk][q][2]))));
        }
      }
    }
  }
  // fold the PSN strip subsets (fixed order), then one partial per CTA
  __syncthreads();
  constexpr int NA = KT * RPG * 3;
  if (ps > 0) {
    float* slot = sm + ((size_t)(ps - 1) * base + tb) * NA;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk)
#pragma unroll
      for (int q = 0; q < RPG; ++q)
#pragma unroll
        for (int t = 0; t < 3; ++t) slot[(kk * RPG + q) * 3 + t] = acc[kk][q][t];
  }
  __syncthreads();
  if (ps == 0) {
    for (int o = 1; o < p.PSN; ++o) {
      const float* slot = sm + ((size_t)(o - 1) * base + tb) * NA;
#pragma unroll
      for (int kk = 0; kk < KT; ++kk)
#pragma unroll
        for (int q = 0; q < RPG; ++q)
#pragma unroll
          for (int t = 0; t < 3; ++t) acc[kk][q][t] += slot[(kk * RPG + q) * 3 + t];
    }
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      float* out = p.ws + ((long long)blockIdx.x * p.K + kg * KT + kk) * (p.NR * 3);
#pragma unroll
      for (int q = 0; q < RPG; ++q) {
        if (!live[q]) continue;
        const int row = g * RPG + q;
        out[row * 3 + 0] = acc[kk][q][0];
        out[row * 3 + 1] = acc[kk][q][1];
        out[row * 3 + 2] = acc[kk][q][2];
      }
    }
  }
}

// Returns 0 when done, -1 when the shape is not one the direct kernel takes, > 0 on error.
static int launch_wgrad_direct(float* dw, const float* dy, const float* x, int N, int C, int H, int W, int K, int R, int S,
                               int stride, int pad_t, int pad_l, int OH, int OW) {
  // measured (config 1, batch 256): C = K = 16: 96 us against 130 us for the split-K implicit GEMM; C = K = 32: 266 us
  // against 222 us, so the implicit GEMM (or, in the FP32-accurate default, the 3xTF32 tensor-core kernel) keeps those
  if (S != 3 || stride != 1 || C > 32 || K > 32 || C < 8 || K % 2 != 0 || C * K > 512) return -1;
#ifdef NG_EMU
  const long long min_pixels = 1024;
#else
  const long long min_pixels = 65536;
#endif
  if ((long long)N * OH * OW < min_pixels) return -1;
  constexpr int KT = 2, RPG = 3;
  WgradDirectP q;
  memset(&q, 0, sizeof(q));
  q.x = x; q.dy = dy;
  q.N = N; q.C = C; q.H = H; q.W = W; q.K = K; q.R = R; q.OH = OH; q.OW = OW;
  q.OWP = (OW + 3) & ~3;
  q.pad_t = pad_t; q.pad_l = pad_l;
  q.NR = C * R;
  q.KG = K / KT;
  q.G = (q.NR + RPG - 1) / RPG;
  const int base = q.KG * q.G;
  if (base > 512 || base < 32) return -1;
  q.PSN = 512 / base;            // strip subsets: fill the CTA up to 512 threads
  if (q.PSN > 8) q.PSN = 8;
  const int threads = base * q.PSN;
  q.xpitch = q.OWP + 4;
  // the largest row block whose dy and x tiles fit 44 KB of shared memory
  int TR = OH;
  for (;; --TR) {
    if (TR < 1) return -1;
    const size_t bytes = sizeof(float) * ((size_t)K * (TR * q.OWP + 4) + (size_t)C * (TR + R - 1) * q.xpitch);
    if (bytes <= 44 * 1024) break;
  }
  q.TR = TR;
  q.row_blocks = (int)ceil_div(OH, TR);
  q.tiles = N * q.row_blocks;
  q.xrows = TR + R - 1;
  q.dpitch = TR * q.OWP + 4;
  size_t smem = sizeof(float) * ((size_t)K * q.dpitch + (size_t)C * q.xrows * q.xpitch);
  const size_t fold = sizeof(float) * (size_t)(q.PSN - 1) * base * KT * RPG * 3;   // reuses the tile memory at the end
  if (fold > smem) smem = fold;
  if (smem > 48 * 1024) return -1;
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  int ctas_per_sm = (int)((200u * 1024u) / (smem + 1024u));
  if (ctas_per_sm > 2048 / threads) ctas_per_sm = 2048 / threads;
  if (ctas_per_sm > 4) ctas_per_sm = 4;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  int grid = sms * ctas_per_sm;
  if (grid > q.tiles) grid = q.tiles;
  const long long out_elems = (long long)K * q.NR * 3;
  float* ws = nullptr;
  if (pool_alloc((void**)&ws, (uint64_t)((long long)grid * out_elems) * sizeof(float))) return 1;
  q.ws = ws;
  ProfScope prof_(NG_PROF_CONV_WGRAD, 2.0 * N * K * C * R * S * (double)OH * OW,
                  4.0 * ((double)N * C * H * W + (double)K * C * R * S + (double)N * K * OH * OW));
  NG_LAUNCH((conv_wgrad_direct_kernel<KT, RPG>), grid, threads, smem, q);
  NG_CHECK_LAUNCH();
  NG_LAUNCH(splitk_reduce_wide_kernel, (int)ceil_div(out_elems, 64), 512, 0, dw, (const float*)ws, grid, out_elems);
  NG_CHECK_LAUNCH();
  pool_free(ws);
  return 0;
}

// =====================================================================================
// Sliding-window direct wgrad (3 x 3, stride 1, narrow layers with K % 4 == 0): the kernel above spends 8 LDS.128 per
// 72 FFMA (the shared-memory return path, 128 B/clk/SM, saturates at half the FFMA rate) and stalls the whole CTA on
// every tile load.  Here thread (kg, c, cs) owns dW[4 kg .. 4 kg + 3][c][0..2][0..2] and walks one 4-pixel column
// strip down the tile's rows: output row tr needs input rows tr .. tr + 2, two of which are already in registers, so a
// row costs 4 LDS.128 of dy + one LDS.128 + LDS.64 of x for 144 FFMA.  Tiles are double-buffered with cp.async
// (zero-filled at the borders), dy is stored [row][strip][k][4] (+ 4 pad words per strip) so that a warp's reads are
// contiguous and the tile writes conflict-free, the x channel pitch is 4 (mod 32) words.  One partial per CTA, fixed-order reduce (splitk_reduce_wide_kernel).
// =====================================================================================
struct WgradSlideP {
  const float* x;
  const float* dy;
  float* ws;  // [gridDim.x][K][C*9]
  int N, C, H, W, K, OH, OW, OWP, pad_t, pad_l;
  int TR, row_blocks, tiles, KG, CS, CSN, SP, xrows, xpitch, cpitch, tile_floats;
  int d_lshift, x_lshift;   // log2 of the lanes that span a dy / x tile row in the copy
  int d_iters, x_iters;     // row iterations of a warp per channel
  FastDiv f_rb;
};

// cp.async of VEC floats with the ignore-src predicate: zeros are written when !ok (the source-size form costs ~20 extra
// instructions per copy).  dst is a shared-window address (a generic pointer under the host emulation).
#ifdef NG_EMU
typedef float* smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr(float* p) { return p; }
#define NG_SMEM_STEP(words) (words)
#else
typedef uint32_t smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr(float* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define NG_SMEM_STEP(words) (4 * (words))
#endif
template <int VEC>
__device__ __forceinline__ void cp_async_zfill(smem_addr_t dst, const float* src, bool ok) {
#ifdef NG_EMU
#pragma unroll
  for (int i = 0; i < VEC; ++i) dst[i] = ok ? src[i] : 0.f;
#else
  if (VEC == 4)
    asm volatile("{\n .reg .pred p;\n setp.eq.u32 p, %2, 0;\n cp.async.cg.shared.global [%0], [%1], 16, p;\n}\n" ::"r"(dst),
                 "l"(src), "r"((int)ok)
                 : "memory");
  else if (VEC == 2)
    asm volatile("{\n .reg .pred p;\n setp.eq.u32 p, %2, 0;\n cp.async.ca.shared.global [%0], [%1], 8, p;\n}\n" ::"r"(dst),
                 "l"(src), "r"((int)ok)
                 : "memory");
  else
    asm volatile("{\n .reg .pred p;\n setp.eq.u32 p, %2, 0;\n cp.async.ca.shared.global [%0], [%1], 4, p;\n}\n" ::"r"(dst),
                 "l"(src), "r"((int)ok)
                 : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef NG_EMU
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() {
#ifndef NG_EMU
  asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
#endif
}

// Tile copy: a warp owns whole channels and walks their rows with (a power-of-two group of) lanes across a row; the
// column tests are hoisted and source offset / destination address advance by a constant per row, so a copy costs about
// ten instructions (a flat element index costs two divisions per copy: a third of the kernel's issue slots).
template <int KT, int VD, int VX>
__device__ __forceinline__ void wgrad_slide_issue(const WgradSlideP& p, float* buf, int tile) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int n = (int)fastdiv((uint32_t)tile, p.f_rb);
  const int oh0 = (tile - n * p.row_blocks) * p.TR;
  // dy tile [TR][CS][SP]: per strip KG groups of KT float4 + 4 pad words
  {
    const float* img = p.dy + (long long)n * p.K * p.OH * p.OW;
    const int cols = p.OWP / VD, lpr = 1 << p.d_lshift, rpi = 32 >> p.d_lshift;
    const int lcol = lane & (lpr - 1), lsub = lane >> p.d_lshift;
    const int dplane = p.OH * p.OW, drows = p.OH - oh0, rstep = p.CS * p.SP;
    for (int col = lcol; col < cols; col += lpr) {
      const int ow = col * VD;
      const bool cok = ow < p.OW;
      for (int kk = warp; kk < p.K; kk += nwarps) {
        int off = kk * dplane + (oh0 + lsub) * p.OW + ow;
        smem_addr_t dst = smem_addr(buf + (kk / KT) * (KT * 4 + 4) + ((kk % KT) << 2) + lsub * rstep + (ow >> 2) * p.SP + (ow & 3));
        // (trip counts come from the host: a loop with a run-time stride makes the compiler divide for the count)
        for (int i = 0, tr = lsub; i < p.d_iters; ++i, tr += rpi, off += rpi * p.OW, dst += NG_SMEM_STEP(rpi * rstep)) {
          if (tr >= p.TR) break;      // (only the last iteration of the upper lane group)
          const bool ok = cok && tr < drows;
          cp_async_zfill<VD>(dst, img + (ok ? off : 0), ok);
        }
      }
    }
  }
  // x tile [C][cpitch], rows of xpitch: column j holds input column j - pad_l
  {
    const float* img = p.x + (long long)n * p.C * p.H * p.W;
    float* xs = buf + p.TR * p.CS * p.SP;
    const int cols = p.xpitch / VX, lpr = 1 << p.x_lshift, rpi = 32 >> p.x_lshift;
    const int lcol = lane & (lpr - 1), lsub = lane >> p.x_lshift;
    const int xplane = p.H * p.W;
    const int rr_lo = p.pad_t - oh0, rr_hi = p.H + p.pad_t - oh0;   // valid rr in [rr_lo, rr_hi)
    for (int col = lcol; col < cols; col += lpr) {
      const int j = col * VX;
      const bool cok = j >= p.pad_l && j < p.W + p.pad_l;
      for (int c = warp; c < p.C; c += nwarps) {
        int off = c * xplane + (oh0 - p.pad_t + lsub) * p.W + j - p.pad_l;
        smem_addr_t dst = smem_addr(xs + c * p.cpitch + lsub * p.xpitch + j);
        for (int i = 0, rr = lsub; i < p.x_iters; ++i, rr += rpi, off += rpi * p.W, dst += NG_SMEM_STEP(rpi * p.xpitch)) {
          if (rr >= p.xrows) break;   // (only the last iteration of the upper lane group)
          const bool ok = cok && rr >= rr_lo && rr < rr_hi;
          cp_async_zfill<VX>(dst, img + (ok ? off : 0), ok);
        }
      }
    }
  }
  cp_async_commit();
}

template <int KT, int VD, int VX>
__global__ void __launch_bounds__(512, 1) conv_wgrad_slide_kernel(const __grid_constant__ WgradSlideP p) {
  NG_DYN_SMEM(float, sm);
  const int tid = threadIdx.x;
  const int base = p.KG * p.C;
  const int ps = tid / base, tb = tid - ps * base;
  const int kg = tb % p.KG, c = tb / p.KG;
  float acc[KT][3][3];
#pragma unroll
  for (int kk = 0; kk < KT; ++kk)
#pragma unroll
    for (int r = 0; r < 3; ++r) acc[kk][r][0] = acc[kk][r][1] = acc[kk][r][2] = 0.f;
  const int dstep = p.CS * p.SP;
  int it = 0;
  if ((int)blockIdx.x < p.tiles) wgrad_slide_issue<KT, VD, VX>(p, sm, blockIdx.x);
  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
    float* buf = sm + (it & 1) * p.tile_floats;
    const int next = tile + gridDim.x;
    if (next < p.tiles) {
      wgrad_slide_issue<KT, VD, VX>(p, sm + ((it + 1) & 1) * p.tile_floats, next);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int n = (int)fastdiv((uint32_t)tile, p.f_rb);
    const int oh0 = (tile - n * p.row_blocks) * p.TR;
    const int rows = p.OH - oh0 < p.TR ? p.OH - oh0 : p.TR;
    const float* dys = buf;
    const float* xs = buf + p.TR * p.CS * p.SP + c * p.cpitch;
    // (threads past base * CSN only round the block up to whole warps: they copy tiles and idle here)
    for (int cs = ps < p.CSN ? ps : p.CS; cs < p.CS; cs += p.CSN) {
      const float* xr = xs + cs * 4;
      const float* dr = dys + cs * p.SP + kg * (KT * 4 + 4);
      // rows tr .. tr + 2 of the window and dy row tr are in registers when a step starts; the step first issues the
      // loads of the next one (window row tr + 3 into the fourth slot, dy row tr + 1 into the other d buffer: past the
      // tile's last row these read the slack behind it and are never used), then runs its 144 FFMA
      float4 xa[4];
      float2 xb[4];
      float4 d[2][KT];
#define NG_SLIDE_LOADX(slot, row)                                               \
  do {                                                                          \
    xa[slot] = *reinterpret_cast<const float4*>(xr + (row) * p.xpitch);         \
    xb[slot] = *reinterpret_cast<const float2*>(xr + (row) * p.xpitch + 4);     \
  } while (0)
#define NG_SLIDE_LOADD(db, row)                                                                                        \
  do {                                                                                                                  \
    _Pragma("unroll") for (int kk = 0; kk < KT; ++kk) d[db][kk] = *reinterpret_cast<const float4*>(dr + (row) * dstep + kk * 4); \
  } while (0)
#define NG_SLIDE_FMA(a3, d, A, B)                                                                         \
  do {                                                                                                    \
    a3[0] = fmaf(d.x, A.x, fmaf(d.y, A.y, fmaf(d.z, A.z, fmaf(d.w, A.w, a3[0]))));                        \
    a3[1] = fmaf(d.x, A.y, fmaf(d.y, A.z, fmaf(d.z, A.w, fmaf(d.w, B.x, a3[1]))));                        \
    a3[2] = fmaf(d.x, A.z, fmaf(d.y, A.w, fmaf(d.z, B.x, fmaf(d.w, B.y, a3[2]))));                        \
  } while (0)
#define NG_SLIDE_STEP(i0, i1, i2, i3, db)                                                                 \
  do {                                                                                                    \
    NG_SLIDE_LOADX(i3, tr + 3);                                                                           \
    NG_SLIDE_LOADD((db) ^ 1, tr + 1);                                                                     \
    _Pragma("unroll") for (int kk = 0; kk < KT; ++kk) {                                                   \
      NG_SLIDE_FMA(acc[kk][0], d[db][kk], xa[i0], xb[i0]);                                                \
      NG_SLIDE_FMA(acc[kk][1], d[db][kk], xa[i1], xb[i1]);                                                \
      NG_SLIDE_FMA(acc[kk][2], d[db][kk], xa[i2], xb[i2]);                                                \
    }                                                                                                     \
  } while (0)
      NG_SLIDE_LOADX(0, 0);
      NG_SLIDE_LOADX(1, 1);
      NG_SLIDE_LOADX(2, 2);
      NG_SLIDE_LOADD(0, 0);
      int tr = 0;
      while (true) {
        NG_SLIDE_STEP(0, 1, 2, 3, 0);
        if (++tr >= rows) break;
        NG_SLIDE_STEP(1, 2, 3, 0, 1);
        if (++tr >= rows) break;
        NG_SLIDE_STEP(2, 3, 0, 1, 0);
        if (++tr >= rows) break;
        NG_SLIDE_STEP(3, 0, 1, 2, 1);
        if (++tr >= rows) break;
      }
#undef NG_SLIDE_STEP
#undef NG_SLIDE_FMA
#undef NG_SLIDE_LOADD
#undef NG_SLIDE_LOADX
    }
    __syncthreads();   // the buffer is refilled by the next iteration's prefetch
  }
  // fold the CSN column-strip subsets (fixed order), then one partial per CTA
  constexpr int NA = KT * 9;
  if (ps > 0 && ps < p.CSN) {
    float* slot = sm + ((size_t)(ps - 1) * base + tb) * NA;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int t = 0; t < 3; ++t) slot[kk * 9 + r * 3 + t] = acc[kk][r][t];
  }
  __syncthreads();
  if (ps == 0) {
    for (int o = 1; o < p.CSN; ++o) {
      const float* slot = sm + ((size_t)(o - 1) * base + tb) * NA;
#pragma unroll
      for (int kk = 0; kk < KT; ++kk)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int t = 0; t < 3; ++t) acc[kk][r][t] += slot[kk * 9 + r * 3 + t];
    }
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      float* out = p.ws + ((long long)blockIdx.x * p.K + kg * KT + kk) * (p.C * 9) + c * 9;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int t = 0; t < 3; ++t) out[r * 3 + t] = acc[kk][r][t];
    }
  }
}

// Returns 0 when done, -1 when the shape is not one this kernel takes, > 0 on error.
static int launch_wgrad_slide(float* dw, const float* dy, const float* x, int N, int C, int H, int W, int K, int R, int S,
                              int stride, int pad_t, int pad_l, int OH, int OW) {
  constexpr int KT = 4;
  if (R != 3 || S != 3 || stride != 1 || C > 32 || K > 32 || C < 8 || K % KT != 0 || C * K > 1024) return -1;
#ifdef NG_EMU
  const long long min_pixels = 1024;
#else
  const long long min_pixels = 65536;
#endif
  if ((long long)N * OH * OW < min_pixels) return -1;
  WgradSlideP q;
  memset(&q, 0, sizeof(q));
  q.x = x; q.dy = dy;
  q.N = N; q.C = C; q.H = H; q.W = W; q.K = K; q.OH = OH; q.OW = OW;
  q.OWP = (OW + 3) & ~3;
  q.pad_t = pad_t; q.pad_l = pad_l;
  q.KG = K / KT;
  q.CS = q.OWP / 4;
  // a strip holds KG groups of KT float4 + 4 pad words (group pitch 20 words: the four groups a warp reads together sit in
  // distinct banks); strip pitch = 4 x odd words, so the 8-byte tile writes of 8 neighbouring strips (one dy row
  // segment) land in 32 distinct banks
  q.SP = q.KG * (KT * 4 + 4);
  if ((q.SP / 4) % 2 == 0) q.SP += 4;
  const int base = q.KG * C;
  if (base > 512 || base < 32) return -1;
  // CTA size: 512 threads, one CTA per SM, or (NG_WGRAD_SLIDE_THREADS=256) two 256-thread CTAs whose barriers and tile
  // copies interleave
  static const int tmax_env = getenv("NG_WGRAD_SLIDE_THREADS") ? atoi(getenv("NG_WGRAD_SLIDE_THREADS")) : 512;
  const int tmax = tmax_env == 256 ? 256 : 512;
  q.CSN = tmax / base;
  if (q.CSN < 1) q.CSN = 1;
  if (q.CSN > q.CS) q.CSN = q.CS;
  const int threads = (base * q.CSN + 31) & ~31;
  const int ctas_per_sm = threads <= 256 ? 2 : 1;
  const size_t tile_budget = ctas_per_sm == 2 ? 108u * 1024u : 200u * 1024u;
  q.xpitch = q.OWP + 4;
  const int sms = (g_sm_count > 0 ? g_sm_count : 148) * ctas_per_sm;
  // rows per tile: two tiles must fit 200 KB; among the row-block counts that do, the one with the fewest
  // (rows + window start-up) x rounds of the persistent grid
  int best_rb = 0, best_tr = 0, best_cp = 0;
  double best_cost = 0.0;
  static const int rb_forced = getenv("NG_WGRAD_SLIDE_RB") ? atoi(getenv("NG_WGRAD_SLIDE_RB")) : 0;   // tuning runs
  for (int rb = 1; rb <= OH && rb <= 64; ++rb) {
    if (rb_forced > 0 && rb != rb_forced) continue;
    const int TR = (int)ceil_div(OH, rb);
    if ((int)ceil_div(OH, TR) != rb) continue;
    int cp = (TR + 2) * q.xpitch;
    cp += ((4 - cp % 32) + 32) % 32;   // channel pitch == 4 (mod 32) words: 8 channels' float4 reads hit 32 distinct banks
    const size_t tile = (size_t)TR * q.CS * q.SP + (size_t)C * cp;
    if (2 * tile * sizeof(float) > tile_budget) continue;
    const long long tiles = (long long)N * rb;
    const long long grid = tiles < sms ? tiles : sms;
    const double cost = (double)ceil_div(tiles, grid) * (TR + 1.0);
    if (best_rb == 0 || cost < best_cost) { best_rb = rb; best_tr = TR; best_cp = cp; best_cost = cost; }
  }
  if (best_rb == 0) return -1;
  q.TR = best_tr;
  q.row_blocks = best_rb;
  q.tiles = N * best_rb;
  q.xrows = best_tr + 2;
  q.cpitch = best_cp;
  q.tile_floats = q.TR * q.CS * q.SP + C * q.cpitch;
  // + slack for the window row prefetched past the last channel's rows of the second buffer
  size_t smem = sizeof(float) * (2 * (size_t)q.tile_floats + q.xpitch + 8);
  const size_t fold = sizeof(float) * (size_t)(q.CSN - 1) * base * KT * 9;
  if (fold > smem) smem = fold;
  if (smem > 220u * 1024u) return -1;
  // widest copies for which every row start and every zero-fill boundary is aligned: dy rows 8 bytes (even OW), x rows
  // 16 bytes (W % 4 == 0, no left padding) or 8
  const int vd = (OW % 2 == 0 && ((uintptr_t)dy & 7) == 0) ? 2 : 1;
  const int vx = (W % 4 == 0 && pad_l % 4 == 0 && ((uintptr_t)x & 15) == 0) ? 4
                 : (W % 2 == 0 && pad_l % 2 == 0 && ((uintptr_t)x & 7) == 0) ? 2 : 1;
  for (q.d_lshift = 0; q.d_lshift < 5 && (1 << q.d_lshift) < q.OWP / vd; ++q.d_lshift) {}
  for (q.x_lshift = 0; q.x_lshift < 5 && (1 << q.x_lshift) < q.xpitch / vx; ++q.x_lshift) {}
  q.f_rb = make_fastdiv(q.row_blocks);
  q.d_iters = (int)ceil_div(q.TR, 32 >> q.d_lshift);
  q.x_iters = (int)ceil_div(q.xrows, 32 >> q.x_lshift);
  const int grid = q.tiles < sms ? q.tiles : sms;
  const long long out_elems = (long long)K * C * 9;
  float* ws = nullptr;
  if (pool_alloc((void**)&ws, (uint64_t)((long long)grid * out_elems) * sizeof(float))) return 1;
  q.ws = ws;
#ifndef NG_EMU
  {
    static bool attr_done = false;
    if (!attr_done) {
      const int big = 220 * 1024;
      cudaError_t e = cudaFuncSetAttribute(conv_wgrad_slide_kernel<KT, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_wgrad_slide_kernel<KT, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_wgrad_slide_kernel<KT, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_wgrad_slide_kernel<KT, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
      if (e != cudaSuccess) {
        pool_free(ws);
        return set_error("conv_wgrad_slide_kernel: %s", cudaGetErrorString(e));
      }
      attr_done = true;
    }
  }
#endif
  ProfScope prof_(NG_PROF_CONV_WGRAD, 2.0 * N * K * C * R * S * (double)OH * OW,
                  4.0 * ((double)N * C * H * W + (double)K * C * R * S + (double)N * K * OH * OW));
  if (vd == 2 && vx == 4) NG_LAUNCH((conv_wgrad_slide_kernel<KT, 2, 4>), grid, threads, smem, q);
  else if (vd == 2 && vx == 2) NG_LAUNCH((conv_wgrad_slide_kernel<KT, 2, 2>), grid, threads, smem, q);
  else if (vd == 2) NG_LAUNCH((conv_wgrad_slide_kernel<KT, 2, 1>), grid, threads, smem, q);
  else NG_LAUNCH((conv_wgrad_slide_kernel<KT, 1, 1>), grid, threads, smem, q);
  NG_CHECK_LAUNCH();
  NG_LAUNCH(splitk_reduce_wide_kernel, (int)ceil_div(out_elems, 64), 512, 0, dw, (const float*)ws, grid, out_elems);
  NG_CHECK_LAUNCH();
  pool_free(ws);
  return 0;
}

}  // namespace ng

using namespace ng;

extern "C" {

int ng_conv2d_fprop(uint64_t y, uint64_t x, uint64_t w, uint64_t bias, int N, int C, int H, int W, int K, int R,
                    int S, int stride, int pad_t, int pad_l, int OH, int OW, int flags) {
  NG_ENSURE_INIT();
  if (check_conv_dims("ng_conv2d_fprop", N, C, H, W, K, R, S, stride, pad_t, pad_l, OH, OW)) return 1;
  ConvGatherP p;
  memset(&p, 0, sizeof(p));
  p.a = dptr<const float>(x);
  p.b = dptr<const float>(w);
  p.bias = bias ? dptr<const float>(bias) : nullptr;
  p.out = dptr<float>(y);
  p.M = N * OH * OW;
  p.J = K;
  p.KC = C;
  p.RS = R * S;
  p.opix = OH * OW;
  p.ow_out = OW;
  p.a_chan_stride = H * W;
  p.a_img_stride = (long long)C * H * W;
  p.a_h = H;
  p.a_w = W;
  p.y_mul = stride; p.y_add = -pad_t;
  p.x_mul = stride; p.x_add = -pad_l;
  p.dstride = 1;
  p.b_sj = (long long)C * R * S;
  p.b_sc = (long long)R * S;
  p.vec_store = (p.opix % 4 == 0) && ((y & 15u) == 0);
  p.relu = (flags & NG_F_RELU) ? 1 : 0;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) { p.tap_dy[r * S + s] = (short)r; p.tap_dx[r * S + s] = (short)s; }
  const double flops = 2.0 * N * K * C * R * S * (double)OH * OW;
  const double bytes = 4.0 * ((double)N * C * H * W + (double)K * C * R * S + (double)N * K * OH * OW);
  const int tok = prof_begin(NG_PROF_CONV_FPROP, flops, bytes);
  int rc = -1;
#ifndef NG_EMU
  const int engine = conv_engine(flags, flops, C, K, R * S, stride);
  if (engine == 2)
    rc = umma::conv_fprop_f16(p.out, p.a, p.b, p.bias, N, C, H, W, K, R, S, stride, pad_t, pad_l, OH, OW, p.relu,
                              (flags & NG_F_X_NHWC) ? 1 : 0, (flags & NG_F_W_STAGED) ? 1 : 0);
  else if (engine == 1)
    rc = umma::conv_fprop_tf32(p.out, p.a, p.b, p.bias, N, C, H, W, K, R, S, stride, pad_t, pad_l, OH, OW, p.relu,
                               (flags & NG_F_TF32) ? 0 : 1, (flags & NG_F_X_NHWC) ? 1 : 0);
#endif
  if (rc < 0 && (flags & (NG_F_X_NHWC | NG_F_W_STAGED)))
    rc = set_error("ng_conv2d_fprop: staged operands given but this shape does not take the tensor-core kernel");
  if (rc < 0 && C <= 4 && R == 3 && S == 3 && stride == 1 && OW % 4 == 0 && (y & 15u) == 0 &&
      (long long)N * OH * OW >= 4096)
    rc = launch_fprop_smallc(p, N, C, H, W, K, OH, OW, pad_t, pad_l);
  if (rc < 0 && stride == 1) rc = launch_conv_direct(p, R, S, false, N, OH, OW);
  if (rc < 0) rc = launch_conv_gather(p, false);
  prof_end(tok);
  return rc;
}

int ng_conv2d_dgrad(uint64_t dx, uint64_t dy, uint64_t w, int N, int C, int H, int W, int K, int R, int S,
                    int stride, int pad_t, int pad_l, int OH, int OW, int flags) {
  NG_ENSURE_INIT();
  if (check_conv_dims("ng_conv2d_dgrad", N, C, H, W, K, R, S, stride, pad_t, pad_l, OH, OW)) return 1;
  // dx[n,c,ih,iw] = sum_{k,r,s} dy[n,k,(ih+pt-r)/st,(iw+pl-s)/st] * w[k,c,r,s]
  ConvGatherP p;
  memset(&p, 0, sizeof(p));
  p.a = dptr<const float>(dy);
  p.b = dptr<const float>(w);
  p.bias = nullptr;
  p.out = dptr<float>(dx);
  p.M = N * H * W;
  p.J = C;
  p.KC = K;
  p.RS = R * S;
  p.opix = H * W;
  p.ow_out = W;
  p.a_chan_stride = OH * OW;
  p.a_img_stride = (long long)K * OH * OW;
  p.a_h = OH;
  p.a_w = OW;
  p.y_mul = 1; p.y_add = pad_t;
  p.x_mul = 1; p.x_add = pad_l;
  p.dstride = stride;
  p.b_sj = (long long)R * S;      // j = c
  p.b_sc = (long long)C * R * S;  // kc = k
  p.vec_store = (p.opix % 4 == 0) && ((dx & 15u) == 0);
  p.relu = 0;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) { p.tap_dy[r * S + s] = (short)(-r); p.tap_dx[r * S + s] = (short)(-s); }
  const double flops = 2.0 * N * K * C * R * S * (double)OH * OW;
  const double bytes = 4.0 * ((double)N * C * H * W + (double)K * C * R * S + (double)N * K * OH * OW);
  const int tok = prof_begin(NG_PROF_CONV_DGRAD, flops, bytes);
  int rc = -1;
#ifndef NG_EMU
  const int engine = conv_engine(flags, flops, C, K, R * S, stride);
  if (engine == 2)
    rc = umma::conv_dgrad_f16(p.out, p.a, p.b, N, C, H, W, K, R, S, stride, pad_t, pad_l, OH, OW,
                              (flags & NG_F_DY_NHWC) ? 1 : 0, (flags & NG_F_W_STAGED) ? 1 : 0);
  else if (engine == 1)
    rc = umma::conv_dgrad_tf32(p.out, p.a, p.b, N, C, H, W, K, R, S, stride, pad_t, pad_l, OH, OW,
                               (flags & NG_F_TF32) ? 0 : 1, (flags & NG_F_DY_NHWC) ? 1 : 0);
#endif
  if (rc < 0 && (flags & (NG_F_DY_NHWC | NG_F_W_STAGED)))
    rc = set_error("ng_conv2d_dgrad: staged operands given but this shape does not take the tensor-core kernel");
  if (rc < 0 && stride == 1) rc = launch_conv_direct(p, R, S, true, N, H, W);
  if (rc < 0) rc = launch_conv_gather(p, stride > 1);
  prof_end(tok);
  return rc;
}

int ng_conv2d_tc_plan(int N, int C, int H, int W, int K, int R, int S, int stride, int OH, int OW, int flags,
                      int* out_fprop_dgrad_wgrad_engine) {
  (void)H; (void)W;
  const double flops = 2.0 * N * K * C * R * S * (double)OH * OW;
  int f = 0, d = 0, w = 0, engine = 0;
#ifndef NG_EMU
  if (stride == 1) engine = conv_engine(flags, flops, C, K, R * S, stride);
  if (engine == 2) {
    f = d = w = 1;
  } else if (engine == 1) {
    f = umma::gather_supported(C, K, R * S) ? 1 : 0;
    d = umma::gather_supported(K, C, R * S) ? 1 : 0;
    w = umma::wgrad_supported(C, K, R * S) ? 1 : 0;
  }
#else
  (void)flops;
#endif
  out_fprop_dgrad_wgrad_engine[0] = f;
  out_fprop_dgrad_wgrad_engine[1] = d;
  out_fprop_dgrad_wgrad_engine[2] = w;
  out_fprop_dgrad_wgrad_engine[3] = engine;
  return 0;
}

int ng_conv2d_stage_nhwc(uint64_t dst, uint64_t src, int64_t N, int64_t C, int64_t HW, int flags) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_STAGE, 0.0, 8.0 * (double)N * C * HW);
#ifndef NG_EMU
  NG_REQUIRE(N > 0 && C > 0 && HW > 0 && N * C * HW < 2147483647LL, "ng_conv2d_stage_nhwc: bad dimensions");
  if (flags & NG_F_STAGE_F16)
    return umma::to_nhwc_f16(dptr<void>(dst), dptr<const float>(src), nullptr, (int)N, (int)C, (int)HW, nullptr);
  return umma::to_nhwc(dptr<float>(dst), dptr<const float>(src), (int)N, (int)C, (int)HW, (flags & NG_F_TF32) ? 1 : 0);
#else
  (void)dst; (void)src; (void)N; (void)C; (void)HW; (void)flags;
  return set_error("ng_conv2d_stage_nhwc: the tensor-core path does not exist in the emulation build");
#endif
}

int ng_conv2d_stage_nhwc_sum(uint64_t dst, uint64_t src, uint64_t sums, int64_t N, int64_t C, int64_t HW, int flags) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_STAGE, 0.0, 8.0 * (double)N * C * HW);
#ifndef NG_EMU
  NG_REQUIRE(N > 0 && C > 0 && HW > 0 && N * C * HW < 2147483647LL && N <= 65535, "ng_conv2d_stage_nhwc_sum: bad dimensions");
  if (flags & NG_F_STAGE_F16)
    return umma::to_nhwc_f16(dptr<void>(dst), dptr<const float>(src), dptr<float>(sums), (int)N, (int)C, (int)HW, nullptr);
  return umma::to_nhwc_sum(dptr<float>(dst), dptr<const float>(src), dptr<float>(sums), (int)N, (int)C, (int)HW,
                           (flags & NG_F_TF32) ? 1 : 0);
#else
  (void)dst; (void)src; (void)sums; (void)N; (void)C; (void)HW; (void)flags;
  return set_error("ng_conv2d_stage_nhwc_sum: the tensor-core path does not exist in the emulation build");
#endif
}

int ng_conv2d_stage_filters_f16(uint64_t dst_fprop, uint64_t dst_dgrad, uint64_t w, int K, int C, int RS) {
  NG_ENSURE_INIT();
#ifndef NG_EMU
  NG_REQUIRE(K > 0 && C > 0 && RS > 0 && dst_fprop != 0, "ng_conv2d_stage_filters_f16: bad arguments");
  return umma::stage_filters_f16(dptr<void>(dst_fprop), dst_dgrad ? dptr<void>(dst_dgrad) : nullptr, dptr<const float>(w), K, C, RS);
#else
  (void)dst_fprop; (void)dst_dgrad; (void)w; (void)K; (void)C; (void)RS;
  return set_error("ng_conv2d_stage_filters_f16: the tensor-core path does not exist in the emulation build");
#endif
}

int ng_conv2d_stage_f16(uint64_t dst, uint64_t src, uint64_t sums, uint64_t absmax_bits, int64_t N, int64_t C, int64_t HW) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_STAGE, 0.0, 8.0 * (double)N * C * HW);
#ifndef NG_EMU
  NG_REQUIRE(N > 0 && C > 0 && HW > 0 && N * C * HW < 2147483647LL && (sums == 0 || N <= 65535),
             "ng_conv2d_stage_f16: bad dimensions");
  return umma::to_nhwc_f16(dptr<void>(dst), dptr<const float>(src), sums ? dptr<float>(sums) : nullptr, (int)N, (int)C,
                           (int)HW, absmax_bits ? dptr<const unsigned int>(absmax_bits) : nullptr);
#else
  (void)dst; (void)src; (void)sums; (void)absmax_bits; (void)N; (void)C; (void)HW;
  return set_error("ng_conv2d_stage_f16: the tensor-core path does not exist in the emulation build");
#endif
}

int ng_conv2d_stage_f16_bn(uint64_t dst, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t mean, uint64_t invstd,
                           uint64_t absmax_bits, int64_t N, int64_t C, int64_t HW, int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_STAGE, 0.0, 8.0 * (double)N * C * HW);
#ifndef NG_EMU
  NG_REQUIRE(N > 0 && C > 0 && HW > 0 && N * C * HW < 2147483647LL, "ng_conv2d_stage_f16_bn: bad dimensions");
  NG_REQUIRE(absmax_bits != 0 && (act == 0 || act == 1), "ng_conv2d_stage_f16_bn: absmax_bits required, act 0 or 1");
  return umma::to_nhwc_f16_bn(dptr<void>(dst), dptr<const float>(x), dptr<const float>(gamma), dptr<const float>(beta),
                              dptr<const float>(mean), dptr<const float>(invstd), act,
                              dptr<const unsigned int>(absmax_bits), (int)N, (int)C, (int)HW);
#else
  (void)dst; (void)x; (void)gamma; (void)beta; (void)mean; (void)invstd; (void)absmax_bits; (void)N; (void)C; (void)HW; (void)act;
  return set_error("ng_conv2d_stage_f16_bn: the tensor-core path does not exist in the emulation build");
#endif
}

int ng_conv2d_stage_f16_bnbwd(uint64_t dst, uint64_t chan_sums, uint64_t dy, uint64_t x, uint64_t gamma, uint64_t beta,
                              uint64_t mean, uint64_t invstd, uint64_t bn_sums, uint64_t bound_bits, int64_t N, int64_t C,
                              int64_t HW, int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_STAGE, 0.0, 12.0 * (double)N * C * HW);
#ifndef NG_EMU
  NG_REQUIRE(N > 0 && C > 0 && HW > 0 && N * C * HW < 2147483647LL && (chan_sums == 0 || N <= 65535),
             "ng_conv2d_stage_f16_bnbwd: bad dimensions");
  NG_REQUIRE(bound_bits != 0 && bn_sums != 0 && (act == 0 || (act == 1 && beta != 0)),
             "ng_conv2d_stage_f16_bnbwd: bound_bits and bn_sums required; act 0, or 1 with beta");
  return umma::to_nhwc_f16_bnbwd(dptr<void>(dst), chan_sums ? dptr<float>(chan_sums) : nullptr, dptr<const float>(dy),
                                 dptr<const float>(x), dptr<const float>(gamma), beta ? dptr<const float>(beta) : nullptr,
                                 dptr<const float>(mean), dptr<const float>(invstd), dptr<const float>(bn_sums), act,
                                 dptr<const unsigned int>(bound_bits), (int)N, (int)C, (int)HW);
#else
  (void)dst; (void)chan_sums; (void)dy; (void)x; (void)gamma; (void)beta; (void)mean; (void)invstd; (void)bn_sums;
  (void)bound_bits; (void)N; (void)C; (void)HW; (void)act;
  return set_error("ng_conv2d_stage_f16_bnbwd: the tensor-core path does not exist in the emulation build");
#endif
}

int ng_conv2d_stage_f16_bnbwd_pool(uint64_t dst, uint64_t chan_sums, uint64_t dpooled, uint64_t pooled, uint64_t x,
                                   uint64_t gamma, uint64_t beta, uint64_t mean, uint64_t invstd, uint64_t bn_sums,
                                   uint64_t bound_bits, int64_t N, int64_t C, int H, int W, int act) {
  NG_ENSURE_INIT();
  ProfScope prof_(NG_PROF_STAGE, 0.0, 10.0 * (double)N * C * H * W);   // x in, pooled + gradient (a quarter each), staged out
#ifndef NG_EMU
  NG_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && N * C * (int64_t)H * W < 2147483647LL && (chan_sums == 0 || N <= 65535),
             "ng_conv2d_stage_f16_bnbwd_pool: bad dimensions");
  NG_REQUIRE(bound_bits != 0 && bn_sums != 0 && beta != 0 && (act == 0 || act == 1),
             "ng_conv2d_stage_f16_bnbwd_pool: bound_bits, bn_sums and beta required; act 0 or 1");
  return umma::to_nhwc_f16_bnbwd_pool(dptr<void>(dst), chan_sums ? dptr<float>(chan_sums) : nullptr,
                                      dptr<const float>(dpooled), dptr<const float>(pooled), dptr<const float>(x),
                                      dptr<const float>(gamma), dptr<const float>(beta), dptr<const float>(mean),
                                      dptr<const float>(invstd), dptr<const float>(bn_sums), act,
                                      dptr<const unsigned int>(bound_bits), (int)N, (int)C, H, W);
#else
  (void)dst; (void)chan_sums; (void)dpooled; (void)pooled; (void)x; (void)gamma; (void)beta; (void)mean; (void)invstd;
  (void)bn_sums; (void)bound_bits; (void)N; (void)C; (void)H; (void)W; (void)act;
  return set_error("ng_conv2d_stage_f16_bnbwd_pool: the tensor-core path does not exist in the emulation build");
#endif
}

int ng_conv2d_wgrad(uint64_t dw, uint64_t dy, uint64_t x, int N, int C, int H, int W, int K, int R, int S,
                    int stride, int pad_t, int pad_l, int OH, int OW, int flags) {
  NG_ENSURE_INIT();
  if (check_conv_dims("ng_conv2d_wgrad", N, C, H, W, K, R, S, stride, pad_t, pad_l, OH, OW)) return 1;
#ifndef NG_EMU
  const double wflops = 2.0 * N * K * C * R * S * (double)OH * OW;
  const int engine = conv_engine(flags, wflops, C, K, R * S, stride);
  if (engine) {
    const int tok = prof_begin(NG_PROF_CONV_WGRAD, wflops,
                               4.0 * ((double)N * C * H * W + (double)K * C * R * S + (double)N * K * OH * OW));
    const int rc = engine == 2
        ? umma::conv_wgrad_f16(dptr<float>(dw), dptr<const void>(dy), dptr<const void>(x), N, C, H, W, K, R, S, stride,
                               pad_t, pad_l, OH, OW, (flags & NG_F_X_NHWC) ? 1 : 0, (flags & NG_F_DY_NHWC) ? 1 : 0)
        : umma::conv_wgrad_tf32(dptr<float>(dw), dptr<const float>(dy), dptr<const float>(x), N, C, H, W, K, R,
                                S, stride, pad_t, pad_l, OH, OW, (flags & NG_F_TF32) ? 0 : 1,
                                (flags & NG_F_X_NHWC) ? 1 : 0, (flags & NG_F_DY_NHWC) ? 1 : 0);
    prof_end(tok);
    if (rc >= 0) return rc;
  }
#endif
  NG_REQUIRE(!(flags & (NG_F_X_NHWC | NG_F_DY_NHWC)),
             "ng_conv2d_wgrad: staged NHWC operands given but this shape does not take the tensor-core kernel");
  {
    // narrow layers (both channel counts <= 32): direct kernels (see conv_wgrad_slide_kernel, conv_wgrad_direct_kernel)
    int rc = launch_wgrad_slide(dptr<float>(dw), dptr<const float>(dy), dptr<const float>(x), N, C, H, W, K, R, S, stride,
                                pad_t, pad_l, OH, OW);
    if (rc >= 0) return rc;
    rc = launch_wgrad_direct(dptr<float>(dw), dptr<const float>(dy), dptr<const float>(x), N, C, H, W, K, R, S,
                                       stride, pad_t, pad_l, OH, OW);
    if (rc >= 0) return rc;
  }
  const int sms_all = g_sm_count > 0 ? g_sm_count : 148;
  {
    // tiny layer: direct kernel over whole images in shared memory (see conv_wgrad_tiny_kernel)
    const long long O = (long long)K * C * R * S;
    const long long img_floats = (long long)C * H * W + (long long)K * OH * OW;
    const int G = O < 256 ? (int)(256 / O) : 1;
    const size_t smem = sizeof(float) * (size_t)(img_floats + (G > 1 ? G * O : 0));
    // (layers with more taps than threads run faster as many short implicit-GEMM splits: measured 17 us
    //  against 30 us for the README CNN's 8 -> 16 channel layer)
    if (2.0 * N * K * C * R * S * (double)OH * OW < 64.0e6 && O <= 256 && smem <= 48 * 1024) {
      WgradTinyP q;
      memset(&q, 0, sizeof(q));
      q.x = dptr<const float>(x);
      q.dy = dptr<const float>(dy);
      q.N = N; q.C = C; q.H = H; q.W = W; q.K = K; q.R = R; q.S = S; q.OH = OH; q.OW = OW;
      q.stride = stride; q.pad_t = pad_t; q.pad_l = pad_l;
      q.O = (int)O;
      q.G = G < OH ? G : (OH > 0 ? OH : 1);
      int grid = N < 2 * sms_all ? N : 2 * sms_all;
      float* ws = nullptr;
      int r = pool_alloc((void**)&ws, (uint64_t)((long long)grid * O) * sizeof(float));
      if (r) return r;
      q.ws = ws;
      const int tok = prof_begin(NG_PROF_CONV_WGRAD, 2.0 * N * K * C * R * S * (double)OH * OW,
                                 4.0 * ((double)N * C * H * W + (double)K * C * R * S + (double)N * K * OH * OW));
      NG_LAUNCH(conv_wgrad_tiny_kernel, grid, 256, smem, q);
      NG_CHECK_LAUNCH();
      if (grid >= 64 && O <= 4096) {
        NG_LAUNCH(splitk_reduce_warp_kernel, (int)ceil_div(O * 32, 256), 256, 0, dptr<float>(dw), (const float*)ws, grid, O);
      } else {
        NG_LAUNCH(splitk_reduce_kernel, (int)ceil_div(O, 256), 256, 0, dptr<float>(dw), (const float*)ws, grid, O,
                  (const float*)nullptr, 1, 0);
      }
      NG_CHECK_LAUNCH();
      pool_free(ws);
      prof_end(tok);
      return 0;
    }
  }
  {
    // image-facing layer: direct kernel (see conv_wgrad_smallc_kernel)
    const int NR = C * R, RPG = NR < 3 ? NR : 3, G = (NR + RPG - 1) / RPG;
    const int px_budget = (40 * 1024 / 4) / K - 4;
    const int TR = px_budget / OW < OH ? px_budget / OW : OH;
    if (S == 3 && stride == 1 && NR <= 12 && K % 32 == 0 && K * G <= 1024 && K * G >= 64 && OW % 4 == 0 && TR >= 1 &&
        (dy & 15u) == 0 && (long long)N * OH * OW >= 4096) {
      WgradSmallP q;
      memset(&q, 0, sizeof(q));
      q.x = dptr<const float>(x);
      q.dy = dptr<const float>(dy);
      q.N = N; q.C = C; q.H = H; q.W = W; q.K = K; q.R = R; q.OH = OH; q.OW = OW;
      q.pad_t = pad_t; q.pad_l = pad_l;
      q.TR = TR;
      q.row_blocks = (int)ceil_div(OH, TR);
      q.tiles = N * q.row_blocks;
      q.NR = NR;
      q.xrows = TR + R - 1;
      q.xpitch = OW + 4;
      q.dpitch = TR * OW + 4;
      const size_t smem = sizeof(float) * ((size_t)K * q.dpitch + (size_t)C * q.xrows * q.xpitch);
      if (smem <= 48 * 1024) {
        // load and compute phases alternate inside a CTA, so several CTAs per SM overlap them:
        // as many as shared memory (<= 48 KB each) and 2048 threads allow, at most 5
        int ctas_per_sm = (int)((220u * 1024u) / (smem + 1024u));
        if (ctas_per_sm > 2048 / (K * G)) ctas_per_sm = 2048 / (K * G);
        if (ctas_per_sm > 5) ctas_per_sm = 5;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        int grid = sms_all * ctas_per_sm;
        if (grid > q.tiles) grid = q.tiles;
        const long long out_elems = (long long)K * NR * 3;
        float* ws = nullptr;
        int r = pool_alloc((void**)&ws, (uint64_t)((long long)grid * out_elems) * sizeof(float));
        if (r) return r;
        q.ws = ws;
        const int tok = prof_begin(NG_PROF_CONV_WGRAD, 2.0 * N * K * C * R * S * (double)OH * OW,
                                   4.0 * ((double)N * C * H * W + (double)K * C * R * S + (double)N * K * OH * OW));
        if (RPG == 1) NG_LAUNCH((conv_wgrad_smallc_kernel<1>), grid, K * G, smem, q);
        else if (RPG == 2) NG_LAUNCH((conv_wgrad_smallc_kernel<2>), grid, K * G, smem, q);
        else NG_LAUNCH((conv_wgrad_smallc_kernel<3>), grid, K * G, smem, q);
        NG_CHECK_LAUNCH();
        NG_LAUNCH(splitk_reduce_wide_kernel, (int)ceil_div(out_elems, 64), 512, 0, dptr<float>(dw), (const float*)ws, grid,
                  out_elems);
        NG_CHECK_LAUNCH();
        pool_free(ws);
        prof_end(tok);
        return 0;
      }
    }
  }
  WgradP p;
  memset(&p, 0, sizeof(p));
  p.x = dptr<const float>(x);
  p.dy = dptr<const float>(dy);
  p.N = N; p.C = C; p.H = H; p.W = W; p.K = K; p.R = R; p.S = S; p.OH = OH; p.OW = OW;
  p.stride = stride; p.pad_t = pad_t; p.pad_l = pad_l;
  p.CRS = C * R * S;
  p.total_kk = (long long)N * OH * OW;
  const int tile = pick_tile(p.CRS, K, false);
  const TileShape ts = tile_shape(tile);
  const int tiles = (int)(ceil_div(p.CRS, ts.bm) * ceil_div(K, ts.bn));
  const long long total_kt = ceil_div(p.total_kk, CT_BK);
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  // split the reduction so that the grid fills whole waves of 2 CTAs per SM (a ragged last wave
  // is the dominant loss for these small-output problems); every split at least 64 K-tiles deep
  // Tiny layers (the README CNN: < 64 MFLOP per pass) are a chain of exposed load latencies per CTA, not
  // FFMA work: cut them into many short splits (>= 8 K-tiles) and oversubscribe the SMs so that several
  // CTAs hide each other's latency; the partials are folded by the wide (8 threads per output) reduce.
  const bool tiny = 2.0 * N * K * C * R * S * (double)OH * OW < 64.0e6;
  const long long min_depth = tiny ? 8 : 64;
  const long long max_splits = total_kt / min_depth > 0 ? total_kt / min_depth : 1;
#ifdef NG_EMU
  const long long tiny_cap = 4;    // the host emulation runs every CUDA thread as a host thread
#else
  const long long tiny_cap = 2048;
#endif
  long long splits = tiny ? pick_splits(tiles, 4 * sms, max_splits < tiny_cap ? max_splits : tiny_cap)
                          : pick_splits(tiles, 2 * sms, max_splits < 512 ? max_splits : 512);
  p.kt_per_split = (int)ceil_div(total_kt, splits);
  splits = ceil_div(total_kt, p.kt_per_split);
  const long long out_elems = (long long)K * p.CRS;
  float* ws = nullptr;
  if (splits > 1) {
    int r = pool_alloc((void**)&ws, (uint64_t)(splits * out_elems) * sizeof(float));
    if (r) return r;
    p.out = ws;
    p.vec_store = (p.CRS % 4 == 0);
  } else {
    p.out = dptr<float>(dw);
    p.vec_store = (p.CRS % 4 == 0) && ((dw & 15u) == 0);
  }
  const int tok = prof_begin(NG_PROF_CONV_WGRAD, 2.0 * N * K * C * R * S * (double)OH * OW,
                             4.0 * ((double)N * C * H * W + (double)K * C * R * S + (double)N * K * OH * OW));
  dim3 grid((unsigned)ceil_div(p.CRS, ts.bm), (unsigned)ceil_div(K, ts.bn), (unsigned)splits);
  NG_TILE_SWITCH_NARROW(tile, NG_LAUNCH((conv_wgrad_kernel<TT>), grid, CT_THREADS, 0, p));
  NG_CHECK_LAUNCH();
  if (splits > 1) {
    if (splits >= 64 && out_elems <= 4096) {
      NG_LAUNCH(splitk_reduce_warp_kernel, (int)ceil_div(out_elems * 32, 256), 256, 0, dptr<float>(dw), (const float*)ws,
                (int)splits, out_elems);
    } else if (splits >= 32 && out_elems <= 65536) {
      NG_LAUNCH(splitk_reduce_wide_kernel, (int)ceil_div(out_elems, 64), 512, 0, dptr<float>(dw), (const float*)ws,
                (int)splits, out_elems);
    } else {
      int g = (int)ceil_div(out_elems, 256);
      if (g > sms * 8) g = sms * 8;
      NG_LAUNCH(splitk_reduce_kernel, g, 256, 0, dptr<float>(dw), (const float*)ws, (int)splits, out_elems,
                (const float*)nullptr, 1, 0);
    }
    NG_CHECK_LAUNCH();
    pool_free(ws);
  }
  prof_end(tok);
  return 0;
}

int ng_gemm(uint64_t c, uint64_t a, uint64_t b, int64_t M, int64_t N, int64_t K, int trans_a, int trans_b,
            int64_t batch, uint64_t bias, int flags) {
  NG_ENSURE_INIT();
  NG_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 1, "ng_gemm: invalid dimensions");
  NG_REQUIRE(M < 2147483647LL && N < 2147483647LL && K < 2147483647LL && batch <= 65535,
             "ng_gemm: dimension too large");
  if (M == 0 || N == 0) return 0;
#ifndef NG_EMU
  // tensor-core GEMM: TF32 mode always; FP32-accurate modes (3xTF32) from 30 MFLOP up — below ~0.2 GFLOP both engines
  // are latency-bound, but the SIMT kernel's 128 x 128 tiles leave a 512 x 256 output on 8 of 148 SMs (measured
  // 512 x 256 x 256: 50 us SIMT, 14.5 us tensor core, the q / k / v projections of BASELINE configs[3])
  const double gflops = 2.0 * (double)M * N * K;
  const bool gemm_tc = (flags & NG_F_TF32) ||
                       ((flags & (NG_F_TF32X3 | NG_F_F16X3)) && ((flags & NG_F_TC_FORCE) || gflops >= 3.0e7));
  if (batch == 1 && K > 0 && gemm_tc) {
    const int tok = prof_begin(NG_PROF_GEMM, 2.0 * (double)M * N * K, 4.0 * ((double)M * K + (double)K * N + (double)M * N));
    const int rc = umma::gemm_tf32(dptr<float>(c), dptr<const float>(a), dptr<const float>(b),
                                   bias ? dptr<const float>(bias) : nullptr, M, N, K, trans_a, trans_b,
                                   (flags & NG_F_RELU) ? 1 : 0, (flags & NG_F_TF32) ? 0 : 1);
    prof_end(tok);
    if (rc >= 0) return rc;
  }
#endif
  // tiny problems: one thread per output (see gemm_small_kernel)
#ifdef NG_EMU
  const long long small_out_cap = 4096;   // the host emulation runs every CUDA thread as a host thread
#else
  const long long small_out_cap = 262144;
#endif
  if (batch == 1 && K > 0 && (double)M * N * K <= 4.0e6 && M * N <= small_out_cap && (M * N >= 256 || K <= 4096)) {
    GemmSmallP q;
    q.a = dptr<const float>(a);
    q.b = dptr<const float>(b);
    q.c = dptr<float>(c);
    q.bias = bias ? dptr<const float>(bias) : nullptr;
    q.M = (int)M; q.N = (int)N; q.K = (int)K;
    q.relu = (flags & NG_F_RELU) ? 1 : 0;
    q.sam = trans_a ? 1 : K; q.sak = trans_a ? M : 1;
    q.sbk = trans_b ? 1 : N; q.sbn = trans_b ? K : 1;
    const int tok = prof_begin(NG_PROF_GEMM, 2.0 * (double)M * N * K, 4.0 * ((double)M * K + (double)K * N + (double)M * N));
    if (K >= 64) NG_LAUNCH((gemm_small_kernel<32>), (int)ceil_div(M * N * 32, 128), 128, 0, q);
    else NG_LAUNCH((gemm_small_kernel<1>), (int)ceil_div(M * N, 128), 128, 0, q);
    NG_CHECK_LAUNCH();
    prof_end(tok);
    return 0;
  }
  // Row-major C[M,N]: the contiguous output index is n → I = N, J = M.
  //   a(i=n, k) = B[k, n]  → !trans_b: B is [K,N], sai = 1,  sak = N ; trans_b: B is [N,K], sai = K, sak = 1
  //   b(k, j=m) = A[m, k]  → !trans_a: A is [M,K], sbj = K,  sbk = 1 ; trans_a: A is [K,M], sbj = 1, sbk = M
  GemmP p;
  memset(&p, 0, sizeof(p));
  p.a = dptr<const float>(b);
  p.b = dptr<const float>(a);
  p.bias_i = bias ? dptr<const float>(bias) : nullptr;
  p.I = (int)N; p.J = (int)M; p.K = (int)K;
  p.sai = trans_b ? K : 1; p.sak = trans_b ? 1 : N;
  p.sbj = trans_a ? 1 : K; p.sbk = trans_a ? M : 1;
  p.ldc = N;
  p.batch_a = K * N;  // stride of the operand bound to `a` (the caller's B)
  p.batch_b = M * K;
  p.batch_c = M * N;
  p.relu = (flags & NG_F_RELU) ? 1 : 0;
  const int tile = pick_tile(p.I, p.J, true);
  const TileShape ts = tile_shape(tile);
  const long long tiles = ceil_div(p.I, ts.bm) * ceil_div(p.J, ts.bn) * batch;
  const long long total_kt = ceil_div(K, CT_BK);
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  long long splits = 1;
  if (batch == 1 && tiles < sms && total_kt >= 64) {
    splits = ceil_div((long long)sms * 2, tiles);
    const long long max_splits = total_kt / 16;
    if (splits > max_splits) splits = max_splits;
    if (splits > 256) splits = 256;
    if (splits < 1) splits = 1;
  }
  p.kt_per_split = (int)ceil_div(total_kt > 0 ? total_kt : 1, splits);
  splits = total_kt > 0 ? ceil_div(total_kt, p.kt_per_split) : 1;
  p.splits = (int)splits;
  const long long out_elems = M * N;
  float* ws = nullptr;
  if (splits > 1) {
    int r = pool_alloc((void**)&ws, (uint64_t)(splits * out_elems) * sizeof(float));
    if (r) return r;
    p.out = ws;
    p.ws_stride = out_elems;
    p.vec_store = (N % 4 == 0);
  } else {
    p.out = dptr<float>(c);
    p.ws_stride = 0;
    p.vec_store = (N % 4 == 0) && ((c & 15u) == 0);
  }
  const int tok = prof_begin(NG_PROF_GEMM, 2.0 * (double)M * N * K * batch,
                             4.0 * batch * ((double)M * K + (double)K * N + (double)M * N));
  dim3 grid((unsigned)ceil_div(p.I, ts.bm), (unsigned)ceil_div(p.J, ts.bn), (unsigned)(batch * splits));
  NG_TILE_SWITCH(tile, NG_LAUNCH((gemm_kernel<TT>), grid, CT_THREADS, 0, p));
  NG_CHECK_LAUNCH();
  if (splits > 1) {
    int g = (int)ceil_div(out_elems, 256);
    if (g > sms * 8) g = sms * 8;
    NG_LAUNCH(splitk_reduce_kernel, g, 256, 0, dptr<float>(c), (const float*)ws, (int)splits, out_elems,
              p.bias_i, (int)N, p.relu);
    NG_CHECK_LAUNCH();
    pool_free(ws);
  }
  prof_end(tok);
  return 0;
}

}  // extern "C"
