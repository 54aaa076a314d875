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

template <int TN, bool STRIDED>
struct ConvGatherLoader {
  const ConvGatherP& p;
  // A: this thread stages rows k_local = khalf + 2q of column i_local
  const float* a_img;
  int y0, x0, i_local, khalf;
  bool i_valid;
  CtKPos apos[4];
  float ra[4];
  // B: elements (klane, j_local + 32q)
  const float* b_col[CtTile<TN>::NB];
  bool b_valid[CtTile<TN>::NB];
  int klane, j_local;
  CtKPos bpos;
  float rb[CtTile<TN>::NB];

  __device__ __forceinline__ ConvGatherLoader(const ConvGatherP& pp, int tile_i0, int tile_j0, int tid) : p(pp) {
    i_local = tid & 127;
    khalf = tid >> 7;
    const int i = tile_i0 + i_local;
    i_valid = i < p.M;
    const int ii = i_valid ? i : 0;
    const int n_img = ii / p.opix, pix = ii - n_img * p.opix;
    const int orow = pix / p.ow_out, ocol = pix - orow * p.ow_out;
    y0 = orow * p.y_mul + p.y_add;
    x0 = ocol * p.x_mul + p.x_add;
    a_img = p.a + (long long)n_img * p.a_img_stride;
#pragma unroll
    for (int q = 0; q < 4; ++q) apos[q].set(khalf + 2 * q, p.RS);
    klane = tid & 7;
    j_local = tid >> 3;
    bpos.set(klane, p.RS);
#pragma unroll
    for (int q = 0; q < CtTile<TN>::NB; ++q) {
      const int jl = j_local + 32 * q;
      const int j = tile_j0 + jl;
      b_valid[q] = (jl < CtTile<TN>::BN) && (j < p.J);
      b_col[q] = p.b + (long long)(b_valid[q] ? j : 0) * p.b_sj;
    }
  }

  __device__ __forceinline__ void fetch() {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v = 0.f;
      if (i_valid && apos[q].slow < p.KC) {
        int y = y0 + p.tap_dy[apos[q].fast];
        int x = x0 + p.tap_dx[apos[q].fast];
        bool ok = true;
        if (STRIDED) {
          ok = (y % p.dstride == 0) && (x % p.dstride == 0);
          y /= p.dstride;
          x /= p.dstride;
        }
        ok = ok && (unsigned)y < (unsigned)p.a_h && (unsigned)x < (unsigned)p.a_w;
        if (ok) v = a_img[(long long)apos[q].slow * p.a_chan_stride + y * p.a_w + x];
      }
      ra[q] = v;
      apos[q].advance(CT_BK, p.RS);
    }
#pragma unroll
    for (int q = 0; q < CtTile<TN>::NB; ++q) {
      float v = 0.f;
      if (b_valid[q] && bpos.slow < p.KC) v = b_col[q][(long long)bpos.slow * p.b_sc + bpos.fast];
      rb[q] = v;
    }
    bpos.advance(CT_BK, p.RS);
  }

  __device__ __forceinline__ void commit(float (*As)[CT_LDA], float (*Bs)[CtTile<TN>::LDB]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) As[khalf + 2 * q][i_local] = ra[q];
#pragma unroll
    for (int q = 0; q < CtTile<TN>::NB; ++q) {
      const int jl = j_local + 32 * q;
      if (jl < CtTile<TN>::BN) Bs[klane][jl] = rb[q];
    }
  }
};

template <int TN, bool STRIDED>
__global__ void __launch_bounds__(CT_THREADS, 2) conv_gather_kernel(const __grid_constant__ ConvGatherP p) {
  const int tid = threadIdx.x;
  const int tile_i0 = blockIdx.x * CT_BM, tile_j0 = blockIdx.y * CtTile<TN>::BN;
  const CtCoord c = ct_coord<TN>(tid);
  float acc[8][TN];
  {
    ConvGatherLoader<TN, STRIDED> ld(p, tile_i0, tile_j0, tid);
    const int num_kt = (p.KC * p.RS + CT_BK - 1) / CT_BK;
    ct_mainloop<TN>(ld, num_kt, c, acc);
  }
  // epilogue: out[(img*J + j)*opix + pix] (+ bias[j]) ; I = img*opix + pix is contiguous per (img, j)
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    const int i = tile_i0 + c.i0 + 32 * f;
    if (i >= p.M) continue;
    if (p.vec_store) {
      const int n_img = i / p.opix, pix = i - n_img * p.opix;
#pragma unroll
      for (int n = 0; n < TN; ++n) {
        const int j = tile_j0 + ct_jcol<TN>(c, n);
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
        for (int n = 0; n < TN; ++n) {
          const int j = tile_j0 + ct_jcol<TN>(c, n);
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
// dense epilogue shared by wgrad and GEMM: out[z][j*ldc + i]
// =====================================================================================
template <int TN>
__device__ __forceinline__ void dense_store(float* out, long long ldc, int I, int J, int tile_i0, int tile_j0,
                                            const CtCoord& c, const float (&acc)[8][TN], const float* bias_i,
                                            int relu, int vec_store) {
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    const int i = tile_i0 + c.i0 + 32 * f;
    if (i >= I) continue;
#pragma unroll
    for (int n = 0; n < TN; ++n) {
      const int j = tile_j0 + ct_jcol<TN>(c, n);
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

template <int TN>
struct WgradLoader {
  const WgradP& p;
  int klane, il;  // this thread stages (klane, il + 32q) of both tiles
  // A (x gather): per q the filter position of column i
  int a_off[4], a_dy[4], a_dx[4];
  bool a_valid[4];
  // B (dy): per q the output channel
  long long b_off[CtTile<TN>::NB];
  bool b_valid[CtTile<TN>::NB];
  // reduction position (n, oh, ow) of row klane
  int n, oh, ow;
  long long kk, kk_end;
  float ra[4], rb[CtTile<TN>::NB];

  __device__ __forceinline__ WgradLoader(const WgradP& pp, int tile_i0, int tile_j0, int tid, long long kk_begin,
                                         long long kk_stop)
      : p(pp) {
    klane = tid & 7;
    il = tid >> 3;
    const int RS = p.R * p.S;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
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
    for (int q = 0; q < CtTile<TN>::NB; ++q) {
      const int jl = il + 32 * q;
      const int j = tile_j0 + jl;
      b_valid[q] = (jl < CtTile<TN>::BN) && (j < p.K);
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
    for (int q = 0; q < 4; ++q) {
      float v = 0.f;
      const int y = yb + a_dy[q], x = xb + a_dx[q];
      if (k_ok && a_valid[q] && (unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W)
        v = x_img[a_off[q] + y * p.W + x];
      ra[q] = v;
    }
    const float* dy_img = p.dy + (long long)n * p.K * p.OH * p.OW + oh * p.OW + ow;
#pragma unroll
    for (int q = 0; q < CtTile<TN>::NB; ++q) {
      float v = 0.f;
      if (k_ok && b_valid[q]) v = dy_img[b_off[q]];
      rb[q] = v;
    }
    kk += CT_BK;
    ow += CT_BK;
    while (ow >= p.OW) { ow -= p.OW; ++oh; }
    while (oh >= p.OH) { oh -= p.OH; ++n; }
  }

  __device__ __forceinline__ void commit(float (*As)[CT_LDA], float (*Bs)[CtTile<TN>::LDB]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) As[klane][il + 32 * q] = ra[q];
#pragma unroll
    for (int q = 0; q < CtTile<TN>::NB; ++q) {
      const int jl = il + 32 * q;
      if (jl < CtTile<TN>::BN) Bs[klane][jl] = rb[q];
    }
  }
};

template <int TN>
__global__ void __launch_bounds__(CT_THREADS, 2) conv_wgrad_kernel(const __grid_constant__ WgradP p) {
  const int tid = threadIdx.x;
  const int tile_i0 = blockIdx.x * CT_BM, tile_j0 = blockIdx.y * CtTile<TN>::BN;
  const CtCoord c = ct_coord<TN>(tid);
  const long long kk_begin = (long long)blockIdx.z * p.kt_per_split * CT_BK;
  long long kk_stop = kk_begin + (long long)p.kt_per_split * CT_BK;
  if (kk_stop > p.total_kk) kk_stop = p.total_kk;
  const int num_kt = (int)((kk_stop - kk_begin + CT_BK - 1) / CT_BK);
  float acc[8][TN];
  {
    WgradLoader<TN> ld(p, tile_i0, tile_j0, tid, kk_begin, kk_stop);
    ct_mainloop<TN>(ld, num_kt, c, acc);
  }
  float* out = p.out + (long long)blockIdx.z * p.K * p.CRS;
  dense_store<TN>(out, p.CRS, p.CRS, p.K, tile_i0, tile_j0, c, acc, nullptr, 0, p.vec_store);
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

template <int TN>
struct GemmLoader {
  const GemmP& p;
  const float* a;
  const float* b;
  int tile_i0, tile_j0, tid;
  int k0, k_end;
  bool a_icontig, b_jcontig;
  float ra[4], rb[CtTile<TN>::NB];

  __device__ __forceinline__ GemmLoader(const GemmP& pp, const float* aa, const float* bb, int ti0, int tj0, int t,
                                        int kb, int ke)
      : p(pp), a(aa), b(bb), tile_i0(ti0), tile_j0(tj0), tid(t), k0(kb), k_end(ke) {
    a_icontig = (p.sai == 1);
    b_jcontig = (p.sbj == 1);
  }
  __device__ __forceinline__ void a_coord(int q, int& kl, int& il) const {
    if (a_icontig) { il = tid & 127; kl = (tid >> 7) + 2 * q; }
    else { kl = tid & 7; il = (tid >> 3) + 32 * q; }
  }
  __device__ __forceinline__ void b_coord(int q, int& kl, int& jl) const {
    const int e = tid + CT_THREADS * q;
    if (b_jcontig) { jl = e % CtTile<TN>::BN; kl = e / CtTile<TN>::BN; }
    else { kl = e & 7; jl = e >> 3; }
  }
  __device__ __forceinline__ void fetch() {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int kl, il;
      a_coord(q, kl, il);
      const int i = tile_i0 + il, k = k0 + kl;
      ra[q] = (i < p.I && k < k_end) ? a[(long long)i * p.sai + (long long)k * p.sak] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < CtTile<TN>::NB; ++q) {
      int kl, jl;
      b_coord(q, kl, jl);
      const int j = tile_j0 + jl, k = k0 + kl;
      rb[q] = (kl < CT_BK && jl < CtTile<TN>::BN && j < p.J && k < k_end)
                  ? b[(long long)j * p.sbj + (long long)k * p.sbk] : 0.f;
    }
    k0 += CT_BK;
  }
  __device__ __forceinline__ void commit(float (*As)[CT_LDA], float (*Bs)[CtTile<TN>::LDB]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int kl, il;
      a_coord(q, kl, il);
      As[kl][il] = ra[q];
    }
#pragma unroll
    for (int q = 0; q < CtTile<TN>::NB; ++q) {
      int kl, jl;
      b_coord(q, kl, jl);
      if (kl < CT_BK && jl < CtTile<TN>::BN) Bs[kl][jl] = rb[q];
    }
  }
};

template <int TN>
__global__ void __launch_bounds__(CT_THREADS, 2) gemm_kernel(const __grid_constant__ GemmP p) {
  const int tid = threadIdx.x;
  const int tile_i0 = blockIdx.x * CT_BM, tile_j0 = blockIdx.y * CtTile<TN>::BN;
  const int batch = blockIdx.z / p.splits, split = blockIdx.z - batch * p.splits;
  const CtCoord c = ct_coord<TN>(tid);
  const int kb = split * p.kt_per_split * CT_BK;
  int ke = kb + p.kt_per_split * CT_BK;
  if (ke > p.K) ke = p.K;
  const int num_kt = (ke - kb + CT_BK - 1) / CT_BK;
  float acc[8][TN];
  {
    GemmLoader<TN> ld(p, p.a + batch * p.batch_a, p.b + batch * p.batch_b, tile_i0, tile_j0, tid, kb, ke);
    ct_mainloop<TN>(ld, num_kt, c, acc);
  }
  float* out = p.out + (long long)split * p.ws_stride + batch * p.batch_c;
  const bool final_pass = (p.splits == 1);
  dense_store<TN>(out, p.ldc, p.I, p.J, tile_i0, tile_j0, c, acc, final_pass ? p.bias_i : nullptr,
                  final_pass ? p.relu : 0, p.vec_store);
}

static inline int pick_tn(int J) {
  if (J <= 16) return 1;
  if (J <= 32) return 2;
  if (J <= 64) return 4;
  return 8;
}

static int launch_conv_gather(const ConvGatherP& p, bool strided) {
  const int tn = pick_tn(p.J);
  const int bn = 16 * tn;
  dim3 grid((unsigned)ceil_div(p.M, CT_BM), (unsigned)ceil_div(p.J, bn), 1);
  if (strided) {
    switch (tn) {
      case 1: NG_LAUNCH((conv_gather_kernel<1, true>), grid, CT_THREADS, 0, p); break;
      case 2: NG_LAUNCH((conv_gather_kernel<2, true>), grid, CT_THREADS, 0, p); break;
      case 4: NG_LAUNCH((conv_gather_kernel<4, true>), grid, CT_THREADS, 0, p); break;
      default: NG_LAUNCH((conv_gather_kernel<8, true>), grid, CT_THREADS, 0, p); break;
    }
  } else {
    switch (tn) {
      case 1: NG_LAUNCH((conv_gather_kernel<1, false>), grid, CT_THREADS, 0, p); break;
      case 2: NG_LAUNCH((conv_gather_kernel<2, false>), grid, CT_THREADS, 0, p); break;
      case 4: NG_LAUNCH((conv_gather_kernel<4, false>), grid, CT_THREADS, 0, p); break;
      default: NG_LAUNCH((conv_gather_kernel<8, false>), grid, CT_THREADS, 0, p); break;
    }
  }
  NG_CHECK_LAUNCH();
  return 0;
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
  const int rc = launch_conv_gather(p, false);
  prof_end(tok);
  return rc;
}

int ng_conv2d_dgrad(uint64_t dx, uint64_t dy, uint64_t w, int N, int C, int H, int W, int K, int R, int S,
                    int stride, int pad_t, int pad_l, int OH, int OW, int flags) {
  NG_ENSURE_INIT();
  (void)flags;
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
  const int rc = launch_conv_gather(p, stride > 1);
  prof_end(tok);
  return rc;
}

int ng_conv2d_wgrad(uint64_t dw, uint64_t dy, uint64_t x, int N, int C, int H, int W, int K, int R, int S,
                    int stride, int pad_t, int pad_l, int OH, int OW, int flags) {
  NG_ENSURE_INIT();
  (void)flags;
  if (check_conv_dims("ng_conv2d_wgrad", N, C, H, W, K, R, S, stride, pad_t, pad_l, OH, OW)) return 1;
  WgradP p;
  memset(&p, 0, sizeof(p));
  p.x = dptr<const float>(x);
  p.dy = dptr<const float>(dy);
  p.N = N; p.C = C; p.H = H; p.W = W; p.K = K; p.R = R; p.S = S; p.OH = OH; p.OW = OW;
  p.stride = stride; p.pad_t = pad_t; p.pad_l = pad_l;
  p.CRS = C * R * S;
  p.total_kk = (long long)N * OH * OW;
  const int tn = pick_tn(K);
  const int bn = 16 * tn;
  const int tiles = (int)(ceil_div(p.CRS, CT_BM) * ceil_div(K, bn));
  const long long total_kt = ceil_div(p.total_kk, CT_BK);
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  // enough splits for ~2 CTAs per SM over 2 waves, each split at least 64 K-tiles deep
  long long splits = ceil_div((long long)sms * 4, tiles);
  const long long max_splits = total_kt / 64 > 0 ? total_kt / 64 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits > 512) splits = 512;
  if (splits < 1) splits = 1;
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
  dim3 grid((unsigned)ceil_div(p.CRS, CT_BM), (unsigned)ceil_div(K, bn), (unsigned)splits);
  switch (tn) {
    case 1: NG_LAUNCH((conv_wgrad_kernel<1>), grid, CT_THREADS, 0, p); break;
    case 2: NG_LAUNCH((conv_wgrad_kernel<2>), grid, CT_THREADS, 0, p); break;
    case 4: NG_LAUNCH((conv_wgrad_kernel<4>), grid, CT_THREADS, 0, p); break;
    default: NG_LAUNCH((conv_wgrad_kernel<8>), grid, CT_THREADS, 0, p); break;
  }
  NG_CHECK_LAUNCH();
  if (splits > 1) {
    int g = (int)ceil_div(out_elems, 256);
    if (g > sms * 8) g = sms * 8;
    NG_LAUNCH(splitk_reduce_kernel, g, 256, 0, dptr<float>(dw), (const float*)ws, (int)splits, out_elems,
              (const float*)nullptr, 1, 0);
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
  const int tn = pick_tn(p.J);
  const int bn = 16 * tn;
  const long long tiles = ceil_div(p.I, CT_BM) * ceil_div(p.J, bn) * batch;
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
  dim3 grid((unsigned)ceil_div(p.I, CT_BM), (unsigned)ceil_div(p.J, bn), (unsigned)(batch * splits));
  switch (tn) {
    case 1: NG_LAUNCH((gemm_kernel<1>), grid, CT_THREADS, 0, p); break;
    case 2: NG_LAUNCH((gemm_kernel<2>), grid, CT_THREADS, 0, p); break;
    case 4: NG_LAUNCH((gemm_kernel<4>), grid, CT_THREADS, 0, p); break;
    default: NG_LAUNCH((gemm_kernel<8>), grid, CT_THREADS, 0, p); break;
  }
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
