// Reductions over arbitrary axis sets (sum / max / min) and the per-channel sum used for bias
// gradients.  Deterministic: two passes through a partials buffer, never atomics.
//
// Replaces the OpenCL `reduce` program of the reference, which runs one serial loop per output
// element (nanograd/nn/ops_gpu.py:87-112,133-155), and unbroadcast (ops_gpu.py:47-49,
// ops_cpu.py:12-18).
#include "ng_common.h"

#include <math.h>

namespace ng {

constexpr int kRMaxDims = 8;

struct ReduceDesc {
  int nk, nr;                  // number of kept / reduced (collapsed) dims
  int64_t ksize[kRMaxDims];    // kept dim sizes, outermost first
  int64_t kstride[kRMaxDims];  // input element strides of kept dims
  int64_t rsize[kRMaxDims];
  int64_t rstride[kRMaxDims];
  int64_t O, R;                // total outputs, total reduced elements per output
  int64_t chunk;               // reduced elements per split
  int S;                       // number of splits of R
};

template <int OP>
__device__ __forceinline__ float r_identity() {
  if (OP == NG_R_SUM) return 0.f;
  if (OP == NG_R_MAX) return -INFINITY;
  return INFINITY;
}
template <int OP>
__device__ __forceinline__ float r_combine(float a, float b) {
  if (OP == NG_R_SUM) return a + b;
  if (OP == NG_R_MAX) return fmaxf(a, b);
  return fminf(a, b);
}

__device__ __forceinline__ int64_t kept_offset(const ReduceDesc& d, int64_t o) {
  int64_t off = 0;
#pragma unroll
  for (int k = kRMaxDims - 1; k >= 0; --k) {
    if (k < d.nk) {
      const int64_t q = o / d.ksize[k];
      off += (o - q * d.ksize[k]) * d.kstride[k];
      o = q;
    }
  }
  return off;
}
__device__ __forceinline__ int64_t red_offset(const ReduceDesc& d, int64_t r) {
  if (d.nr == 1) return r * d.rstride[0];
  int64_t off = 0;
#pragma unroll
  for (int k = kRMaxDims - 1; k >= 0; --k) {
    if (k < d.nr) {
      const int64_t q = r / d.rsize[k];
      off += (r - q * d.rsize[k]) * d.rstride[k];
      r = q;
    }
  }
  return off;
}

// One thread per (output, split): lanes run along the outputs (coalesced when the innermost
// input dim is kept; also used when R is tiny, e.g. pooling windows).
// partial layout: [S][O]
template <int OP>
__global__ void __launch_bounds__(256) reduce_thread_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                            ReduceDesc d) {
  const int64_t total = d.O * d.S;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total;
       w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = w / d.O, o = w - s * d.O;
    const float* base = in + kept_offset(d, o);
    const int64_t r0 = s * d.chunk;
    int64_t r1 = r0 + d.chunk;
    if (r1 > d.R) r1 = d.R;
    float acc = r_identity<OP>();
    for (int64_t r = r0; r < r1; ++r) acc = r_combine<OP>(acc, base[red_offset(d, r)]);
    out[w] = acc;
  }
}

// One block per (output, split): threads run along the reduced elements (coalesced when the
// innermost input dim is reduced).
template <int OP>
__global__ void __launch_bounds__(256) reduce_block_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                           ReduceDesc d) {
  __shared__ float scratch[32];
  const int64_t total = d.O * d.S;
  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int64_t s = w / d.O, o = w - s * d.O;
    const float* base = in + kept_offset(d, o);
    const int64_t r0 = s * d.chunk;
    int64_t r1 = r0 + d.chunk;
    if (r1 > d.R) r1 = d.R;
    float acc = r_identity<OP>();
    for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x)
      acc = r_combine<OP>(acc, base[red_offset(d, r)]);
    // block combine
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc = r_combine<OP>(acc, __shfl_xor_sync(0xffffffffu, acc, off));
    __syncthreads();
    if (lane == 0) scratch[warp] = acc;
    __syncthreads();
    if (warp == 0) {
      float t = (lane < (int)(blockDim.x >> 5)) ? scratch[lane] : r_identity<OP>();
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) t = r_combine<OP>(t, __shfl_xor_sync(0xffffffffu, t, off));
      if (lane == 0) out[w] = t;
    }
  }
}

template <int OP>
static int run_reduce(float* out, const float* in, ReduceDesc d, bool inner_reduced) {
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  const bool use_block = inner_reduced && d.R >= 128;
  // choose the split so that the grid has a few waves of work
  int64_t S = 1;
  if (use_block) {
    const int64_t want_blocks = (int64_t)sms * 8;
    if (d.O < want_blocks && d.R > 8192) {
      S = want_blocks / d.O;
      const int64_t maxS = ceil_div(d.R, 4096);
      if (S > maxS) S = maxS;
      if (S > 1024) S = 1024;
      if (S < 1) S = 1;
    }
  } else {
    const int64_t want_threads = (int64_t)sms * 2048;
    if (d.O < want_threads && d.R > 256) {
      S = want_threads / d.O;
      const int64_t maxS = ceil_div(d.R, 64);
      if (S > maxS) S = maxS;
      if (S > 1024) S = 1024;
      if (S < 1) S = 1;
    }
  }
  d.S = (int)S;
  d.chunk = ceil_div(d.R, S);
  d.S = (int)ceil_div(d.R, d.chunk);
  float* dst = out;
  float* partial = nullptr;
  if (d.S > 1) {
    int r = pool_alloc((void**)&partial, (uint64_t)d.S * d.O * sizeof(float));
    if (r) return r;
    dst = partial;
  }
  const int64_t work = d.O * d.S;
  if (use_block) {
    int64_t grid = work < (int64_t)sms * 16 ? work : (int64_t)sms * 16;
    NG_LAUNCH((reduce_block_kernel<OP>), (int)grid, 256, 0, dst, in, d);
  } else {
    int64_t grid = ceil_div(work, 256);
    if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
    NG_LAUNCH((reduce_thread_kernel<OP>), (int)grid, 256, 0, dst, in, d);
  }
  NG_CHECK_LAUNCH();
  if (d.S > 1) {
    // second pass: out[o] = combine over s of partial[s][o]
    ReduceDesc e;
    memset(&e, 0, sizeof(e));
    e.nk = 1; e.ksize[0] = d.O; e.kstride[0] = 1;
    e.nr = 1; e.rsize[0] = d.S; e.rstride[0] = d.O;
    e.O = d.O; e.R = d.S; e.S = 1; e.chunk = d.S;
    int64_t grid = ceil_div(e.O, 256);
    if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
    NG_LAUNCH((reduce_thread_kernel<OP>), (int)grid, 256, 0, out, (const float*)partial, e);
    NG_CHECK_LAUNCH();
    pool_free(partial);
  }
  return 0;
}

// out[s][c] = sum over n in split s, p of in[n][c][p]; grid = (C, S)
// A warp walks whole (image, channel) planes, its lanes the contiguous pixels: coalesced, and no index
// division anywhere (the per-element 64-bit div/mod of a flat loop cost 36 us on the 11x11 planes of the
// README CNN's bias gradient).  Two planes in flight per warp.
__global__ void __launch_bounds__(256) channel_sum_kernel(float* __restrict__ partial, const float* __restrict__ in,
                                                          int64_t N, int64_t C, int64_t P, int64_t n_per_split,
                                                          int vec_ok) {
  __shared__ float scratch[32];
  const int64_t c = blockIdx.x, s = blockIdx.y;
  const int64_t n0 = s * n_per_split;
  int64_t n1 = n0 + n_per_split;
  if (n1 > N) n1 = N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float acc0 = 0.f, acc1 = 0.f;
  for (int64_t n = n0 + warp; n < n1; n += 2 * nwarps) {
    const float* r0 = in + (n * C + c) * P;
    const bool two = n + nwarps < n1;
    const float* r1 = two ? in + ((n + nwarps) * C + c) * P : r0;
    if (vec_ok) {
      const int64_t P4 = P >> 2;
      for (int64_t q = lane; q < P4; q += 32) {
        const float4 v = reinterpret_cast<const float4*>(r0)[q];
        const float4 w = two ? reinterpret_cast<const float4*>(r1)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        acc0 += (v.x + v.y) + (v.z + v.w);
        acc1 += (w.x + w.y) + (w.z + w.w);
      }
    } else {
      for (int64_t q = lane; q < P; q += 32) {
        acc0 += r0[q];
        if (two) acc1 += r1[q];
      }
    }
  }
  const float t = block_sum(acc0 + acc1, scratch);
  if (threadIdx.x == 0) partial[s * C + c] = t;
}

}  // namespace ng

using namespace ng;

extern "C" {

int ng_reduce(int op, uint64_t out, uint64_t in, int ndim, const int64_t* shape, const int32_t* mask) {
  NG_ENSURE_INIT();
  NG_REQUIRE(ndim >= 0 && ndim <= 16, "ng_reduce: ndim %d unsupported", ndim);
  // contiguous strides, drop size-1 dims, merge neighbours with the same flag
  int64_t strides[16];
  {
    int64_t st = 1;
    for (int i = ndim - 1; i >= 0; --i) { strides[i] = st; st *= shape[i]; }
  }
  int64_t csize[16], cstride[16];
  int cmask[16];
  int nc = 0;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] == 1) continue;
    const int m = mask[i] ? 1 : 0;
    if (nc > 0 && cmask[nc - 1] == m) {
      csize[nc - 1] *= shape[i];
      cstride[nc - 1] = strides[i];
    } else {
      csize[nc] = shape[i]; cstride[nc] = strides[i]; cmask[nc] = m; nc++;
    }
  }
  ReduceDesc d;
  memset(&d, 0, sizeof(d));
  d.O = 1; d.R = 1;
  for (int i = 0; i < nc; ++i) {
    if (cmask[i]) {
      NG_REQUIRE(d.nr < kRMaxDims, "ng_reduce: too many reduced dim groups");
      d.rsize[d.nr] = csize[i]; d.rstride[d.nr] = cstride[i]; d.nr++; d.R *= csize[i];
    } else {
      NG_REQUIRE(d.nk < kRMaxDims, "ng_reduce: too many kept dim groups");
      d.ksize[d.nk] = csize[i]; d.kstride[d.nk] = cstride[i]; d.nk++; d.O *= csize[i];
    }
  }
  if (d.O == 0) return 0;
  if (d.nr == 0) {  // nothing to reduce (all reduced axes had size 1): copy
    d.nr = 1; d.rsize[0] = 1; d.rstride[0] = 0; d.R = 1;
  }
  NG_REQUIRE(d.R > 0, "ng_reduce: reduction over an empty axis");
  const bool inner_reduced = nc > 0 && cmask[nc - 1] == 1;
  switch (op) {
    case NG_R_SUM: return run_reduce<NG_R_SUM>(dptr<float>(out), dptr<const float>(in), d, inner_reduced);
    case NG_R_MAX: return run_reduce<NG_R_MAX>(dptr<float>(out), dptr<const float>(in), d, inner_reduced);
    case NG_R_MIN: return run_reduce<NG_R_MIN>(dptr<float>(out), dptr<const float>(in), d, inner_reduced);
    default: return set_error("ng_reduce: unknown op %d", op);
  }
}

int ng_channel_sum(uint64_t out, uint64_t in, int64_t N, int64_t C, int64_t P) {
  NG_ENSURE_INIT();
  if (C <= 0) return 0;
  NG_REQUIRE(N > 0 && P > 0, "ng_channel_sum: empty reduction");
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  int64_t S = ceil_div((int64_t)sms * 4, C);
  if (S > N) S = N;
  // keep at least ~16K elements per block so the second pass stays negligible
  while (S > 1 && (N / S) * P < 16384) S--;
  const int64_t n_per = ceil_div(N, S);
  S = ceil_div(N, n_per);
  NG_REQUIRE(S <= 65535, "ng_channel_sum: split too large");
  const int vec = (P % 4 == 0) && ((in & 15u) == 0);
  float* dst = dptr<float>(out);
  float* partial = nullptr;
  if (S > 1) {
    int r = pool_alloc((void**)&partial, (uint64_t)S * C * sizeof(float));
    if (r) return r;
    dst = partial;
  }
  NG_LAUNCH(channel_sum_kernel, dim3((unsigned)C, (unsigned)S), 256, 0, dst, dptr<const float>(in), N, C, P,
            n_per, vec);
  NG_CHECK_LAUNCH();
  if (S > 1) {
    ReduceDesc e;
    memset(&e, 0, sizeof(e));
    e.nk = 1; e.ksize[0] = C; e.kstride[0] = 1;
    e.nr = 1; e.rsize[0] = S; e.rstride[0] = C;
    e.O = C; e.R = S; e.S = 1; e.chunk = S;
    NG_LAUNCH((reduce_thread_kernel<NG_R_SUM>), (int)ceil_div(C, 256), 256, 0, dptr<float>(out),
              (const float*)partial, e);
    NG_CHECK_LAUNCH();
    pool_free(partial);
  }
  return 0;
}

}  // extern "C"
