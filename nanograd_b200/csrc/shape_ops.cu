// Data-movement kernels: N-d slice with zero fill (also padding), axis permutation, one-hot.
//
// Replaces the reference's `gslice`, `perm` and `one_hot` OpenCL programs
// (nanograd/nn/ops_gpu.py:162-208,234-248) — all three are rebuilt from source on every call
// there — and states inner_slice / OneHot of ops_cpu.py:20-49.
#include "ng_common.h"

namespace ng {

constexpr int kSMaxDims = 8;

struct SliceDesc {
  int ndim;
  int64_t out_shape[kSMaxDims];
  int64_t in_shape[kSMaxDims];
  int64_t in_stride[kSMaxDims];
  int64_t start[kSMaxDims];
};

__global__ void __launch_bounds__(256) slice_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                    int64_t n, SliceDesc d) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t rem = i, off = 0;
    bool inside = true;
#pragma unroll
    for (int k = kSMaxDims - 1; k >= 0; --k) {
      if (k < d.ndim) {
        const int64_t q = rem / d.out_shape[k];
        const int64_t idx = rem - q * d.out_shape[k] + d.start[k];
        rem = q;
        inside = inside && idx >= 0 && idx < d.in_shape[k];
        off += idx * d.in_stride[k];
      }
    }
    out[i] = inside ? in[off] : 0.f;
  }
}

struct PermDesc {
  int ndim;
  int64_t out_shape[kSMaxDims];
  int64_t in_stride_for_out[kSMaxDims];  // input stride of the axis that lands on output axis k
};

__global__ void __launch_bounds__(256) permute_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                      int64_t n, PermDesc d) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t rem = i, off = 0;
#pragma unroll
    for (int k = kSMaxDims - 1; k >= 0; --k) {
      if (k < d.ndim) {
        const int64_t q = rem / d.out_shape[k];
        off += (rem - q * d.out_shape[k]) * d.in_stride_for_out[k];
        rem = q;
      }
    }
    out[i] = in[off];
  }
}

// 2-D transpose through a padded shared-memory tile: both the read and the write are coalesced.
// in is [rows][cols], out is [cols][rows]; grid.z batches independent matrices.
__global__ void __launch_bounds__(256) transpose2d_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                          int rows, int cols) {
  __shared__ float tile[32][33];
  const int64_t mat = (int64_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int r = r0 + ty + j, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + j][tx] = in[mat + (int64_t)r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + ty + j, r = r0 + tx;
    if (r < rows && c < cols) out[mat + (int64_t)c * rows + r] = tile[tx][ty + j];
  }
}

__global__ void __launch_bounds__(256) onehot_kernel(float* __restrict__ out, const float* __restrict__ labels,
                                                     int64_t rows, int64_t classes) {
  const int64_t n = rows * classes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / classes, c = i - r * classes;
    // labels are float32-encoded integers; the cast truncates exactly like astype(int), ops_cpu.py:42
    int64_t lab = (int64_t)labels[r];
    if (lab < 0) lab += classes;  // NumPy fancy indexing wraps negative indices
    out[i] = (lab == c) ? 1.f : 0.f;
  }
}

static inline int sgrid(int64_t n) {
  int64_t g = ceil_div(n, 256);
  const int64_t cap = (int64_t)(g_sm_count > 0 ? g_sm_count : 148) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace ng

using namespace ng;

extern "C" {

int ng_slice(uint64_t out, uint64_t in, int ndim, const int64_t* in_shape, const int64_t* out_shape,
             const int64_t* start) {
  NG_ENSURE_INIT();
  NG_REQUIRE(ndim >= 1 && ndim <= kSMaxDims, "ng_slice: ndim %d not in [1, %d]", ndim, kSMaxDims);
  SliceDesc d;
  memset(&d, 0, sizeof(d));
  d.ndim = ndim;
  int64_t n = 1, st = 1;
  for (int i = ndim - 1; i >= 0; --i) {
    d.out_shape[i] = out_shape[i];
    d.in_shape[i] = in_shape[i];
    d.in_stride[i] = st;
    d.start[i] = start[i];
    st *= in_shape[i];
    n *= out_shape[i];
  }
  if (n <= 0) return 0;
  NG_LAUNCH(slice_kernel, sgrid(n), 256, 0, dptr<float>(out), dptr<const float>(in), n, d);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_permute(uint64_t out, uint64_t in, int ndim, const int64_t* in_shape, const int32_t* perm) {
  NG_ENSURE_INIT();
  NG_REQUIRE(ndim >= 1 && ndim <= kSMaxDims, "ng_permute: ndim %d not in [1, %d]", ndim, kSMaxDims);
  int64_t in_stride[kSMaxDims];
  int64_t n = 1;
  {
    int64_t st = 1;
    for (int i = ndim - 1; i >= 0; --i) { in_stride[i] = st; st *= in_shape[i]; }
    n = st;
  }
  if (n <= 0) return 0;
  bool seen[kSMaxDims] = {false};
  for (int i = 0; i < ndim; ++i) {
    NG_REQUIRE(perm[i] >= 0 && perm[i] < ndim && !seen[perm[i]], "ng_permute: invalid permutation");
    seen[perm[i]] = true;
  }
  // batched 2-D transpose fast path: perm swaps the last two axes and keeps the others
  bool last2 = ndim >= 2 && perm[ndim - 1] == ndim - 2 && perm[ndim - 2] == ndim - 1;
  for (int i = 0; i + 2 < ndim && last2; ++i) last2 = (perm[i] == i);
  if (last2 && in_shape[ndim - 2] <= 0x7fffffff && in_shape[ndim - 1] <= 0x7fffffff) {
    const int rows = (int)in_shape[ndim - 2], cols = (int)in_shape[ndim - 1];
    int64_t batch = 1;
    for (int i = 0; i + 2 < ndim; ++i) batch *= in_shape[i];
    if (batch <= 65535 && ceil_div(rows, 32) <= 65535) {
      dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32), (unsigned)batch);
      NG_LAUNCH(transpose2d_kernel, grid, 256, 0, dptr<float>(out), dptr<const float>(in), rows, cols);
      NG_CHECK_LAUNCH();
      return 0;
    }
  }
  PermDesc d;
  memset(&d, 0, sizeof(d));
  d.ndim = ndim;
  for (int k = 0; k < ndim; ++k) {
    d.out_shape[k] = in_shape[perm[k]];
    d.in_stride_for_out[k] = in_stride[perm[k]];
  }
  NG_LAUNCH(permute_kernel, sgrid(n), 256, 0, dptr<float>(out), dptr<const float>(in), n, d);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_onehot(uint64_t out, uint64_t labels, int64_t rows, int64_t classes) {
  NG_ENSURE_INIT();
  const int64_t n = rows * classes;
  if (n <= 0) return 0;
  NG_LAUNCH(onehot_kernel, sgrid(n), 256, 0, dptr<float>(out), dptr<const float>(labels), rows, classes);
  NG_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
