// Fused log-softmax and negative-log-likelihood kernels (row-wise reductions, one warp per row).
//
// The reference builds log_softmax out of nine primitive graph nodes
// (nanograd/tensor.py:327-331: max -> reshape -> sub -> exp -> sum -> log -> reshape -> sub) and
// NLLLoss as onehot -> mul -> mean -> neg (nanograd/nn/module.py:674-686; OneHot:
// nanograd/nn/ops_cpu.py:39-49).  Note the reference's NLL divides by rows*cols, not rows.
#include "ng_common.h"

#include <math.h>

namespace ng {

__global__ void __launch_bounds__(256) logsoftmax_fwd_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                             int64_t rows, int64_t cols) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* x = in + r * cols;
    float m = -INFINITY;
    for (int64_t c = lane; c < cols; c += 32) m = fmaxf(m, x[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int64_t c = lane; c < cols; c += 32) s += expf(x[c] - m);
    s = warp_sum(s);
    const float lse = logf(s);
    float* o = out + r * cols;
    for (int64_t c = lane; c < cols; c += 32) o[c] = (x[c] - m) - lse;
  }
}

// dx = dy - exp(logp) * sum_j dy_j   (the max term cancels analytically)
__global__ void __launch_bounds__(256) logsoftmax_bwd_kernel(float* __restrict__ dx, const float* __restrict__ dy,
                                                             const float* __restrict__ logp, int64_t rows,
                                                             int64_t cols) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* g = dy + r * cols;
    const float* lp = logp + r * cols;
    float s = 0.f;
    for (int64_t c = lane; c < cols; c += 32) s += g[c];
    s = warp_sum(s);
    float* o = dx + r * cols;
    for (int64_t c = lane; c < cols; c += 32) o[c] = g[c] - expf(lp[c]) * s;
  }
}

// loss[0] = -(sum_b logp[b, label_b]) / (rows * cols); single block → deterministic order
__global__ void __launch_bounds__(256) nll_fwd_kernel(float* __restrict__ loss, const float* __restrict__ logp,
                                                       const float* __restrict__ labels, int64_t rows, int64_t cols) {
  __shared__ float scratch[32];
  float acc = 0.f;
  for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) {
    int64_t lab = (int64_t)labels[r];
    if (lab < 0) lab += cols;
    if (lab >= 0 && lab < cols) acc += logp[r * cols + lab];
  }
  const float t = block_sum(acc, scratch);
  if (threadIdx.x == 0) loss[0] = -t / (float)(rows * cols);
}

__global__ void __launch_bounds__(256) nll_bwd_kernel(float* __restrict__ dlogp, const float* __restrict__ dloss,
                                                      const float* __restrict__ labels, int64_t rows, int64_t cols) {
  const int64_t n = rows * cols;
  const float g = -dloss[0] / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    int64_t lab = (int64_t)labels[r];
    if (lab < 0) lab += cols;
    dlogp[i] = (lab == c) ? g : 0.f;
  }
}

// ---- MSELoss fused: ((p - t)^2).mean()  (module.py:689-712: Add, Neg, Pow, Sum, Mul nodes) --------------
__global__ void __launch_bounds__(256) mse_partial_kernel(float* __restrict__ partial, const float* __restrict__ p,
                                                          const float* __restrict__ t, int64_t n) {
  __shared__ float scratch[32];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = p[i] - t[i];
    acc += d * d;
  }
  const float s = block_sum(acc, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) mse_finish_kernel(float* __restrict__ loss, const float* __restrict__ partial, int m,
                                                         float inv_n) {
  __shared__ float scratch[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < m; i += blockDim.x) acc += partial[i];
  const float s = block_sum(acc, scratch);
  if (threadIdx.x == 0) loss[0] = s * inv_n;
}

// d = sign * dloss * 2 (p - t) / n   (sign = +1 for the prediction, -1 for the target)
__global__ void __launch_bounds__(256) mse_bwd_kernel(float* __restrict__ d, const float* __restrict__ dloss,
                                                      const float* __restrict__ p, const float* __restrict__ t, int64_t n,
                                                      float sign) {
  const float g = sign * dloss[0] * 2.f / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = g * (p[i] - t[i]);
}

static inline int row_grid(int64_t rows) {
  int64_t g = ceil_div(rows, 8);  // 8 warps per block
  const int64_t cap = (int64_t)(g_sm_count > 0 ? g_sm_count : 148) * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---- evaluation metrics on the device (the reference copies logits to the host every step to take an
// argmax, tests/helpers.py:310-313,326-334) --------------------------------------------------------------
// out[r] = index of the first maximum of row r (np.argmax semantics), as float32
__global__ void __launch_bounds__(256) argmax_rows_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                          int64_t rows, int64_t cols) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* row = in + r * cols;
  float best = row[0];
  int64_t arg = 0;
  for (int64_t c = 1; c < cols; ++c) {
    const float v = row[c];
    if (v > best) { best = v; arg = c; }
  }
  out[r] = (float)arg;
}

// out[0] = number of i with a[i] == b[i] (one block; exact while n < 2^24)
__global__ void __launch_bounds__(1024) count_equal_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                           const float* __restrict__ b, int64_t n) {
  __shared__ float scratch[32];
  float cnt = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) cnt += (a[i] == b[i]) ? 1.f : 0.f;
  const float total = block_sum(cnt, scratch);
  if (threadIdx.x == 0) out[0] = total;
}

}  // namespace ng

using namespace ng;

extern "C" {

int ng_logsoftmax_fwd(uint64_t out, uint64_t in, int64_t rows, int64_t cols) {
  NG_ENSURE_INIT();
  if (rows <= 0 || cols <= 0) return 0;
  NG_LAUNCH(logsoftmax_fwd_kernel, row_grid(rows), 256, 0, dptr<float>(out), dptr<const float>(in), rows, cols);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_logsoftmax_bwd(uint64_t dx, uint64_t dy, uint64_t logp, int64_t rows, int64_t cols) {
  NG_ENSURE_INIT();
  if (rows <= 0 || cols <= 0) return 0;
  NG_LAUNCH(logsoftmax_bwd_kernel, row_grid(rows), 256, 0, dptr<float>(dx), dptr<const float>(dy),
            dptr<const float>(logp), rows, cols);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_nll_fwd(uint64_t loss1, uint64_t logp, uint64_t labels, int64_t rows, int64_t cols) {
  NG_ENSURE_INIT();
  NG_REQUIRE(rows > 0 && cols > 0, "ng_nll_fwd: empty input");
  NG_LAUNCH(nll_fwd_kernel, 1, 256, 0, dptr<float>(loss1), dptr<const float>(logp), dptr<const float>(labels), rows,
            cols);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_nll_bwd(uint64_t dlogp, uint64_t dloss1, uint64_t labels, int64_t rows, int64_t cols) {
  NG_ENSURE_INIT();
  if (rows <= 0 || cols <= 0) return 0;
  int64_t g = ceil_div(rows * cols, 256);
  const int64_t cap = (int64_t)(g_sm_count > 0 ? g_sm_count : 148) * 16;
  if (g > cap) g = cap;
  NG_LAUNCH(nll_bwd_kernel, (int)g, 256, 0, dptr<float>(dlogp), dptr<const float>(dloss1),
            dptr<const float>(labels), rows, cols);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_mse_fwd(uint64_t loss1, uint64_t pred, uint64_t target, int64_t n) {
  NG_ENSURE_INIT();
  NG_REQUIRE(n > 0, "ng_mse_fwd: empty input");
  int64_t blocks = ceil_div(n, 1024);
  const int64_t cap = (int64_t)(g_sm_count > 0 ? g_sm_count : 148) * 4;
  if (blocks > cap) blocks = cap;
  float* partial = nullptr;
  int r = pool_alloc((void**)&partial, (uint64_t)blocks * sizeof(float));
  if (r) return r;
  NG_LAUNCH(mse_partial_kernel, (int)blocks, 256, 0, partial, dptr<const float>(pred), dptr<const float>(target), n);
  NG_CHECK_LAUNCH();
  NG_LAUNCH(mse_finish_kernel, 1, 256, 0, dptr<float>(loss1), (const float*)partial, (int)blocks, 1.f / (float)n);
  NG_CHECK_LAUNCH();
  pool_free(partial);
  return 0;
}

int ng_mse_bwd(uint64_t d, uint64_t dloss1, uint64_t pred, uint64_t target, int64_t n, int for_target) {
  NG_ENSURE_INIT();
  if (n <= 0) return 0;
  int64_t g = ceil_div(n, 256);
  const int64_t cap = (int64_t)(g_sm_count > 0 ? g_sm_count : 148) * 16;
  if (g > cap) g = cap;
  NG_LAUNCH(mse_bwd_kernel, (int)g, 256, 0, dptr<float>(d), dptr<const float>(dloss1), dptr<const float>(pred),
            dptr<const float>(target), n, for_target ? -1.f : 1.f);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_argmax_rows(uint64_t out, uint64_t in, int64_t rows, int64_t cols) {
  NG_ENSURE_INIT();
  NG_REQUIRE(rows >= 0 && cols > 0, "ng_argmax_rows: bad shape");
  if (rows == 0) return 0;
  NG_LAUNCH(argmax_rows_kernel, (int)ceil_div(rows, 256), 256, 0, dptr<float>(out), dptr<const float>(in), rows, cols);
  NG_CHECK_LAUNCH();
  return 0;
}

int ng_count_equal(uint64_t out1, uint64_t a, uint64_t b, int64_t n) {
  NG_ENSURE_INIT();
  NG_REQUIRE(n >= 0 && n < (1LL << 24), "ng_count_equal: n must be below 2^24 (float32 counter)");
  NG_LAUNCH(count_equal_kernel, 1, 1024, 0, dptr<float>(out1), dptr<const float>(a), dptr<const float>(b), n);
  NG_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
