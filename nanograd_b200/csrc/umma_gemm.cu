// Tensor-core GEMM (tcgen05 + TMA + TMEM) for Linear / matmul: the contraction core of umma_conv.cu
// without the convolution address generator.
//
//   C[M,N] (row-major) = op(A) @ op(B) (+ bias[N]) (+ ReLU),  op = optional transpose, FP32 in / out.
//
// Both operands are read in place by TMA, whatever their storage order:
//   * an operand whose reduction index k is contiguous (A not transposed, B transposed: x @ W.T of
//     nn.Linear, module.py:248-266) is a K-major UMMA operand: 2-D box {32 k, rows}, 128B swizzle;
//   * an operand whose M/N index is contiguous (A transposed, B not transposed: the two backward
//     products of ops_cpu.py:193-204) is an MN-major operand: the row-major matrix is declared as
//     (32, K, rows/32) so one 3-D box {32, 32 k, groups} lands as [group][k][32] in the
//     128B/32B-atom swizzle, the only MN-major layout the tensor core takes for 32-bit data.
// No transpose kernel and no staging copy.  TF32 mode rounds nothing (the tensor core truncates);
// the FP32-accurate 3xTF32 mode splits hi/lo in shared memory exactly like the convolution kernels.
// Small outputs split the reduction over CTAs (workspace + fixed-order reduce with the bias/ReLU epilogue).
#include "umma_common.cuh"

#ifndef NG_EMU
#include <cuda_runtime.h>
#include <stdlib.h>

namespace ng {
namespace umma {

constexpr int GM_THREADS = 192;
constexpr int GM_THREADS_X3 = 320;
constexpr int GM_MAX_STAGES = 8;
constexpr int GM_KCHUNK = 32;

struct GemmUP {
  CUtensorMap tm_a, tm_b;
  float* out;          // C, or the split-K workspace [splits][M][N]
  const float* bias;   // applied here only when splits == 1
  int M, N, K;
  int n_tile, m_tiles, n_tiles, splits, chunks_per_split, k_chunks;
  int a_mn, b_mn;      // operand is MN-major
  uint32_t idesc, stage_a_bytes, stage_b_bytes, tmem_cols;
  int stages, x3, relu, vec_store;
};

static inline int pow2_ceil_i(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

__global__ void __launch_bounds__(GM_THREADS_X3, 1) umma_gemm_kernel(const __grid_constant__ GemmUP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = p.stage_a_bytes + p.stage_b_bytes;
  uint8_t* lo_ring = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(lo_ring + (p.x3 ? 2u * stage_bytes : 0u));
  uint64_t* empty_bar = full_bar + GM_MAX_STAGES;
  uint64_t* split_bar = empty_bar + GM_MAX_STAGES;
  uint64_t* lo_empty = split_bar + GM_MAX_STAGES;  // [2]
  uint64_t* tmem_full = lo_empty + 2;              // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_a);
    tma_prefetch_desc(&p.tm_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&split_bar[s], 4);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
      mbar_init(&lo_empty[s], 1);
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

  const int total_items = p.m_tiles * p.n_tiles * p.splits;

  // item -> (m tile, n tile, split); k chunk range of the split
#define GM_DECODE_ITEM(item)                                              \
  int _t = (item);                                                        \
  const int split = _t % p.splits; _t /= p.splits;                        \
  const int n_t = _t % p.n_tiles;                                         \
  const int m_t = _t / p.n_tiles;                                         \
  const int kc0 = split * p.chunks_per_split;                             \
  int kc1 = kc0 + p.chunks_per_split;                                     \
  if (kc1 > p.k_chunks) kc1 = p.k_chunks;                                 \
  const int n_kc = kc1 > kc0 ? kc1 - kc0 : 0;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        GM_DECODE_ITEM(item);
        (void)n_kc;
        for (int kc = kc0; kc < kc1; ++kc) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
          if (p.a_mn) tma_load_3d(sa, &p.tm_a, &full_bar[s], 0, kc * GM_KCHUNK, m_t * 4);
          else tma_load_2d(sa, &p.tm_a, &full_bar[s], kc * GM_KCHUNK, m_t * 128);
          if (p.b_mn) tma_load_3d(sa + p.stage_a_bytes, &p.tm_b, &full_bar[s], 0, kc * GM_KCHUNK, n_t * (p.n_tile >> 5));
          else tma_load_2d(sa + p.stage_a_bytes, &p.tm_b, &full_bar[s], kc * GM_KCHUNK, n_t * p.n_tile);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    int s = 0;
    uint32_t ph = 0;
    int as = 0;
    uint32_t aph = 0;
    uint32_t lo_i = 0;
    const uint32_t a_hi = p.a_mn ? smem_desc_hi(512u, SWZ_128B_ATOM32) : smem_desc_hi(1024u, SWZ_128B);
    const uint32_t b_hi = p.b_mn ? smem_desc_hi(512u, SWZ_128B_ATOM32) : smem_desc_hi(1024u, SWZ_128B);
    const uint32_t a_lbo = p.a_mn ? 4096u : 16u, b_lbo = p.b_mn ? 4096u : 16u;
    const uint32_t a_step = p.a_mn ? 64u : 2u, b_step = p.b_mn ? 64u : 2u;  // K = 8 advance, in 16-byte units
    const uint32_t smem0 = smem_u32(smem), lo0 = smem_u32(lo_ring);
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      GM_DECODE_ITEM(item);
      (void)n_t; (void)m_t;
      mbar_wait(&tmem_empty[as], aph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.n_tile);
      uint32_t accumulate = 0;
      for (int kc = 0; kc < n_kc; ++kc) {
        mbar_wait(p.x3 ? &split_bar[s] : &full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem0 + (uint32_t)s * stage_bytes;
        uint32_t a_lo = smem_desc_lo(sa, a_lbo), b_lo = smem_desc_lo(sa + p.stage_a_bytes, b_lbo);
        if (p.x3) {
          const uint32_t la = lo0 + (lo_i & 1u) * stage_bytes;
          uint32_t al_lo = smem_desc_lo(la, a_lbo), bl_lo = smem_desc_lo(la + p.stage_a_bytes, b_lbo);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              mma_tf32(d_tmem, desc64(al_lo, a_hi), desc64(b_lo, b_hi), p.idesc, accumulate);
              mma_tf32(d_tmem, desc64(a_lo, a_hi), desc64(bl_lo, b_hi), p.idesc, 1u);
              mma_tf32(d_tmem, desc64(a_lo, a_hi), desc64(b_lo, b_hi), p.idesc, 1u);
              accumulate = 1;
              a_lo += a_step; al_lo += a_step; b_lo += b_step; bl_lo += b_step;
            }
            mma_commit(&empty_bar[s]);
            mma_commit(&lo_empty[lo_i & 1u]);
          }
        } else {
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              mma_tf32(d_tmem, desc64(a_lo, a_hi), desc64(b_lo, b_hi), p.idesc, accumulate);
              accumulate = 1;
              a_lo += a_step; b_lo += b_step;
            }
            mma_commit(&empty_bar[s]);
          }
        }
        accumulate = 1;
        __syncwarp();
        ++lo_i;
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) mma_commit(&tmem_full[as]);
      __syncwarp();
      as ^= 1;
      if (as == 0) aph ^= 1u;
    }
  } else if (warp >= 6) {
    const int t = threadIdx.x - 192;
    int s = 0;
    uint32_t ph = 0;
    uint32_t lo_i = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      GM_DECODE_ITEM(item);
      (void)n_t; (void)m_t;
      for (int kc = 0; kc < n_kc; ++kc) {
        mbar_wait(&lo_empty[lo_i & 1u], ((lo_i >> 1) & 1u) ^ 1u);
        mbar_wait(&full_bar[s], ph);
        split_stage_hi_lo(smem + (size_t)s * stage_bytes, lo_ring + (size_t)(lo_i & 1u) * stage_bytes, stage_bytes, t);
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[s]);
        ++lo_i;
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;
    int as = 0;
    uint32_t aph = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      GM_DECODE_ITEM(item);
      const int m = m_t * 128 + q * 32 + lane;
      const bool row_ok = m < p.M;
      float* dst_row = p.out + ((long long)split * p.M + m) * p.N;
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      for (int j0 = 0; j0 < p.n_tile; j0 += 32) {
        uint32_t r[32];
        if (n_kc > 0) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.n_tile + j0), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        const int n0 = n_t * p.n_tile + j0;
        if (row_ok && n0 < p.N) {
          if (p.vec_store && n0 + 32 <= p.N) {
            float4* d4 = reinterpret_cast<float4*>(dst_row + n0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 v = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                     __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
              if (p.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + i);
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
              }
              if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              d4[i] = v;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (n0 + i < p.N) {
                float v = __uint_as_float(r[i]);
                if (p.bias) v += __ldg(p.bias + n0 + i);
                if (p.relu) v = fmaxf(v, 0.f);
                dst_row[n0 + i] = v;
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
#undef GM_DECODE_ITEM

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// out[e] = sum_s ws[s][e] (+ bias[e % N]) (ReLU) — fixed summation order
__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(float* __restrict__ out, const float* __restrict__ ws,
                                                                 int S, long long n, const float* __restrict__ bias, int N,
                                                                 int relu) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += ws[(long long)s * n + e];
    if (bias) acc += bias[e % N];
    if (relu) acc = fmaxf(acc, 0.f);
    out[e] = acc;
  }
}

bool gemm_supported(long long M, long long N, long long K, int trans_a, int trans_b) {
  if (M < 32 || N < 16 || K < 32 || K % 4 != 0) return false;
  if (M >= 2147483647LL || N >= 2147483647LL || K >= 2147483647LL) return false;
  if (trans_a && M % 32 != 0) return false;   // MN-major A: 32-row groups
  if (!trans_a && K % 4 != 0) return false;   // row stride multiple of 16 bytes
  if (!trans_b && N % 32 != 0) return false;  // MN-major B
  if (trans_a && M % 4 != 0) return false;
  return true;
}

// C[M,N] = op(A) @ op(B) (+bias) ; returns 0 done, -1 unsupported (use the SIMT kernel), > 0 error
int gemm_tf32(float* c, const float* a, const float* b, const float* bias, long long M, long long N, long long K,
              int trans_a, int trans_b, int relu, int x3) {
  if (!gemm_supported(M, N, K, trans_a, trans_b)) return -1;
  if ((reinterpret_cast<uintptr_t>(a) & 15u) || (reinterpret_cast<uintptr_t>(b) & 15u)) return -1;
  GemmUP p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.a_mn = trans_a ? 1 : 0;
  p.b_mn = trans_b ? 0 : 1;
  const int n_cap = x3 ? 128 : 256;
  int n_tile = (int)ceil_div(N, 32) * 32;
  if (n_tile > n_cap) n_tile = n_cap;
  p.n_tile = n_tile;
  p.m_tiles = (int)ceil_div(M, 128);
  p.n_tiles = (int)ceil_div(N, n_tile);
  p.k_chunks = (int)ceil_div(K, GM_KCHUNK);
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  const int tiles = p.m_tiles * p.n_tiles;
  int splits = 1;
  if (tiles * 2 <= sms && p.k_chunks >= 8) {
    splits = sms / tiles;
    if (splits > p.k_chunks / 4) splits = p.k_chunks / 4;
    if (splits < 1) splits = 1;
  }
  p.chunks_per_split = (int)ceil_div(p.k_chunks, splits);
  splits = (int)ceil_div(p.k_chunks, p.chunks_per_split);
  p.splits = splits;
  p.x3 = x3 ? 1 : 0;
  p.relu = (splits == 1) ? relu : 0;
  p.bias = (splits == 1) ? bias : nullptr;
  p.idesc = make_idesc_tf32(128, n_tile, p.a_mn, p.b_mn);
  p.stage_a_bytes = 128u * GM_KCHUNK * 4u;
  p.stage_b_bytes = (uint32_t)n_tile * GM_KCHUNK * 4u;
  const uint32_t stage_bytes = p.stage_a_bytes + p.stage_b_bytes;
  int stages = (int)((227u * 1024u - 2048u) / stage_bytes) - (x3 ? 2 : 0);
  if (stages > GM_MAX_STAGES) stages = GM_MAX_STAGES;
  if (stages < 2) return -1;
  p.stages = stages;
  p.tmem_cols = (uint32_t)pow2_ceil_i(2 * n_tile < 32 ? 32 : 2 * n_tile);

  float* ws = nullptr;
  if (splits > 1) {
    if (pool_alloc((void**)&ws, (uint64_t)splits * M * N * sizeof(float))) return 1;
    p.out = ws;
    p.vec_store = (N % 4 == 0) ? 1 : 0;
  } else {
    p.out = c;
    p.vec_store = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(c) & 15u) == 0) &&
                  (!bias || (reinterpret_cast<uintptr_t>(bias) & 15u) == 0);
  }
  int rc = 0;
  if (p.a_mn) {  // A stored [K][M]
    const uint64_t dims[3] = {32u, (uint64_t)K, (uint64_t)(M / 32)};
    const uint64_t strides[2] = {(uint64_t)M * 4u, 128u};
    const uint32_t box[3] = {32u, (uint32_t)GM_KCHUNK, 4u};
    if (encode_tensor_map(&p.tm_a, a, 3, dims, strides, box, SWZ_128B_ATOM32)) rc = -1;
  } else {       // A stored [M][K]
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)K * 4u};
    const uint32_t box[2] = {(uint32_t)GM_KCHUNK, 128u};
    if (encode_tensor_map(&p.tm_a, a, 2, dims, strides, box, SWZ_128B)) rc = -1;
  }
  if (rc == 0) {
    if (p.b_mn) {  // B stored [K][N]
      const uint64_t dims[3] = {32u, (uint64_t)K, (uint64_t)(N / 32)};
      const uint64_t strides[2] = {(uint64_t)N * 4u, 128u};
      const uint32_t box[3] = {32u, (uint32_t)GM_KCHUNK, (uint32_t)(n_tile / 32)};
      if (encode_tensor_map(&p.tm_b, b, 3, dims, strides, box, SWZ_128B_ATOM32)) rc = -1;
    } else {       // B stored [N][K]
      const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
      const uint64_t strides[1] = {(uint64_t)K * 4u};
      const uint32_t box[2] = {(uint32_t)GM_KCHUNK, (uint32_t)n_tile};
      if (encode_tensor_map(&p.tm_b, b, 2, dims, strides, box, SWZ_128B)) rc = -1;
    }
  }
  if (rc == 0) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(227 * 1024));
      if (e != cudaSuccess) rc = set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      configured = (e == cudaSuccess);
    }
  }
  if (rc == 0) {
    const size_t smem_bytes = (size_t)(stages + (x3 ? 2 : 0)) * stage_bytes + 1024 + 512;
    const int items = tiles * splits;
    const int grid = items < sms ? items : sms;
    NG_LAUNCH(umma_gemm_kernel, grid, x3 ? GM_THREADS_X3 : GM_THREADS, smem_bytes, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = set_error("umma_gemm_kernel launch failed: %s", cudaGetErrorString(e));
  }
  if (rc == 0 && splits > 1) {
    const long long n = M * N;
    int g = (int)ceil_div(n, 256);
    if (g > sms * 8) g = sms * 8;
    NG_LAUNCH(gemm_splitk_reduce_kernel, g, 256, 0, c, (const float*)ws, splits, n, bias, (int)N, relu);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = set_error("gemm_splitk_reduce_kernel launch failed: %s", cudaGetErrorString(e));
  }
  if (ws) pool_free(ws);
  return rc;
}

}  // namespace umma
}  // namespace ng
#endif  // NG_EMU
