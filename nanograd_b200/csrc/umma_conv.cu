// Tensor-core (tcgen05, TF32 inputs / FP32 accumulate) implicit-GEMM convolution for sm_100a.
//
// Opt-in path (flag NG_F_TF32) behind the same C-ABI entry points as the FP32 SIMT kernels of
// conv_gemm.cu; stated tolerance rtol 1e-2 (scale-relative) against the reference's NumPy path
// (nanograd/nn/ops_cpu.py:345-400).  The C ABI keeps the reference's layouts (NCHW activations,
// KCRS filters, NCHW results).
//
// Measured on B200 (tools/umma_probe*.cu): a TMA box must start at a 16-byte aligned innermost
// coordinate, so the +-1 pixel tap shifts cannot be expressed on the W axis of an NCHW tensor.
// Each call therefore stages a channels-last (NHWC) copy of the gathered activation tensor
// (one bandwidth-bound transpose), after which every filter tap is a TMA box shifted along the
// non-innermost W/H axes, with zero padding supplied by TMA out-of-bounds fill.
//
// fprop / dgrad ("gather" kernel), stride 1:
//   GEMM view  D[pixel, j] = sum_{tap, kc} A[pixel shifted by tap, kc] * B[tap][j][kc]
//   A  activations NHWC, per (tap, 32-channel chunk) ONE 4-D TMA box {32 c, w, rows, imgs} ->
//      shared memory [128 pixels][32 c] = canonical K-major UMMA operand, 128B swizzle.
//   B  filters, re-laid once per call as [tap][j][kc] (K-major, 128B swizzle), 3-D TMA box.
//   D  128 pixels x N_TILE output channels, FP32 in TMEM, double buffered so the epilogue of tile
//      t (TMEM -> registers -> +bias/ReLU -> coalesced NCHW stores) overlaps the MMAs of tile t+1.
//   Persistent CTAs (one per SM), warp-specialised: warp 0 TMA producer, warp 1 MMA issuer
//   (single elected thread), warps 2-5 epilogue.
#include "umma_common.cuh"

#ifndef NG_EMU
#include <cuda_runtime.h>
#include <stdlib.h>

namespace ng {
namespace umma {

// ------------------------------------------------------------------ driver entry point
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return set_error("cuTensorMapEncodeTiled is not available from the driver (%s)", cudaGetErrorString(e));
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

int encode_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle, int fp16) {
  if (load_encode()) return 1;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle == SWZ_128B) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  else if (swizzle == SWZ_64B) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle == SWZ_32B) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle == SWZ_128B_ATOM32) sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  CUresult r = g_encode(out, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                        bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// ------------------------------------------------------------------ filter re-layout
// dst[(rs*J + j)*KC + kc] = w[j*sj + kc*sc + rs]
__global__ void __launch_bounds__(256) filter_relayout_kernel(float* __restrict__ dst, const float* __restrict__ w,
                                                              int J, int KC, int RS, long long sj, long long sc,
                                                              int round) {
  const long long n = (long long)J * KC * RS;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int kc = (int)(e % KC);
    const long long t = e / KC;
    const int j = (int)(t % J);
    const int rs = (int)(t / J);
    const float v = w[(long long)j * sj + (long long)kc * sc + rs];
    dst[e] = round ? tf32_round(v) : v;
  }
}

// ------------------------------------------------------------------ NCHW -> NHWC staging copy
// out[n][p][c] = in[n][c][p], p = h*W + w.  32 x 32 tiles through shared memory, coalesced both ways.
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(float* __restrict__ out, const float* __restrict__ in, int C,
                                                           int HW, int round) {
  __shared__ float tile[32][33];
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const long long img = (long long)blockIdx.z * C * HW;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, p = p0 + tx;
    tile[ty + 8 * i][tx] = (c < C && p < HW) ? in[img + (long long)c * HW + p] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty + 8 * i, c = c0 + tx;
    if (p < HW && c < C) {
      const float v = tile[tx][ty + 8 * i];
      out[img + (long long)p * C + c] = round ? tf32_round(v) : v;
    }
  }
}

// The same staging copy for dy that also produces the bias gradient: a block walks every pixel tile
// of one (image, 32-channel group), keeps the running channel sums in registers and writes one
// partial per (image, channel); channel_partial_sum_kernel adds the images in a fixed order.
__global__ void __launch_bounds__(256) nchw_to_nhwc_sum_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                               float* __restrict__ partial, int C, int HW, int round) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32;
  const long long img = (long long)blockIdx.y * C * HW;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int p0 = 0; p0 < HW; p0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty + 8 * i, p = p0 + tx;
      const float v = (c < C && p < HW) ? in[img + (long long)c * HW + p] : 0.f;
      tile[ty + 8 * i][tx] = v;
      acc[i] += v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + ty + 8 * i, c = c0 + tx;
      if (p < HW && c < C) {
        const float v = tile[tx][ty + 8 * i];
        out[img + (long long)p * C + c] = round ? tf32_round(v) : v;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int c = c0 + ty + 8 * i;
    if (tx == 0 && c < C) partial[(long long)blockIdx.y * C + c] = v;
  }
}

// out[split][c] = sum over the rows n of the split of partial[n][c]; block = 32 channels x 32 row lanes,
// grid (C / 32, splits)
__global__ void __launch_bounds__(1024) channel_partial_sum_kernel(float* __restrict__ out, const float* __restrict__ partial,
                                                                   long long N, int C, long long rows_per_split) {
  __shared__ float red[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long long n0 = (long long)blockIdx.y * rows_per_split;
  long long n1 = n0 + rows_per_split;
  if (n1 > N) n1 = N;
  float acc = 0.f;
  if (c < C)
    for (long long n = n0 + ty; n < n1; n += 32) acc += partial[n * C + c];
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][tx];
    out[(long long)blockIdx.y * C + c] = t;
  }
}

// sums[c] = sum_n partial[n][c] in a fixed order; many rows are folded in two levels so that the first one fills
// the machine
int channel_partial_sum(float* sums, const float* partial, long long N, int C) {
  const unsigned cb = (unsigned)ceil_div(C, 32);
  if (N <= 2048) {
    NG_LAUNCH(channel_partial_sum_kernel, dim3(cb, 1), 1024, 0, sums, partial, N, C, N);
    NG_CHECK_LAUNCH();
    return 0;
  }
  const long long per = 512;
  const long long splits = ceil_div(N, per);
  float* mid = nullptr;
  if (pool_alloc((void**)&mid, (uint64_t)splits * C * sizeof(float))) return 1;
  NG_LAUNCH(channel_partial_sum_kernel, dim3(cb, (unsigned)splits), 1024, 0, mid, partial, N, C, per);
  NG_CHECK_LAUNCH();
  NG_LAUNCH(channel_partial_sum_kernel, dim3(cb, 1), 1024, 0, sums, (const float*)mid, splits, C, splits);
  NG_CHECK_LAUNCH();
  pool_free(mid);
  return 0;
}

int to_nhwc_sum(float* out, const float* in, float* sums, int N, int C, int HW, int round) {
  float* partial = nullptr;
  if (pool_alloc((void**)&partial, (uint64_t)N * C * sizeof(float))) return 1;
  dim3 grid((unsigned)ceil_div(C, 32), (unsigned)N);
  NG_LAUNCH(nchw_to_nhwc_sum_kernel, grid, 256, 0, out, in, partial, C, HW, round);
  NG_CHECK_LAUNCH();
  if (channel_partial_sum(sums, partial, N, C)) return 1;
  pool_free(partial);
  return 0;
}

int to_nhwc(float* out, const float* in, int N, int C, int HW, int round) {
  dim3 grid((unsigned)ceil_div(HW, 32), (unsigned)ceil_div(C, 32), (unsigned)N);
  NG_LAUNCH(nchw_to_nhwc_kernel, grid, 256, 0, out, in, C, HW, round);
  NG_CHECK_LAUNCH();
  return 0;
}

// Ablation switches (NG_UMMA_EXP=<bits>) exist only in builds with -DNG_UMMA_ABLATE:
// even an always-false runtime test on the per-stage paths costs measurable time in these kernels.
#ifdef NG_UMMA_ABLATE
#define NG_EXP(bits) ((p.exp_flags & (bits)) != 0)
#define NG_EXP_FLAGS (p.exp_flags)
#else
#define NG_EXP(bits) false
#define NG_EXP_FLAGS 0
#endif

// ------------------------------------------------------------------ gather kernel (fprop / dgrad)
constexpr int G_THREADS = 224;       // TF32 mode: producer, MMA, 4 epilogue warps, second producer
constexpr int G_THREADS_X3 = 480;    // 3xTF32 mode: + 2 groups of 4 operand-split warps before the second producer
constexpr int G_CCHUNK = 32;        // channels per pipeline stage (4 MMAs of K = 8)
constexpr int G_MAX_STAGES = 8;
// Ring depths of the 3xTF32 hand-off.  The chain MMA commit -> slot free -> split warps -> MMA issue is
// longer than one stage of MMAs, so two lo buffers / three A buffers left the tensor pipe idle between
// stages: four of each (all 512 TMEM columns at n_tile = 128) measured 6-7 % faster, six lo buffers slower
// again (they cost a TMA stage).
constexpr unsigned G_LO_RING = 4;   // 3xTF32: B lo-part buffers in shared memory
constexpr int G_A_RING = 4;         // 3xTF32: A (hi, lo) buffers in TMEM, 64 columns each

struct GatherP {
  CUtensorMap tm_a;  // activations NHWC, dims (C, W, H, N)
  CUtensorMap tm_b;  // filters [tap][j][kc], dims (KC, J, RS)
  float* out;
  const float* bias;
  int bias_vec;   // bias is 16-byte aligned: the epilogue reads it four values at a time
  int n_img, J, KC, RS, OH, OW;
  int w_tile, rows_tile, imgs_tile;   // w_tile * rows_tile * imgs_tile == 128
  int tiles_w, tiles_h, tiles_img;
  int n_tile, n_tiles;                // MMA N and the number of output-channel tiles
  int c_chunks;
  uint32_t idesc;
  uint32_t stage_a_bytes, stage_b_bytes;
  int stages;
  uint32_t tmem_cols;
  int relu;
  int x3;         // 3xTF32 (FP32-accurate) mode: stages hold hi and lo halves
  int exp_flags;  // experiments (NG_UMMA_EXP): 1 = skip the split arithmetic, 2 = skip the MMAs, 4 = skip the epilogue stores
  uint32_t* dbg;  // host-mapped progress words (NG_UMMA_DEBUG=1), else null
  short tap_dx[64], tap_dy[64];
};

#define NG_DBG(i, v)                                   \
  do {                                                 \
    if (p.dbg) { ((volatile uint32_t*)p.dbg)[i] = (v); __threadfence_system(); } \
  } while (0)

template <bool X3>
__global__ void __launch_bounds__(X3 ? G_THREADS_X3 : G_THREADS, 1) umma_conv_gather_kernel(const __grid_constant__ GatherP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // Shared memory: `stages` operand sets (A tile + B tile) filled by TMA, then — 3xTF32 mode only — a
  // ring of G_LO_RING "lo" sets.  A lo set lives only from the split to the end of its MMAs, an operand set
  // from the TMA issue on, so most of the capacity stays available for loads in flight.
  // 3xTF32 mode: the split warps move the A tile into TMEM (hi and lo, G_A_RING buffers of 64 columns)
  // and write only B's lo part back to shared memory (ring of G_LO_RING), so the MMAs read just B from shared
  // memory: 32 KB of shared-memory traffic per K step of a 128x128 tile instead of 48 KB.
  const uint32_t stage_bytes = p.stage_a_bytes + p.stage_b_bytes;
  uint8_t* lo_ring = smem + (size_t)p.stages * stage_bytes;   // x3: 2 x stage_b_bytes (B lo)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(lo_ring + (X3 ? G_LO_RING * p.stage_b_bytes : 0u));
  uint64_t* empty_bar = full_bar + G_MAX_STAGES;
  uint64_t* split_bar = empty_bar + G_MAX_STAGES;
  uint64_t* lo_empty = split_bar + G_MAX_STAGES;   // [G_LO_RING]
  uint64_t* a_empty = lo_empty + G_LO_RING;        // [G_A_RING]
  uint64_t* tmem_full = a_empty + G_A_RING;
  uint64_t* tmem_empty = tmem_full + 2;
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
    for (int s = 0; s < G_A_RING; ++s) mbar_init(&a_empty[s], 1);
    for (int s = 2; s < (int)G_LO_RING; ++s) mbar_init(&lo_empty[s], 1);
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
  const uint32_t tmem_a0 = tmem_base + (uint32_t)(2 * p.n_tile);   // x3: A ring after the two accumulators
  if (threadIdx.x == 0 && blockIdx.x == 0) { NG_DBG(4, tmem_base); NG_DBG(5, 0x1111u); }

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_img;
  const int total_tiles = m_tiles * p.n_tiles;
  const int k_iters = p.RS * p.c_chunks;

  const int producer2 = X3 ? 14 : 6;   // two TMA-issuing warps alternate stages (see W_THREADS below)
  if (warp == 0 || warp == producer2) {
    if (lane == 0) {
      const uint32_t pid = warp == 0 ? 0u : 1u;
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
              tma_load_4d(sa, &p.tm_a, &full_bar[s], cc * G_CCHUNK, cx, cy, img0);
              tma_load_3d(sa + p.stage_a_bytes, &p.tm_b, &full_bar[s], cc * G_CCHUNK, n_t * p.n_tile, rs);
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
    uint32_t lo_i = 0;  // running stage counter: lo set = lo_i & 1
    const uint32_t dhi = smem_desc_hi(1024u, SWZ_128B);
    const uint32_t smem0 = smem_u32(smem), lo0 = smem_u32(lo_ring);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[as], aph ^ 1u, p.dbg, 0x200u + as);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.n_tile);
      uint32_t accumulate = 0;
      int cc = 0;
      for (int it = 0; it < k_iters; ++it) {
        int ksteps = (p.KC - cc * G_CCHUNK + 7) >> 3;
        if (ksteps > 4) ksteps = 4;
        if (NG_EXP(2)) ksteps = 0;
        mbar_wait(X3 ? &split_bar[s] : &full_bar[s], ph, p.dbg, 0x300u + s);
        tc_fence_after();
        const uint32_t sa = smem0 + (uint32_t)s * stage_bytes;
        uint32_t a_lo = smem_desc_lo(sa, 16u), b_lo = smem_desc_lo(sa + p.stage_a_bytes, 16u);
        if (X3) {
          uint32_t bl_lo = smem_desc_lo(lo0 + (lo_i % G_LO_RING) * p.stage_b_bytes, 16u);
          const uint32_t aj = lo_i % G_A_RING;
          uint32_t a_t = tmem_a0 + aj * 64u;   // hi: columns [0,32), lo: [32,64) of the A buffer
          if (elect_one()) {
            if (ksteps == 4) {   // the common case, fully unrolled: this thread's time per stage is the kernel's critical path
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                mma_tf32_ts(d_tmem, a_t + 32u + 8u * k, desc64(b_lo + 2u * k, dhi), p.idesc, k == 0 ? accumulate : 1u);
                mma_tf32_ts(d_tmem, a_t + 8u * k, desc64(bl_lo + 2u * k, dhi), p.idesc, 1u);
                mma_tf32_ts(d_tmem, a_t + 8u * k, desc64(b_lo + 2u * k, dhi), p.idesc, 1u);
              }
            } else {
              for (int k = 0; k < ksteps; ++k) {
                mma_tf32_ts(d_tmem, a_t + 32u, desc64(b_lo, dhi), p.idesc, accumulate);
                mma_tf32_ts(d_tmem, a_t, desc64(bl_lo, dhi), p.idesc, 1u);
                mma_tf32_ts(d_tmem, a_t, desc64(b_lo, dhi), p.idesc, 1u);
                accumulate = 1;
                a_t += 8u; b_lo += 2; bl_lo += 2;   // 8 TMEM columns / 32 bytes along K
              }
            }
            mma_commit(&empty_bar[s]);
            mma_commit(&lo_empty[lo_i % G_LO_RING]);
            mma_commit(&a_empty[aj]);
          }
        } else {
          if (elect_one()) {
            if (ksteps == 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma_tf32(d_tmem, desc64(a_lo + 2u * k, dhi), desc64(b_lo + 2u * k, dhi), p.idesc, k == 0 ? accumulate : 1u);
            } else {
              for (int k = 0; k < ksteps; ++k) {
                mma_tf32(d_tmem, desc64(a_lo, dhi), desc64(b_lo, dhi), p.idesc, accumulate);
                accumulate = 1;
                a_lo += 2; b_lo += 2;
              }
            }
            mma_commit(&empty_bar[s]);
          }
        }
        accumulate = ksteps > 0 ? 1u : accumulate;
        __syncwarp();
        ++lo_i;
        if (++cc == p.c_chunks) cc = 0;
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) mma_commit(&tmem_full[as]);
      __syncwarp();
      as ^= 1;
      if (as == 0) aph ^= 1u;
    }
  } else if (warp >= 6) {
    // 3xTF32 operand split (these warps exist only in x3 launches)
    // Two groups of four split warps take alternate stages (group g: stage counters g, g+2, ...), so
    // the barrier wake-up / LDS / TMEM-store latency of one stage overlaps the other group's work.
    // Thread <-> pixel row m of the A tile <-> TMEM lane m (a warp may touch lane quarter warp % 4).
    const int grp = (warp - 6) >> 2;
    const int t = (threadIdx.x - 192) & 127;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t smem0 = smem_u32(smem);
    int my_tiles = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) ++my_tiles;
    const uint32_t total_iters = (uint32_t)my_tiles * (uint32_t)k_iters;
    int s = grp % p.stages;   // stage index and phase are stepped: no division on the per-stage path
    uint32_t ph = (uint32_t)(grp / p.stages) & 1u;
    for (uint32_t lo_i = (uint32_t)grp; lo_i < total_iters; lo_i += 2u) {
      const uint32_t aj = lo_i % G_A_RING;
      mbar_wait(&a_empty[aj], ((lo_i / G_A_RING) & 1u) ^ 1u, p.dbg, 0x700u + aj);
      mbar_wait(&lo_empty[lo_i % G_LO_RING], ((lo_i / G_LO_RING) & 1u) ^ 1u, p.dbg, 0x600u + (lo_i % G_LO_RING));
      mbar_wait(&full_bar[s], ph, p.dbg, 0x500u + s);
      tc_fence_after();
      if (!NG_EXP(1)) {
        // A: own 128-byte row (32 channels), 16-byte chunks un-swizzled (chunk ^ (row & 7))
        const uint32_t row = smem0 + (uint32_t)s * stage_bytes + (uint32_t)m * 128u;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = lds128(row + (uint32_t)((c ^ (m & 7)) << 4));
          hi[4 * c + 0] = __float_as_uint(v.x); hi[4 * c + 1] = __float_as_uint(v.y);
          hi[4 * c + 2] = __float_as_uint(v.z); hi[4 * c + 3] = __float_as_uint(v.w);
          lo[4 * c + 0] = __float_as_uint(tf32_lo(v.x)); lo[4 * c + 1] = __float_as_uint(tf32_lo(v.y));
          lo[4 * c + 2] = __float_as_uint(tf32_lo(v.z)); lo[4 * c + 3] = __float_as_uint(tf32_lo(v.w));
        }
        const uint32_t a_t = tmem_a0 + aj * 64u + lane_addr;
        tmem_st_32x32(a_t, hi);
        tmem_st_32x32(a_t + 32u, lo);
        // B: lo part to the shared-memory ring
        split_stage_hi_lo(smem + (size_t)s * stage_bytes + p.stage_a_bytes,
                          lo_ring + (size_t)(lo_i % G_LO_RING) * p.stage_b_bytes, p.stage_b_bytes, t, NG_EXP_FLAGS);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&split_bar[s]);
      s += 2;
      if (s >= p.stages) { s -= p.stages; ph ^= 1u; }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;
    const int g = m / p.w_tile, wcol = m - g * p.w_tile;
    const int gi = g / p.rows_tile, gr = g - gi * p.rows_tile;
    const long long opix = (long long)p.OH * p.OW;
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
      if (blockIdx.x == 0 && threadIdx.x == 64) NG_DBG(10, (uint32_t)(tile + 1));
      for (int j0 = 0; j0 < p.n_tile; j0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.n_tile + j0);
        tmem_ld_32x32(taddr, r);
        tmem_ld_wait();
        const int jbase = n_t * p.n_tile + j0;
        if (valid && !NG_EXP(4)) {
          // These warps share issue slots with the split warps and the MMA thread, so the store loop is kept
          // lean: one running pointer, no per-element bounds test on full chunks, bias as four-wide loads.
          float* dst = out_pix + (long long)jbase * opix;
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

static inline int pow2_ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

// Shapes the tensor-core kernels take (everything else runs on the FP32 SIMT kernels).
bool gather_supported(int KC, int J, int RS) { return KC % 4 == 0 && KC >= 8 && RS <= 64 && J >= 8; }
bool wgrad_supported(int C, int K, int RS) { return C % 32 == 0 && K % 32 == 0 && RS <= 64; }

// Launches the gather kernel.  `act` is the gathered activation tensor [n_img, KC, AH, AW] (NCHW);
// `filt` the KCRS filters addressed as filt[j*sj + kc*sc + rs]; the output is [n_img, J, OH, OW].
// Returns 0 on success, -1 when the shape is not supported by this path (caller falls back to
// the SIMT kernel), > 0 on error.
static int launch_gather(float* out, const float* act, const float* filt, const float* bias, int n_img, int KC, int AH,
                         int AW, int J, int R, int S, long long sj, long long sc, int OH, int OW, const int* tap_dy,
                         const int* tap_dx, int relu, int x3, int act_is_nhwc) {
  const int RS = R * S;
  if (!gather_supported(KC, J, RS)) return -1;
  GatherP p;
  memset(&p, 0, sizeof(p));
  p.out = out;
  p.bias = bias;
  p.bias_vec = (bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & 15u) == 0) ? 1 : 0;
  p.n_img = n_img; p.J = J; p.KC = KC; p.RS = RS; p.OH = OH; p.OW = OW;
  int w_tile = pow2_ceil(OW);
  if (w_tile > 32) w_tile = 32;
  int rows = 128 / w_tile;
  const int oh2 = pow2_ceil(OH);
  if (rows > oh2) rows = oh2;
  p.w_tile = w_tile;
  p.rows_tile = rows;
  p.imgs_tile = 128 / (w_tile * rows);
  p.tiles_w = (int)ceil_div(OW, p.w_tile);
  p.tiles_h = (int)ceil_div(OH, rows);
  p.tiles_img = (int)ceil_div(n_img, p.imgs_tile);
  int n_tile = (int)ceil_div(J, 32) * 32;
  const int n_cap = x3 ? 128 : 256;  // x3 stages hold hi and lo halves: keep >= 3 stages in 227 KB
  if (n_tile > n_cap) n_tile = n_cap;
  p.n_tile = n_tile;
  p.x3 = x3 ? 1 : 0;
  p.n_tiles = (int)ceil_div(J, n_tile);
  p.c_chunks = (int)ceil_div(KC, G_CCHUNK);
  p.idesc = make_idesc_tf32(128, n_tile, /*a K-major*/ 0, /*b K-major*/ 0);
  p.stage_a_bytes = 128u * G_CCHUNK * 4u;
  p.stage_b_bytes = (uint32_t)n_tile * 128u;
  const uint32_t stage_bytes = p.stage_a_bytes + p.stage_b_bytes;
  const uint32_t budget = 227u * 1024u - 2048u;
  // x3: two B-lo sets share the capacity
  int stages = (int)((budget - (x3 ? G_LO_RING * p.stage_b_bytes : 0u)) / stage_bytes);
  if (stages > G_MAX_STAGES) stages = G_MAX_STAGES;
  if (stages < 2) return -1;
  p.stages = stages;
  // TMEM: two accumulators (+ the A ring in x3 mode: 2*128 + 4*64 = 512 columns)
  p.tmem_cols = (uint32_t)pow2_ceil(2 * n_tile + (x3 ? G_A_RING * 64 : 0) < 32 ? 32 : 2 * n_tile + (x3 ? G_A_RING * 64 : 0));
  p.relu = relu;
  { const char* e = getenv("NG_UMMA_EXP"); p.exp_flags = e ? atoi(e) : 0; }
  for (int t = 0; t < RS; ++t) { p.tap_dx[t] = (short)tap_dx[t]; p.tap_dy[t] = (short)tap_dy[t]; }

  // staging: activations -> NHWC, filters -> [tap][j][kc]
  float* act_t = nullptr;
  float* wt = nullptr;
  if (!act_is_nhwc && pool_alloc((void**)&act_t, (uint64_t)n_img * KC * AH * AW * sizeof(float))) return 1;
  if (pool_alloc((void**)&wt, (uint64_t)RS * J * KC * sizeof(float))) { pool_free(act_t); return 1; }
  int rc = 0;
  const float* act_nhwc = act_is_nhwc ? act : act_t;
  if (!act_is_nhwc) rc = to_nhwc(act_t, act, n_img, KC, AH * AW, /*round=*/x3 ? 0 : 1);
  if (rc == 0) {
    const long long n = (long long)RS * J * KC;
    int g = (int)ceil_div(n, 256);
    if (g > 148 * 8) g = 148 * 8;
    NG_LAUNCH(filter_relayout_kernel, g, 256, 0, wt, filt, J, KC, RS, sj, sc, x3 ? 0 : 1);
    const uint64_t dims[4] = {(uint64_t)KC, (uint64_t)AW, (uint64_t)AH, (uint64_t)n_img};
    const uint64_t strides[3] = {(uint64_t)KC * 4u, (uint64_t)AW * KC * 4u, (uint64_t)AH * AW * KC * 4u};
    const uint32_t box[4] = {(uint32_t)G_CCHUNK, (uint32_t)p.w_tile, (uint32_t)p.rows_tile, (uint32_t)p.imgs_tile};
    if (encode_tensor_map(&p.tm_a, act_nhwc, 4, dims, strides, box, SWZ_128B)) rc = -1;
  }
  if (rc == 0) {
    const uint64_t dims[3] = {(uint64_t)KC, (uint64_t)J, (uint64_t)RS};
    const uint64_t strides[2] = {(uint64_t)KC * 4u, (uint64_t)J * KC * 4u};
    const uint32_t box[3] = {(uint32_t)G_CCHUNK, (uint32_t)n_tile, 1u};
    if (encode_tensor_map(&p.tm_b, wt, 3, dims, strides, box, SWZ_128B)) rc = -1;
  }
  if (rc == 0) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(umma_conv_gather_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(227 * 1024));
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(umma_conv_gather_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(227 * 1024));
      if (e != cudaSuccess) rc = set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      configured = (e == cudaSuccess);
    }
  }
  if (rc == 0) {
    const size_t smem_bytes = (size_t)stages * stage_bytes + (x3 ? G_LO_RING * p.stage_b_bytes : 0u) + 1024 + 512;
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_img * p.n_tiles;
    const int sms = g_sm_count > 0 ? g_sm_count : 148;
    const int grid = total_tiles < sms ? total_tiles : sms;
    static uint32_t* dbg_host = nullptr;
    const bool debug = getenv("NG_UMMA_DEBUG") != nullptr;
    if (debug) {
      if (!dbg_host) cudaHostAlloc((void**)&dbg_host, 64 * sizeof(uint32_t), cudaHostAllocMapped);
      memset(dbg_host, 0, 64 * sizeof(uint32_t));
      p.dbg = dbg_host;
      fprintf(stderr, "[umma] gather: tiles %d grid %d n_tile %d stages %d w_tile %d rows %d imgs %d idesc %08x tmem %u\n",
              total_tiles, grid, p.n_tile, p.stages, p.w_tile, p.rows_tile, p.imgs_tile, p.idesc, p.tmem_cols);
    }
    if (x3) NG_LAUNCH(umma_conv_gather_kernel<true>, grid, G_THREADS_X3, smem_bytes, p);
    else NG_LAUNCH(umma_conv_gather_kernel<false>, grid, G_THREADS, smem_bytes, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = set_error("umma_conv_gather_kernel launch failed: %s", cudaGetErrorString(e));
    if (debug) {
      cudaError_t e2 = cudaStreamSynchronize(g_stream);
      fprintf(stderr, "[umma] sync: %s ; dbg:", cudaGetErrorString(e2));
      for (int i = 0; i < 12; ++i) fprintf(stderr, " %08x", dbg_host[i]);
      fprintf(stderr, "\n");
    }
  }
  pool_free(wt);
  pool_free(act_t);
  return rc;
}

int conv_fprop_tf32(float* y, const float* x, const float* w, const float* bias, int N, int C, int H, int W, int K,
                    int R, int S, int stride, int pad_t, int pad_l, int OH, int OW, int relu, int x3, int x_is_nhwc) {
  if (stride != 1) return -1;
  int dy[64], dx[64];
  if (R * S > 64) return -1;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) { dy[r * S + s] = r - pad_t; dx[r * S + s] = s - pad_l; }
  return launch_gather(y, x, w, bias, N, C, H, W, K, R, S, (long long)C * R * S, (long long)R * S, OH, OW, dy, dx,
                       relu, x3, x_is_nhwc);
}

int conv_dgrad_tf32(float* dx_out, const float* dyv, const float* w, int N, int C, int H, int W, int K, int R, int S,
                    int stride, int pad_t, int pad_l, int OH, int OW, int x3, int dy_is_nhwc) {
  if (stride != 1) return -1;
  int dy[64], dx[64];
  if (R * S > 64) return -1;
  // dx[n,c,y,x] = sum_{k,r,s} dy[n,k,y+pt-r,x+pl-s] * w[k,c,r,s]
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) { dy[r * S + s] = pad_t - r; dx[r * S + s] = pad_l - s; }
  return launch_gather(dx_out, dyv, w, nullptr, N, K, OH, OW, C, R, S, (long long)R * S, (long long)C * R * S, H, W,
                       dy, dx, 0, x3, dy_is_nhwc);
}


// ------------------------------------------------------------------ wgrad kernel
//   per filter tap t = (r, s):   D_t[c][k] = sum_{n,oh,ow} x[n, oh+r-pt, ow+s-pl, c] * dy[n, oh, ow, k]
// Both operands come from the NHWC staging copies, so both are MN-major (channel-contiguous) UMMA
// operands: shared memory [pixel][32 channels] rows of 128 B in the 128B/32B-atom swizzle, the
// only layout the tensor core accepts for MN-major 32-bit data.  The reduction runs over pixels
// (8 per MMA, 32 per pipeline stage).
//   M = 128 rows = (128 / Cm) taps x Cm input channels  (Cm = min(C, 128)): narrow layers put
//       several taps of the same channel tile side by side so the MMA is always 128 rows tall;
//   N = up to 256 output channels;  one 5-D TMA box per tap for x, one for dy, per stage.
// CTA = (tap group, c tile, k tile, split of the pixel range); partial sums go to a workspace
// [split][tap][c][k] and a second pass reduces the splits in a fixed order (deterministic) and
// writes KCRS.
// One thread needs ~200 cycles to issue a TMA box (plus ~150 for the barrier handshake) whatever the
// box size (tools/tma_bw.cu), which caps a single producer at 8-20 B/clk/SM for the 8-16 KB boxes of
// this kernel; issue scales linearly with the number of producer WARPS, so two warps alternate chunks.
constexpr int W_THREADS = 224;      // producer, MMA, 4 epilogue warps, second producer
constexpr int W_THREADS_X3 = 480;   // ... + 2 groups of 4 operand-split warps before the second producer
constexpr unsigned W_RING = 4;      // X3: x (hi, lo) buffers in TMEM (64 columns each) and dy-lo buffers in shared memory
constexpr int W_A_RING_MAX = 4;
constexpr int W_LO_RING_MAX = 4;
constexpr int W_MAX_STAGES = 6;

struct WgradUP {
  CUtensorMap tm_x;   // x NHWC as (32, W, H, C/32, N)
  CUtensorMap tm_dy;  // dy NHWC as (32, OW, OH, K/32, N)
  float* ws;          // [splits][RS][C][K]
  int C, K, RS, S;
  int Cm, taps_per_cta, tap_groups, c_tiles, n_tile, n_tiles;
  int w_c, rows_c, chunks_w, chunks_h, total_chunks, chunks_per_split, splits;
  int pad_t, pad_l;
  uint32_t idesc, stage_a_bytes, stage_b_bytes, tmem_cols;
  int stages;
  int x3;
  int exp_flags;         // experiments (NG_UMMA_EXP): 1 skip x split loads, 2 skip MMAs, 4 skip TMEM stores, 16 skip dy split, 32 skip promotion
  uint32_t* dbg;
};

// X3 = false: TF32 mode, one TMEM accumulator for the whole pixel range of the CTA.
// X3 = true : 3xTF32 mode.  The tensor core's FP32 accumulation is not round-to-nearest, so a long
//   accumulation chain drifts (measured 1.3e-4 relative after 17k pixels); the reduction is
//   therefore cut into groups of W_PROMOTE chunks that ping-pong between two TMEM accumulators,
//   and the epilogue warps add each finished group into FP32 registers (round-to-nearest) while
//   the next group is being multiplied.
constexpr int W_PROMOTE = 8;       // chunks (of 32 pixels) per TMEM accumulation group in X3 mode
constexpr int W_X3_MAX_N = 128;    // output channels per CTA in X3 mode (register accumulator row)

template <bool X3>
__global__ void __launch_bounds__(X3 ? W_THREADS_X3 : W_THREADS, 1)
umma_conv_wgrad_kernel(const __grid_constant__ WgradUP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = p.stage_a_bytes + p.stage_b_bytes;
  // X3: the x operand (hi, lo) goes to TMEM (ring of W_RING buffers of 64 columns, written by the
  // split warps), only dy's lo part goes back to shared memory (ring of W_RING)
  uint8_t* lo_ring = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(lo_ring + (X3 ? W_RING * p.stage_b_bytes : 0u));
  uint64_t* empty_bar = full_bar + W_MAX_STAGES;
  uint64_t* split_bar = empty_bar + W_MAX_STAGES;
  uint64_t* lo_empty = split_bar + W_MAX_STAGES;   // [W_LO_RING_MAX]
  uint64_t* a_empty = lo_empty + W_LO_RING_MAX;    // [W_A_RING_MAX]
  uint64_t* acc_full = a_empty + W_A_RING_MAX;     // [2]
  uint64_t* acc_empty = acc_full + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_x);
    tma_prefetch_desc(&p.tm_dy);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&split_bar[s], 4);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);
    }
    for (int b = 0; b < W_LO_RING_MAX; ++b) mbar_init(&lo_empty[b], 1);
    for (int b = 0; b < W_A_RING_MAX; ++b) mbar_init(&a_empty[b], 1);
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
  const uint32_t tmem_a0 = tmem_base + (uint32_t)(2 * p.n_tile);   // X3: A ring after the two accumulators

  // work item
  int item = blockIdx.x;
  const int split = item % p.splits; item /= p.splits;
  const int n_t = item % p.n_tiles; item /= p.n_tiles;
  const int c_t = item % p.c_tiles;
  const int tg = item / p.c_tiles;
  const int chunk_begin = split * p.chunks_per_split;
  int chunk_end = chunk_begin + p.chunks_per_split;
  if (chunk_end > p.total_chunks) chunk_end = p.total_chunks;
  const int n_chunks = chunk_end > chunk_begin ? chunk_end - chunk_begin : 0;
  const int n_groups = X3 ? (n_chunks + W_PROMOTE - 1) / W_PROMOTE : (n_chunks > 0 ? 1 : 0);
  const int tap0 = tg * p.taps_per_cta;
  int n_taps = p.RS - tap0;
  if (n_taps > p.taps_per_cta) n_taps = p.taps_per_cta;
  const int groups_per_tap = p.Cm >> 5;
  const uint32_t tap_bytes = (uint32_t)groups_per_tap * 4096u;

  constexpr int kProducer2 = X3 ? 14 : 6;
  if (warp == 0 || warp == kProducer2) {
    if (lane == 0) {
      // producer `pid` takes chunks pid, pid + 2, ...; the chunk position is stepped, not divided
      const int pid = warp == 0 ? 0 : 1;
      int s = pid % p.stages;
      uint32_t ph = (uint32_t)(pid / p.stages) & 1u;
      int t = chunk_begin + pid;
      int cw = t % p.chunks_w; t /= p.chunks_w;
      int chh = t % p.chunks_h;
      int img = t / p.chunks_h;
      int tdx[4], tdy[4];
#pragma unroll
      for (int tl = 0; tl < 4; ++tl) {
        const int tap = tap0 + tl;
        tdy[tl] = tap / p.S - p.pad_t;
        tdx[tl] = tap % p.S - p.pad_l;
      }
      const uint32_t tx_bytes = (uint32_t)n_taps * tap_bytes + p.stage_b_bytes;
      const int cg0 = c_t * groups_per_tap, ng0 = n_t * (p.n_tile >> 5);
      for (int i = pid; i < n_chunks; i += 2) {
        const int ow0 = cw * p.w_c, oh0 = chh * p.rows_c;
        mbar_wait(&empty_bar[s], ph ^ 1u, p.dbg, 0x100u + s);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
#pragma unroll
        for (int tl = 0; tl < 4; ++tl)
          if (tl < n_taps)
            tma_load_5d(sa + (size_t)tl * tap_bytes, &p.tm_x, &full_bar[s], 0, ow0 + tdx[tl], oh0 + tdy[tl], cg0, img);
        tma_load_5d(sa + p.stage_a_bytes, &p.tm_dy, &full_bar[s], 0, ow0, oh0, ng0, img);
        s += 2;
        if (s >= p.stages) { s -= p.stages; ph ^= 1u; }
        cw += 2;
        while (cw >= p.chunks_w) { cw -= p.chunks_w; ++chh; }
        while (chh >= p.chunks_h) { chh -= p.chunks_h; ++img; }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform control flow, one elected lane issues (see elect_one()).
    int s = 0;
    uint32_t ph = 0;
    int ch = 0;
    const uint32_t dhi = smem_desc_hi(512u, SWZ_128B_ATOM32);
    const uint32_t smem0 = smem_u32(smem), lo0 = smem_u32(lo_ring);
    for (int g = 0; g < n_groups; ++g) {
      const int b = X3 ? (g & 1) : 0;
      const uint32_t d_tmem = tmem_base + (uint32_t)(b * p.n_tile);
      if (X3) {
        mbar_wait(&acc_empty[b], (uint32_t)(((g >> 1) & 1) ^ 1), p.dbg, 0x200u + b);
        tc_fence_after();
      }
      const int g_end = X3 ? (ch + W_PROMOTE < n_chunks ? ch + W_PROMOTE : n_chunks) : n_chunks;
      uint32_t accumulate = 0;
      for (; ch < g_end; ++ch) {
        mbar_wait(X3 ? &split_bar[s] : &full_bar[s], ph, p.dbg, 0x300u + s);
        tc_fence_after();
        const uint32_t sa = smem0 + (uint32_t)s * stage_bytes;
        uint32_t a_lo = smem_desc_lo(sa, 4096u), b_lo = smem_desc_lo(sa + p.stage_a_bytes, 4096u);
        if (X3) {
          const uint32_t lj = (uint32_t)ch % W_RING;
          uint32_t bl_lo = smem_desc_lo(lo0 + lj * p.stage_b_bytes, 4096u);
          const uint32_t aj = lj;
          uint32_t a_t = tmem_a0 + aj * 64u;   // hi: columns [0,32), lo: [32,64)
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (NG_EXP(2)) break;
              mma_tf32_ts(d_tmem, a_t + 32u, desc64(b_lo, dhi), p.idesc, accumulate);
              mma_tf32_ts(d_tmem, a_t, desc64(bl_lo, dhi), p.idesc, 1u);
              mma_tf32_ts(d_tmem, a_t, desc64(b_lo, dhi), p.idesc, 1u);
              accumulate = 1;
              a_t += 8u; b_lo += 64; bl_lo += 64;   // 8 pixels: 8 TMEM columns / 1024 bytes
            }
            mma_commit(&empty_bar[s]);
            mma_commit(&a_empty[aj]);
          }
        } else {
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              mma_tf32(d_tmem, desc64(a_lo, dhi), desc64(b_lo, dhi), p.idesc, accumulate);
              accumulate = 1;
              a_lo += 64; b_lo += 64;
            }
            mma_commit(&empty_bar[s]);
          }
        }
        accumulate = 1;
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) mma_commit(&acc_full[b]);
      __syncwarp();
    }
  } else if (warp >= 6) {
    // 3xTF32 operand split (X3 launches only).  A last tap group with fewer taps leaves part of
    // the A region unwritten by TMA: splitting that garbage is harmless (those D rows are never stored).
    if (X3) {
      // Two groups of four split warps take alternate chunks.  Thread <-> row m = (tap, channel) of the
      // x operand <-> TMEM lane m; its 32 pixel values come from one shared-memory column of the
      // [pixel][32 channels] group (32-byte chunks swizzled by pixel & 3).
      const int grp = (warp - 6) >> 2;
      const int t = (threadIdx.x - 192) & 127;
      const int q = warp & 3;
      const int m = q * 32 + lane;
      const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
      const uint32_t smem0 = smem_u32(smem);
      const uint32_t col_base = (uint32_t)(m >> 5) * 4096u + (uint32_t)(m & 7) * 4u;  // group, position in 32B chunk
      const uint32_t chunk32 = (uint32_t)((m & 31) >> 3);
      // stage index and phase are stepped (no division on the per-chunk path); the rings are powers of two
      int s = grp % p.stages;
      uint32_t ph = (uint32_t)(grp / p.stages) & 1u;
      for (int ch = grp; ch < n_chunks; ch += 2) {
        const uint32_t aj = (uint32_t)ch % W_RING, lj = aj;
        const uint32_t ring_parity = (((uint32_t)ch / W_RING) & 1u) ^ 1u;
        mbar_wait(&a_empty[aj], ring_parity, p.dbg, 0x700u + aj);
        mbar_wait(&full_bar[s], ph, p.dbg, 0x500u + s);
        tc_fence_after();
        const uint32_t sa = smem0 + (uint32_t)s * stage_bytes + col_base;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          float v = 1.f;
          if (!NG_EXP(1))
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sa + (uint32_t)k * 128u + ((chunk32 ^ (uint32_t)(k & 3)) << 5)));
          hi[k] = __float_as_uint(v);
          lo[k] = __float_as_uint(tf32_lo(v));
        }
        const uint32_t a_t = tmem_a0 + aj * 64u + lane_addr;
        if (!NG_EXP(4)) {
          tmem_st_32x32(a_t, hi);
          tmem_st_32x32(a_t + 32u, lo);
        }
        split_stage_hi_lo(smem + (size_t)s * stage_bytes + p.stage_a_bytes,
                          lo_ring + (size_t)lj * p.stage_b_bytes, p.stage_b_bytes, t, NG_EXP_FLAGS & 16);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[s]);
        s += 2;
        if (s >= p.stages) { s -= p.stages; ph ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int tl = m / p.Cm, cl = m - tl * p.Cm;
    const int tap = tap0 + tl;
    const int c = c_t * p.Cm + cl;
    float* dst_row = p.ws + (((long long)split * p.RS + tap) * p.C + c) * p.K + (long long)n_t * p.n_tile;
    const bool row_ok = (tl < n_taps);
    if (X3) {
      float acc[W_X3_MAX_N];
#pragma unroll
      for (int i = 0; i < W_X3_MAX_N; ++i) acc[i] = 0.f;
      for (int g = 0; g < n_groups; ++g) {
        const int b = g & 1;
        mbar_wait(&acc_full[b], (uint32_t)((g >> 1) & 1), p.dbg, 0x400u + b);
        tc_fence_after();
#pragma unroll
        for (int j0 = 0; j0 < W_X3_MAX_N; j0 += 32) {
          if (j0 < p.n_tile && !NG_EXP(32)) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * p.n_tile + j0), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[j0 + i] += __uint_as_float(r[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[b]);
      }
#pragma unroll
      for (int j0 = 0; j0 < W_X3_MAX_N; j0 += 32) {
        if (j0 < p.n_tile && row_ok && n_t * p.n_tile + j0 < p.K) {
          float4* d4 = reinterpret_cast<float4*>(dst_row + j0);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            d4[i] = make_float4(acc[j0 + 4 * i], acc[j0 + 4 * i + 1], acc[j0 + 4 * i + 2], acc[j0 + 4 * i + 3]);
        }
      }
    } else {
      if (n_chunks > 0) {
        mbar_wait(&acc_full[0], 0, p.dbg, 0x400u);
        tc_fence_after();
      }
      for (int j0 = 0; j0 < p.n_tile; j0 += 32) {
        uint32_t r[32];
        if (n_chunks > 0) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)j0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        if (row_ok && n_t * p.n_tile + j0 < p.K) {
          float4* d4 = reinterpret_cast<float4*>(dst_row + j0);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            d4[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                __uint_as_float(r[4 * i + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// dw[k][c][rs] = sum_split ws[split][rs][c][k]   (fixed summation order)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(float* __restrict__ dw, const float* __restrict__ ws, int splits,
                                                           int RS, int C, int K) {
  const long long n = (long long)RS * C * K;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(long long)s * n + e];
    const int k = (int)(e % K);
    const long long t = e / K;
    const int c = (int)(t % C);
    const int rs = (int)(t / C);
    dw[((long long)k * C + c) * RS + rs] = acc;
  }
}

int conv_wgrad_tf32(float* dw, const float* dyv, const float* x, int N, int C, int H, int W, int K, int R, int S,
                    int stride, int pad_t, int pad_l, int OH, int OW, int x3, int x_is_nhwc, int dy_is_nhwc) {
  const int RS = R * S;
  if (stride != 1 || !wgrad_supported(C, K, RS)) return -1;
  WgradUP p;
  memset(&p, 0, sizeof(p));
  p.C = C; p.K = K; p.RS = RS; p.S = S;
  p.pad_t = pad_t; p.pad_l = pad_l;
  p.Cm = (C % 128 == 0) ? 128 : ((C % 64 == 0) ? 64 : 32);
  p.taps_per_cta = 128 / p.Cm;
  p.tap_groups = (int)ceil_div(RS, p.taps_per_cta);
  p.c_tiles = C / p.Cm;
  const int n_cap = x3 ? W_X3_MAX_N : 256;
  p.n_tile = K < n_cap ? K : n_cap;
  p.n_tiles = (int)ceil_div(K, p.n_tile);
  p.x3 = x3 ? 1 : 0;
  int w_c = pow2_ceil(OW);
  if (w_c > 32) w_c = 32;
  p.w_c = w_c;
  p.rows_c = 32 / w_c;
  p.chunks_w = (int)ceil_div(OW, w_c);
  p.chunks_h = (int)ceil_div(OH, p.rows_c);
  p.total_chunks = N * p.chunks_h * p.chunks_w;
  const int items = p.tap_groups * p.c_tiles * p.n_tiles;
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  int splits = sms / items;
  if (splits < 1) splits = 1;
  if (splits > p.total_chunks) splits = p.total_chunks;
  p.chunks_per_split = (int)ceil_div(p.total_chunks, splits);
  splits = (int)ceil_div(p.total_chunks, p.chunks_per_split);
  p.splits = splits;
  // X3: the x operand is read from TMEM (K-major by construction), dy stays MN-major in shared memory
  p.idesc = make_idesc_tf32(128, p.n_tile, /*a MN-major*/ x3 ? 0 : 1, /*b MN-major*/ 1);
  p.stage_a_bytes = 128u * 128u;
  p.stage_b_bytes = (uint32_t)p.n_tile * 128u;
  const uint32_t stage_bytes = p.stage_a_bytes + p.stage_b_bytes;
  { const char* e = getenv("NG_UMMA_EXP"); p.exp_flags = e ? atoi(e) : 0; }
  int stages = (int)((227u * 1024u - 2048u - (x3 ? W_RING * p.stage_b_bytes : 0u)) / stage_bytes);
  if (stages > W_MAX_STAGES) stages = W_MAX_STAGES;
  if (stages < 2) return -1;
  p.stages = stages;
  p.tmem_cols = (uint32_t)pow2_ceil(x3 ? 2 * p.n_tile + (int)W_RING * 64 : (p.n_tile < 32 ? 32 : p.n_tile));

  float *x_t = nullptr, *dy_t = nullptr, *ws = nullptr;
  if (!x_is_nhwc && pool_alloc((void**)&x_t, (uint64_t)N * C * H * W * sizeof(float))) return 1;
  if (!dy_is_nhwc && pool_alloc((void**)&dy_t, (uint64_t)N * K * OH * OW * sizeof(float))) { pool_free(x_t); return 1; }
  if (pool_alloc((void**)&ws, (uint64_t)splits * RS * C * K * sizeof(float))) { pool_free(x_t); pool_free(dy_t); return 1; }
  p.ws = ws;
  int rc = 0;
  const float* x_nhwc = x_is_nhwc ? x : x_t;
  const float* dy_nhwc = dy_is_nhwc ? dyv : dy_t;
  if (!x_is_nhwc) rc = to_nhwc(x_t, x, N, C, H * W, x3 ? 0 : 1);
  if (rc == 0 && !dy_is_nhwc) rc = to_nhwc(dy_t, dyv, N, K, OH * OW, x3 ? 0 : 1);
  if (rc == 0) {
    const uint64_t dims[5] = {32u, (uint64_t)W, (uint64_t)H, (uint64_t)(C / 32), (uint64_t)N};
    const uint64_t strides[4] = {(uint64_t)C * 4u, (uint64_t)W * C * 4u, 128u, (uint64_t)H * W * C * 4u};
    const uint32_t box[5] = {32u, (uint32_t)w_c, (uint32_t)p.rows_c, (uint32_t)(p.Cm / 32), 1u};
    if (encode_tensor_map(&p.tm_x, x_nhwc, 5, dims, strides, box, SWZ_128B_ATOM32)) rc = -1;
  }
  if (rc == 0) {
    const uint64_t dims[5] = {32u, (uint64_t)OW, (uint64_t)OH, (uint64_t)(K / 32), (uint64_t)N};
    const uint64_t strides[4] = {(uint64_t)K * 4u, (uint64_t)OW * K * 4u, 128u, (uint64_t)OH * OW * K * 4u};
    const uint32_t box[5] = {32u, (uint32_t)w_c, (uint32_t)p.rows_c, (uint32_t)(p.n_tile / 32), 1u};
    if (encode_tensor_map(&p.tm_dy, dy_nhwc, 5, dims, strides, box, SWZ_128B_ATOM32)) rc = -1;
  }
  if (rc == 0) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(umma_conv_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(227 * 1024));
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(umma_conv_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(227 * 1024));
      if (e != cudaSuccess) rc = set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      configured = (e == cudaSuccess);
    }
  }
  if (rc == 0) {
    const size_t smem_bytes = (size_t)stages * stage_bytes + (x3 ? W_RING * p.stage_b_bytes : 0u) + 1024 + 512;
    const int grid = items * splits;
    static uint32_t* dbg_host = nullptr;
    const bool debug = getenv("NG_UMMA_DEBUG") != nullptr;
    if (debug) {
      if (!dbg_host) cudaHostAlloc((void**)&dbg_host, 64 * sizeof(uint32_t), cudaHostAllocMapped);
      memset(dbg_host, 0, 64 * sizeof(uint32_t));
      p.dbg = dbg_host;
      fprintf(stderr, "[umma] wgrad: grid %d items %d splits %d chunks %d/split Cm %d taps/cta %d n_tile %d stages %d w_c %d rows_c %d\n",
              grid, items, splits, p.chunks_per_split, p.Cm, p.taps_per_cta, p.n_tile, stages, w_c, p.rows_c);
    }
    if (x3) NG_LAUNCH(umma_conv_wgrad_kernel<true>, grid, W_THREADS_X3, smem_bytes, p);
    else NG_LAUNCH(umma_conv_wgrad_kernel<false>, grid, W_THREADS, smem_bytes, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = set_error("umma_conv_wgrad_kernel launch failed: %s", cudaGetErrorString(e));
    if (debug) {
      cudaError_t e2 = cudaStreamSynchronize(g_stream);
      fprintf(stderr, "[umma] wgrad sync: %s ; dbg: %08x %08x %08x\n", cudaGetErrorString(e2), dbg_host[0], dbg_host[1],
              dbg_host[2]);
    }
  }
  if (rc == 0) {
    const long long n = (long long)RS * C * K;
    int g = (int)ceil_div(n, 256);
    if (g > sms * 8) g = sms * 8;
    NG_LAUNCH(wgrad_reduce_kernel, g, 256, 0, dw, (const float*)ws, splits, RS, C, K);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = set_error("wgrad_reduce_kernel launch failed: %s", cudaGetErrorString(e));
  }
  pool_free(ws);
  pool_free(dy_t);
  pool_free(x_t);
  return rc;
}

}  // namespace umma
}  // namespace ng
#endif  // NG_EMU
