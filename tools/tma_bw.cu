// Probe: how fast can one producer thread per SM stream channels-last (NHWC) activation tiles into
// shared memory with TMA, as a function of the tensor-map rank / box shape?  No math: a consumer
// thread releases every stage as soon as it lands.  Mirrors the load pattern of the wgrad kernel
// (csrc/umma_conv.cu): per 32-pixel chunk, `taps` shifted boxes of `Cm` channels.
//   mode 0: 5-D map (32, W, H, C/32, N), box (32, w_c, rows_c, Cm/32, 1)      -- what the kernel does
//   mode 1: 4-D map (32, W, H*N, C/32),  box (32, w_c, rows_c, Cm/32)        -- image dim merged
//   mode 2: 3-D map (32, W*H*N, C/32),   box (32, 32, Cm/32)                 -- linear pixel run
//   mode 3: 2-D map (32, W*H*N) per 32-channel group, Cm/32 loads of box (32, 32)
//   mode 4: 2-D map (C, W*H*N) without swizzle, box (Cm, 32): full pixel rows (Cm*4 contiguous bytes)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/tma_bw tools/tma_bw.cu
// usage: tma_bw mode C HW Cm taps stages [N=512] [ctas=148] [sub=1: 32*sub pixels per box] [reps=1] [producers=1]
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>
#include "../nanograd_b200/csrc/umma_common.cuh"
namespace ng { int set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); return 1; } }
using namespace ng::umma;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct BwP {
  CUtensorMap tm;
  int mode, W, H, N, groups, taps, stages, w_c, rows_c, chunks_w, chunks_h, total_chunks, chunks_per_cta;
  uint32_t tap_bytes;
  int reps, producers, same_warp;
  unsigned long long* prof;   // [3] cycles of CTA 0 / producer 0 in wait, expect_tx, TMA issue
  unsigned long long* cycles;
};

__global__ void __launch_bounds__(160, 1) tma_bw_kernel(const __grid_constant__ BwP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = p.tap_bytes * (uint32_t)p.taps;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + 16;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int chunk_begin = blockIdx.x * p.chunks_per_cta;
  int chunk_end = chunk_begin + p.chunks_per_cta;
  if (chunk_end > p.total_chunks) chunk_end = p.total_chunks;
  const long long t0 = clock64();
  const int n_items = (chunk_end - chunk_begin) * p.reps;
  const bool is_prod = p.same_warp ? ((int)threadIdx.x < p.producers) : ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < p.producers);
  if (is_prod) {
    const int pid = p.same_warp ? threadIdx.x : (threadIdx.x >> 5);
    long long c_wait = 0, c_exp = 0, c_tma = 0;
    for (int it = pid; it < n_items; it += p.producers) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      const int ch = chunk_begin + it % (chunk_end - chunk_begin);
      int t = ch;
      const int cw = t % p.chunks_w; t /= p.chunks_w;
      const int chh = t % p.chunks_h;
      const int img = t / p.chunks_h;
      const int ow0 = cw * p.w_c, oh0 = chh * p.rows_c;
      long long c0 = clock64();
      mbar_wait(&empty_bar[s], ph ^ 1u, nullptr, 0);
      long long c1 = clock64();
      uint8_t* sa = smem + (size_t)s * stage_bytes;
      mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
      long long c2 = clock64();
      c_wait += c1 - c0; c_exp += c2 - c1;
      for (int tl = 0; tl < p.taps; ++tl) {
        const int r = tl / 3, sx = tl % 3;
        uint8_t* dst = sa + (size_t)tl * p.tap_bytes;
        if (p.mode == 0) tma_load_5d(dst, &p.tm, &full_bar[s], 0, ow0 + sx - 1, oh0 + r - 1, 0, img);
        else if (p.mode == 1) tma_load_4d(dst, &p.tm, &full_bar[s], 0, ow0 + sx - 1, img * p.H + oh0 + r, 0);
        else {
          int pix = (img * p.H + oh0 + r) * p.W + ow0 + sx;   // linear run (no halo semantics)
          if (pix + 32 > p.N * p.H * p.W) pix = p.N * p.H * p.W - 32;
          if (p.mode == 2) tma_load_3d(dst, &p.tm, &full_bar[s], 0, pix, 0);
          else if (p.mode == 3) {
            for (int g = 0; g < p.groups; ++g) tma_load_2d(dst + g * 4096, &p.tm, &full_bar[s], g * 32, pix);
          } else tma_load_2d(dst, &p.tm, &full_bar[s], 0, pix);
        }
      }
      c_tma += clock64() - c2;
    }
    if (blockIdx.x == 0 && pid == 0) { p.prof[0] = c_wait; p.prof[1] = c_exp; p.prof[2] = c_tma; p.prof[3] = (n_items + p.producers - 1) / p.producers; }
  } else if (threadIdx.x == 32 * 4) {
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < n_items; ++it) {
      mbar_wait(&full_bar[s], ph, nullptr, 0);
      mbar_arrive(&empty_bar[s]);
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
    p.cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
}

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: tma_bw mode C HW Cm taps stages [N] [ctas]\n"); return 2; }
  const int mode = atoi(argv[1]), C = atoi(argv[2]), HW = atoi(argv[3]), Cm = atoi(argv[4]), taps = atoi(argv[5]),
            stages = atoi(argv[6]);
  const int N = argc > 7 ? atoi(argv[7]) : 512, ctas = argc > 8 ? atoi(argv[8]) : 148;
  const int sub = argc > 9 ? atoi(argv[9]) : 1;
  const int reps = argc > 10 ? atoi(argv[10]) : 1, producers = argc > 11 ? atoi(argv[11]) : 1;
  const int H = HW, W = HW;
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const size_t elems = (size_t)N * H * W * C;
  float* dx; cudaMalloc(&dx, elems * 4); cudaMemset(dx, 0, elems * 4);
  BwP p; memset(&p, 0, sizeof(p));
  p.reps = reps; p.producers = producers;
  p.mode = mode; p.W = W; p.H = H; p.N = N; p.groups = Cm / 32; p.taps = taps; p.stages = stages;
  p.w_c = W < 32 ? W : 32; p.rows_c = 32 / p.w_c * sub;
  p.chunks_w = (W + p.w_c - 1) / p.w_c; p.chunks_h = (H + p.rows_c - 1) / p.rows_c;
  p.total_chunks = N * p.chunks_h * p.chunks_w;
  p.chunks_per_cta = (p.total_chunks + ctas - 1) / ctas;
  p.tap_bytes = (uint32_t)Cm * 32u * 4u * (uint32_t)sub;
  cuuint64_t dims[5], strides[4]; cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  int rank = 0; CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B;
  if (mode == 0) {
    rank = 5;
    cuuint64_t d[5] = {32u, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 32), (cuuint64_t)N};
    cuuint64_t st[4] = {(cuuint64_t)C * 4u, (cuuint64_t)W * C * 4u, 128u, (cuuint64_t)H * W * C * 4u};
    cuuint32_t b[5] = {32u, (cuuint32_t)p.w_c, (cuuint32_t)p.rows_c, (cuuint32_t)(Cm / 32), 1u};
    memcpy(dims, d, sizeof(d)); memcpy(strides, st, sizeof(st)); memcpy(box, b, sizeof(b));
  } else if (mode == 1) {
    rank = 4;
    cuuint64_t d[4] = {32u, (cuuint64_t)W, (cuuint64_t)H * N, (cuuint64_t)(C / 32)};
    cuuint64_t st[3] = {(cuuint64_t)C * 4u, (cuuint64_t)W * C * 4u, 128u};
    cuuint32_t b[4] = {32u, (cuuint32_t)p.w_c, (cuuint32_t)p.rows_c, (cuuint32_t)(Cm / 32)};
    memcpy(dims, d, sizeof(d)); memcpy(strides, st, sizeof(st)); memcpy(box, b, sizeof(b));
  } else if (mode == 2) {
    rank = 3;
    cuuint64_t d[3] = {32u, (cuuint64_t)W * H * N, (cuuint64_t)(C / 32)};
    cuuint64_t st[2] = {(cuuint64_t)C * 4u, 128u};
    cuuint32_t b[3] = {32u, 32u * (cuuint32_t)sub, (cuuint32_t)(Cm / 32)};
    memcpy(dims, d, sizeof(d)); memcpy(strides, st, sizeof(st)); memcpy(box, b, sizeof(b));
  } else if (mode == 3) {
    rank = 2;
    cuuint64_t d[2] = {(cuuint64_t)C, (cuuint64_t)W * H * N};
    cuuint64_t st[1] = {(cuuint64_t)C * 4u};
    cuuint32_t b[2] = {32u, 32u};
    memcpy(dims, d, sizeof(d)); memcpy(strides, st, sizeof(st)); memcpy(box, b, sizeof(b));
  } else {
    rank = 2; swz = CU_TENSOR_MAP_SWIZZLE_NONE;
    cuuint64_t d[2] = {(cuuint64_t)C, (cuuint64_t)W * H * N};
    cuuint64_t st[1] = {(cuuint64_t)C * 4u};
    cuuint32_t b[2] = {(cuuint32_t)Cm, 32u};
    memcpy(dims, d, sizeof(d)); memcpy(strides, st, sizeof(st)); memcpy(box, b, sizeof(b));
  }
  CUresult r = enc(&p.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "encode failed %d\n", (int)r); return 1; }
  cudaMalloc(&p.cycles, ctas * 8);
  cudaMalloc(&p.prof, 4 * 8);
  p.same_warp = argc > 12 ? atoi(argv[12]) : 0;
  const size_t smem = (size_t)stages * p.tap_bytes * taps + 1024 + 512;
  cudaFuncSetAttribute(tma_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(e0);
    tma_bw_kernel<<<ctas, 160, smem>>>(p);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  std::vector<unsigned long long> cyc(ctas);
  cudaMemcpy(cyc.data(), p.cycles, ctas * 8, cudaMemcpyDeviceToHost);
  unsigned long long cmax = 0; for (auto c : cyc) if (c > cmax) cmax = c;
  const double bytes = (double)p.total_chunks * taps * p.tap_bytes * reps;
  printf("mode %d C %d HW %d Cm %d taps %d stages %d sub %d reps %d prod %d (stage %u KB): %.3f ms  %.2f TB/s  %.1f B/clk/SM  (tensor %.0f MB)\n", mode, C, HW,
         Cm, taps, stages, sub, reps, producers, p.tap_bytes * taps / 1024, best, bytes / best / 1e9, bytes / ctas / (double)cmax, elems * 4 / 1e6);
  unsigned long long prof[4]; cudaMemcpy(prof, p.prof, 32, cudaMemcpyDeviceToHost);
  printf("   producer 0 of CTA 0: %llu items; cycles per item: wait %.0f, expect_tx %.0f, tma issue %.0f\n", prof[3], (double)prof[0] / prof[3], (double)prof[1] / prof[3], (double)prof[2] / prof[3]);
  return 0;
}
