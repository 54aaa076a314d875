// Probe: tcgen05.mma.cta_group::2 (a CTA pair, M = 256) with the protocol the conv kernels would need:
//   * each CTA TMA-loads its own 128 rows of A and ITS HALF of B's N rows into its own shared memory,
//     signalling its own mbarrier;
//   * a thread of each CTA then arrives on the LEADER's "ready" barrier (remote arrive through
//     mapa / shared::cluster for rank 1) -- the stand-in for the operand-split warps;
//   * the leader issues the MMAs (cta_group::2) and commits with .multicast::cluster to a barrier in
//     both CTAs; each CTA reads its own 128 accumulator rows from its own TMEM.
//   D[256][N] = A[256][32] * B[N][32]^T, K-major operands, 128B swizzle, compared with a CPU reference.
// mode 0: SS (A from shared memory);  mode 1: TS (A from TMEM, written by tcgen05.st in each CTA).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/bin/umma_pair_probe tools/umma_pair_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>
#include "../nanograd_b200/csrc/umma_common.cuh"
namespace ng { int set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); return 1; } }
using namespace ng::umma;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct PairP {
  CUtensorMap tm_a, tm_b;   // A [256][32], B [N][32]
  float* d_out;             // [256][N]
  int N, mode;
  uint32_t idesc;
  volatile uint32_t* dbg;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, volatile uint32_t* dbg, uint32_t who) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 22)) { if (dbg) { dbg[0] = 0xDEAD0000u | who; __threadfence_system(); } __trap(); }
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_probe(const __grid_constant__ PairP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                 // 128 rows x 128 B = 16 KB
  uint8_t* sb = smem + 16384;         // (N/2) rows x 128 B
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 16384 + 16384);
  uint64_t* ready_bar = full_bar + 1;   // leader's: 2 arrivals (one per CTA)
  uint64_t* done_bar = full_bar + 2;    // in both CTAs: MMA completion (multicast commit)
  uint32_t* slot = reinterpret_cast<uint32_t*>(full_bar + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int half = p.N / 2;
  if (threadIdx.x == 0) {
    mbar_init(full_bar, 1);
    mbar_init(ready_bar, 2);
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc2(slot, 256); tmem_relinquish2(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(full_bar, 16384u + (uint32_t)half * 128u);
    tma_load_2d(sa, &p.tm_a, full_bar, 0, (int)rank * 128);
    tma_load_2d(sb, &p.tm_b, full_bar, 0, (int)rank * half);
  }
  // everyone waits for the local tiles
  mbar_wait(full_bar, 0, p.dbg, 1);
  if (p.mode == 1) {
    // TS mode: thread m copies row m of the (swizzled) A tile into TMEM columns [128, 160) (after the accumulator)
    const int m = threadIdx.x;
    uint32_t v[32];
    const uint32_t row = smem_u32(sa) + (uint32_t)m * 128u;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 f = lds128(row + (uint32_t)((c ^ (m & 7)) << 4));
      v[4 * c] = __float_as_uint(f.x); v[4 * c + 1] = __float_as_uint(f.y);
      v[4 * c + 2] = __float_as_uint(f.z); v[4 * c + 3] = __float_as_uint(f.w);
    }
    tmem_st_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 128u, v);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive_cluster(mapa_u32(smem_u32(ready_bar), 0));
  if (rank == 0 && warp == 1) {
    mbar_wait_cluster(ready_bar, 0, p.dbg, 2);
    tc_fence_after();
    if (elect_one()) {
      uint32_t acc = 0;
      for (int k = 0; k < 4; ++k) {
        const uint64_t bdesc = make_smem_desc(smem_u32(sb) + k * 32, 16, 1024, SWZ_128B);
        if (p.mode == 0) {
          const uint64_t adesc = make_smem_desc(smem_u32(sa) + k * 32, 16, 1024, SWZ_128B);
          mma2_tf32_ss(tmem_base, adesc, bdesc, p.idesc, acc);
        } else {
          mma2_tf32_ts(tmem_base, tmem_base + 128u + (uint32_t)k * 8u, bdesc, p.idesc, acc);
        }
        acc = 1;
      }
      mma2_commit_multicast(done_bar, 3);
    }
    __syncwarp();
  }
  mbar_wait_cluster(done_bar, 0, p.dbg, 3);
  tc_fence_after();
  for (int j0 = 0; j0 < p.N; j0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)j0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.d_out[((int)rank * 128 + warp * 32 + lane) * p.N + j0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc2(tmem_base, 256); }
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int N = argc > 2 ? atoi(argv[2]) : 64;
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  std::vector<float> ha(256 * 32), hb(N * 32);
  for (int m = 0; m < 256; ++m) for (int k = 0; k < 32; ++k) ha[m * 32 + k] = (float)(((m * 3 + k * 5) % 17) - 8) * 0.25f;
  for (int j = 0; j < N; ++j) for (int k = 0; k < 32; ++k) hb[j * 32 + k] = (float)(((j * 7 + k * 3) % 11) - 5) * 0.5f;
  float *da, *db, *dd;
  cudaMalloc(&da, ha.size() * 4); cudaMalloc(&db, hb.size() * 4); cudaMalloc(&dd, 256 * N * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0xff, 256 * N * 4);
  PairP p; memset(&p, 0, sizeof(p));
  p.d_out = dd; p.N = N; p.mode = mode;
  p.idesc = make_idesc_tf32(256, N, 0, 0);
  {
    cuuint64_t dims[2] = {32, 256}, str[1] = {128}; cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
    CUresult r = enc(&p.tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, da, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "encode a failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {32, (cuuint64_t)N}, str[1] = {128}; cuuint32_t box[2] = {32, (cuuint32_t)(N / 2)}, es[2] = {1, 1};
    CUresult r = enc(&p.tm_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, db, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "encode b failed %d\n", (int)r); return 1; }
  }
  uint32_t* dbg; cudaHostAlloc((void**)&dbg, 64, cudaHostAllocMapped); memset(dbg, 0, 64);
  p.dbg = dbg;
  cudaFuncSetAttribute(pair_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  pair_probe<<<2, 128, 48 * 1024>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %d N %d: sync: %s dbg %08x idesc %08x\n", mode, N, cudaGetErrorString(e), dbg[0], p.idesc);
  if (e != cudaSuccess) return 1;
  std::vector<float> hd(256 * N);
  cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < 256; ++m) for (int j = 0; j < N; ++j) {
    double ref = 0; for (int k = 0; k < 32; ++k) ref += (double)ha[m * 32 + k] * hb[j * 32 + k];
    const double err = fabs(ref - hd[m * N + j]);
    if (!(err <= 1e-3)) { if (bad < 6) printf("  D[%d][%d] = %g, want %g\n", m, j, hd[m * N + j], ref); ++bad; }
    if (err > maxerr) maxerr = err;
  }
  printf("mode %d N %d: max err %g, %d bad of %d\n", mode, N, maxerr, bad, 256 * N);
  return bad != 0;
}
