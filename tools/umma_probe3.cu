// Probe: tcgen05.mma with the A operand in TMEM (TS mode), written there by tcgen05.st from registers
// (thread = row m, 32 consecutive columns = 32 K elements), B K-major via TMA (128B swizzle).
//   D[128][32] = A[128][32] * B[32 j][32 k]^T   compared with a CPU reference.
#include <cstdio>
#include <cstdlib>
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
struct P3 { CUtensorMap tm_b; const float* a; float* d_out; uint32_t idesc; volatile uint32_t* dbg; };

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe3(const __grid_constant__ P3 p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sb = smem;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4096);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *slot;
  if (threadIdx.x == 0) { mbar_arrive_expect_tx(bar, 4096); tma_load_2d(sb, &p.tm_b, bar, 0, 0); }
  // A -> TMEM columns 32..63: thread m holds row m
  {
    const int m = warp * 32 + lane;
    uint32_t v[32];
    for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(p.a[m * 32 + k]);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(tmem_base + ((uint32_t)(warp * 32) << 16) + 32u), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
        "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
        "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  mbar_wait(bar, 0, p.dbg, 1);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    for (int k = 0; k < 4; ++k) {
      const uint64_t bdesc = make_smem_desc(smem_u32(sb) + k * 32, 16, 1024, SWZ_128B);
      mma_tf32_ts(tmem_base, tmem_base + 32u + (uint32_t)k * 8u, bdesc, p.idesc, acc);
      acc = 1;
    }
    mma_commit(bar2);
  }
  mbar_wait(bar2, 0, p.dbg, 2);
  tc_fence_after();
  uint32_t r[32];
  tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) p.d_out[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}
int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  std::vector<float> ha(128 * 32), hb(32 * 32);
  for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) ha[m * 32 + k] = (float)(((m * 3 + k * 5) % 17) - 8) * 0.25f;
  for (int j = 0; j < 32; ++j) for (int k = 0; k < 32; ++k) hb[j * 32 + k] = (float)(((j * 7 + k * 3) % 11) - 5) * 0.5f;
  float *da, *db, *dd; cudaMalloc(&da, ha.size() * 4); cudaMalloc(&db, hb.size() * 4); cudaMalloc(&dd, 128 * 32 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice);
  uint32_t* dbg; cudaHostAlloc((void**)&dbg, 256, cudaHostAllocMapped); memset(dbg, 0, 256);
  P3 p; memset(&p, 0, sizeof(p));
  cuuint64_t gdim[2] = {32, 32}, gstr[1] = {128}; cuuint32_t bdim[2] = {32, 32}, estr[2] = {1, 1};
  if (enc(&p.tm_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, db, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
  p.a = da; p.d_out = dd; p.dbg = dbg; p.idesc = make_idesc_tf32(128, 32, 0, 0);
  cudaFuncSetAttribute(probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  probe3<<<1, 128, 4096 + 1024 + 256>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s dbg %x %x\n", cudaGetErrorString(e), dbg[0], dbg[1]);
  if (e != cudaSuccess) return 2;
  std::vector<float> hd(128 * 32); cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0; double maxerr = 0;
  for (int m = 0; m < 128; ++m) for (int j = 0; j < 32; ++j) {
    double ref = 0; for (int k = 0; k < 32; ++k) ref += (double)ha[m * 32 + k] * hb[j * 32 + k];
    double err = fabs(ref - hd[m * 32 + j]); if (err > maxerr) maxerr = err;
    if (err > 1e-3 * (1 + fabs(ref))) { if (bad < 6) printf("  D[%d][%d] got %g want %g\n", m, j, hd[m * 32 + j], ref); bad++; }
  }
  printf("TS-mode MMA check: %d mismatches of 4096, max abs err %g\n", bad, maxerr);
  return 0;
}
