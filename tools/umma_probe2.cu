// Probe: TF32 tcgen05 MMA with BOTH operands MN-major (the wgrad form on NHWC tensors):
//   D[k][c] = sum_pix A[pix][k] * B[pix][c],  A = dy_t [P][K=128], B = x_t [P][C=64], P = 32 pixels.
// Shared memory: per 32-channel group a TMA box {32 ch, 32 pix} with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
// -> [pix][32 ch] rows of 128 B; UMMA layout type 1 (SWIZZLE_128B_BASE32B), LBO = group stride,
// SBO = stride between 4-row K groups.  argv: lbo sbo layout_type tma_swizzle(0 none,3 128B,4 atom32)
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
static EncodeTiledFn g_enc;
static int encode2d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint32_t b0, uint32_t b1, int sw) {
  cuuint64_t gdim[2] = {d0, d1}, gstr[1] = {d0 * 4};
  cuuint32_t bdim[2] = {b0, b1}, estr[2] = {1, 1};
  CUresult r = g_enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, bdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "encode failed %d\n", (int)r); return 1; }
  return 0;
}
struct P2 {
  CUtensorMap tm_a, tm_b;
  float* d_out;   // 128 x 64
  float* dump;    // A stage dump (4096 floats of group 0)
  uint32_t lbo, sbo, ltype, idesc, kstep;
  volatile uint32_t* dbg;
};
__global__ void __launch_bounds__(128, 1) probe2(const __grid_constant__ P2 p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;               // 4 groups x 4 KB
  uint8_t* sb = smem + 16384;       // 2 groups x 4 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 16384 + 8192);
    for (int g = 0; g < 4; ++g) asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(sa + g * 4096)), "l"(reinterpret_cast<uint64_t>(&p.tm_a)), "r"(smem_u32(bar)), "r"(g * 32), "r"(0) : "memory");
    for (int g = 0; g < 2; ++g) asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(sb + g * 4096)), "l"(reinterpret_cast<uint64_t>(&p.tm_b)), "r"(smem_u32(bar)), "r"(g * 32), "r"(0) : "memory");
  }
  mbar_wait(bar, 0, p.dbg, 1);
  for (int i = threadIdx.x; i < 1024; i += 128) p.dump[i] = reinterpret_cast<float*>(sa)[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    uint32_t acc = 0;
    for (int k = 0; k < 4; ++k) {   // 4 MMAs of 8 pixels
      const uint64_t adesc = make_smem_desc(smem_u32(sa) + k * p.kstep, p.lbo, p.sbo, p.ltype);
      const uint64_t bdesc = make_smem_desc(smem_u32(sb) + k * p.kstep, p.lbo, p.sbo, p.ltype);
      mma_tf32(tmem_base, adesc, bdesc, p.idesc, acc);
      acc = 1;
    }
    mma_commit(bar2);
  }
  mbar_wait(bar2, 0, p.dbg, 2);
  tc_fence_after();
  for (int j0 = 0; j0 < 64; j0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + j0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.d_out[(warp * 32 + lane) * 64 + j0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}
int main(int argc, char** argv) {
  const int P = 32, K = 128, C = 64;
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  g_enc = (EncodeTiledFn)fn;
  std::vector<float> ha(P * K), hb(P * C);
  for (int i = 0; i < P; ++i) for (int k = 0; k < K; ++k) ha[i * K + k] = (float)(((i * 5 + k * 3) % 13) - 6) * 0.25f;
  for (int i = 0; i < P; ++i) for (int c = 0; c < C; ++c) hb[i * C + c] = (float)(((i * 7 + c * 11) % 9) - 4) * 0.5f;
  float *da, *db, *dd, *ddump;
  cudaMalloc(&da, ha.size() * 4); cudaMalloc(&db, hb.size() * 4); cudaMalloc(&dd, 128 * 64 * 4); cudaMalloc(&ddump, 4096);
  cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice);
  uint32_t* dbg; cudaHostAlloc((void**)&dbg, 256, cudaHostAllocMapped); memset(dbg, 0, 256);
  P2 p; memset(&p, 0, sizeof(p));
  p.lbo = argc > 1 ? atoi(argv[1]) : 4096;
  p.sbo = argc > 2 ? atoi(argv[2]) : 512;
  p.ltype = argc > 3 ? atoi(argv[3]) : 1;
  const int sw = argc > 4 ? atoi(argv[4]) : 4;
  p.kstep = argc > 5 ? atoi(argv[5]) : 1024;
  if (encode2d(&p.tm_a, da, K, P, 32, 32, sw)) return 1;
  if (encode2d(&p.tm_b, db, C, P, 32, 32, sw)) return 1;
  p.d_out = dd; p.dump = ddump; p.dbg = dbg;
  p.idesc = make_idesc_tf32(128, 64, 1, 1);
  cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe2<<<1, 128, 16384 + 8192 + 1024 + 256>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  printf("lbo %u sbo %u ltype %u tma_swz %d kstep %u: sync: %s dbg %x %x %x\n", p.lbo, p.sbo, p.ltype, sw, p.kstep, cudaGetErrorString(e), dbg[0], dbg[1], dbg[2]);
  if (e != cudaSuccess) return 2;
  std::vector<float> hd(128 * 64), hdump(1024);
  cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hdump.data(), ddump, 4096, cudaMemcpyDeviceToHost);
  // smem image of A group 0: row = pixel (128 B), find where element (pix, k) landed
  printf("smem A rows 0..5 positions of channels 0,8,16,24 (index within row):");
  for (int row = 0; row < 6; ++row) {
    printf(" |");
    for (int kk = 0; kk < 32; kk += 8) {
      int pos = -1;
      for (int i = 0; i < 32; ++i) {
        // identify by comparing the 4 consecutive values (chunk of 16B): values repeat mod 13 so compare 8 values
        bool ok = true;
        for (int t = 0; t < 8 && ok; ++t) if (i + t >= 32 + 0 || hdump[row * 32 + ((i + t) % 32)] != ha[row * K + ((kk + t) % 32)]) ok = (i + t < 32) ? (hdump[row * 32 + i + t] == ha[row * K + kk + t]) : true;
        if (ok) { pos = i; break; }
      }
      printf(" %d", pos);
    }
  }
  printf("\n");
  int bad = 0; double maxerr = 0;
  for (int k = 0; k < K; ++k) for (int c = 0; c < C; ++c) {
    double ref = 0;
    for (int i = 0; i < P; ++i) ref += (double)ha[i * K + k] * hb[i * C + c];
    const double err = fabs(ref - hd[k * 64 + c]);
    if (err > maxerr) maxerr = err;
    if (err > 1e-3 * (1 + fabs(ref))) { if (bad < 4) printf("  D[%d][%d] got %g want %g\n", k, c, hd[k * 64 + c], ref); bad++; }
  }
  int okg[4][2] = {};
  for (int k = 0; k < K; ++k) for (int c = 0; c < C; ++c) {
    double ref = 0;
    for (int i = 0; i < P; ++i) ref += (double)ha[i * K + k] * hb[i * C + c];
    if (fabs(ref - hd[k * 64 + c]) <= 1e-3 * (1 + fabs(ref))) okg[k / 32][c / 32]++;
  }
  printf("MMA check: %d mismatches of %d, max abs err %g ; ok per (m-group, n-group):", bad, K * C, maxerr);
  for (int a = 0; a < 4; ++a) for (int b = 0; b < 2; ++b) printf(" %d", okg[a][b]);
  printf("\n");
  return 0;
}
