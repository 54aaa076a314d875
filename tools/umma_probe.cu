// Hardware probe for the building blocks of nanograd_b200/csrc/umma_conv.cu (run on a B200):
//   1. a 4-D TMA box of an NCHW tensor declared in dimension order (W, C, H, N) lands in shared
//      memory as [row][c][w] with the 128B swizzle (dumped and checked on the host);
//   2. that tile is a valid MN-major tcgen05 A operand and a [j][c] tile a valid K-major B operand:
//      D[128 px][32 j] = A * B is compared with a CPU reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I nanograd_b200/csrc -o tools/bin/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>
#include "../nanograd_b200/csrc/umma_common.cuh"

namespace ng {
int set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); return 1; }
}
using namespace ng::umma;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_enc;

static int encode(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                  const uint32_t* box, CUtensorMapSwizzle sw) {
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; if (i + 1 < rank) gstr[i] = strides[i]; }
  CUresult r = g_enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "encode failed %d\n", (int)r); return 1; }
  return 0;
}

struct ProbeP {
  CUtensorMap tm_a, tm_b;
  float* dump_a;  // 4096 floats
  float* dump_b;  // 32*32 floats
  float* d_out;   // 128 x 32
  int cx, cy;
  int do_mma;
  int which;        // bit 0: load A, bit 1: load B
  const CUtensorMap* g_a;  // tensor maps in global memory (used when which & 4)
  const CUtensorMap* g_b;
  uint32_t a_lbo, a_sbo, a_swz, idesc;
  int prefill;
  volatile uint32_t* dbg;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ ProbeP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* sa = reinterpret_cast<float*>(smem);             // 16 KB
  float* sb = reinterpret_cast<float*>(smem + 16384);     // 4 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 4096);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
    p.dbg[0] = 1;
  }
  if (warp == 0) { tmem_alloc(slot, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot;
  if (threadIdx.x == 0) {
    p.dbg[1] = 2; __threadfence_system();
    const CUtensorMap* ta = (p.which & 4) ? p.g_a : &p.tm_a;
    const CUtensorMap* tb = (p.which & 4) ? p.g_b : &p.tm_b;
    mbar_arrive_expect_tx(bar, ((p.which & 1) ? 16384 : 0) + ((p.which & 2) ? 4096 : 0));
    p.dbg[2] = 3; __threadfence_system();
    if (p.which & 1) tma_load_4d(sa, ta, bar, p.cx, 0, p.cy, 0);
    p.dbg[3] = 4; __threadfence_system();
    if (p.which & 2) tma_load_3d(sb, tb, bar, 0, 0, 0);
    p.dbg[4] = 5; __threadfence_system();
  }
  mbar_wait(bar, 0, p.dbg + 16, 1);
  if (threadIdx.x == 0) { p.dbg[5] = 6; __threadfence_system(); }
  for (int i = threadIdx.x; i < 4096; i += 128) p.dump_a[i] = sa[i];
  for (int i = threadIdx.x; i < 1024; i += 128) p.dump_b[i] = sb[i];
  if (!p.do_mma) { __syncthreads(); if (warp == 0) tmem_dealloc(tmem_base, 32); return; }
  if (p.prefill) {
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint((float)(threadIdx.x * 100 + j));
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(tmem_base + ((uint32_t)(warp * 32) << 16)), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
        "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
        "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    uint32_t acc = p.prefill ? 1u : 0u;
    for (int k = 0; k < 4; ++k) {
      const uint64_t adesc = make_smem_desc(smem_u32(sa) + k * p.a_sbo, p.a_lbo, p.a_sbo, p.a_swz);
      const uint64_t bdesc = make_smem_desc(smem_u32(sb) + k * 32, 16, 1024, SWZ_128B);
      mma_tf32(tmem_base, adesc, bdesc, p.idesc, acc);
      acc = 1;
    }
    p.dbg[6] = 7; __threadfence_system();
    mma_commit(bar2);
    p.dbg[7] = 8; __threadfence_system();
  }
  mbar_wait(bar2, 0, p.dbg + 16, 2);
  tc_fence_after();
  uint32_t r[32];
  tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) p.d_out[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 32);
}

int main(int argc, char** argv) {
  const int N = 2, C = 32, H = 8, W = 32, J = 32;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  g_enc = (EncodeTiledFn)fn;
  if (!g_enc) { printf("no encode fn\n"); return 1; }
  std::vector<float> hx(N * C * H * W), hw(J * C);
  for (int n = 0; n < N; ++n) for (int c = 0; c < C; ++c) for (int h = 0; h < H; ++h) for (int w = 0; w < W; ++w)
    hx[((n * C + c) * H + h) * W + w] = (float)(c * 0.125 + h * 8 + w * 0.25 + 1);   // exact in TF32 (small ints / 8)
  for (int j = 0; j < J; ++j) for (int c = 0; c < C; ++c) hw[j * C + c] = (float)(((j * 7 + c * 3) % 11) - 5) * 0.5f;
  float *dx, *dw, *dump_a, *dump_b, *d_out;
  cudaMalloc(&dx, hx.size() * 4); cudaMalloc(&dw, hw.size() * 4);
  cudaMalloc(&dump_a, 4096 * 4); cudaMalloc(&dump_b, 1024 * 4); cudaMalloc(&d_out, 128 * 32 * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
  uint32_t* dbg;
  cudaHostAlloc((void**)&dbg, 64 * 4, cudaHostAllocMapped);
  int variant = argc > 1 ? atoi(argv[1]) : 0;   // 0: permuted dims (W,C,H,N); 1: natural dims (W,H,C,N) for comparison
  int do_mma = argc > 2 ? atoi(argv[2]) : 1;
  ProbeP p;
  memset(&p, 0, sizeof(p));
  if (variant == 0) {
    const uint64_t dims[4] = {W, C, H, N};
    const uint64_t strides[3] = {(uint64_t)H * W * 4, (uint64_t)W * 4, (uint64_t)C * H * W * 4};
    const uint32_t box[4] = {32, 32, 4, 1};
    if (encode(&p.tm_a, dx, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  } else {
    const uint64_t dims[4] = {W, H, C, N};
    const uint64_t strides[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)C * H * W * 4};
    const uint32_t box[4] = {32, 4, 32, 1};
    if (encode(&p.tm_a, dx, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    const uint64_t dims[3] = {C, J, 1};
    const uint64_t strides[2] = {(uint64_t)C * 4, (uint64_t)J * C * 4};
    const uint32_t box[3] = {32, 32, 1};
    if (encode(&p.tm_b, dw, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  p.dump_a = dump_a; p.dump_b = dump_b; p.d_out = d_out;
  p.cx = argc > 4 ? atoi(argv[4]) : -1; p.cy = argc > 5 ? atoi(argv[5]) : 1;
  p.do_mma = do_mma;
  p.which = argc > 3 ? atoi(argv[3]) : 3;
  {
    CUtensorMap* g;
    cudaMalloc(&g, 2 * sizeof(CUtensorMap));
    cudaMemcpy(g, &p.tm_a, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    cudaMemcpy(g + 1, &p.tm_b, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    p.g_a = g; p.g_b = g + 1;
  }
  printf("which %d sizeof(ProbeP) %zu alignof %zu\n", p.which, sizeof(ProbeP), alignof(ProbeP));
  p.a_lbo = argc > 6 ? atoi(argv[6]) : 32 * 128; p.a_sbo = argc > 7 ? atoi(argv[7]) : 1024; p.a_swz = SWZ_128B;
  printf("lbo %u sbo %u\n", p.a_lbo, p.a_sbo);
  p.idesc = make_idesc_tf32(128, 32, argc > 8 ? atoi(argv[8]) : 1, 0);
  p.prefill = argc > 9 ? atoi(argv[9]) : 0;
  printf("idesc %08x prefill %d\n", p.idesc, p.prefill);
  p.dbg = dbg;
  memset(dbg, 0, 64 * 4);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe_kernel<<<1, 128, 16384 + 4096 + 1024 + 256>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  printf("variant %d do_mma %d: sync: %s ; dbg:", variant, do_mma, cudaGetErrorString(e));
  for (int i = 0; i < 20; ++i) printf(" %x", dbg[i]);
  printf("\n");
  if (e != cudaSuccess) return 2;
  std::vector<float> ha(4096), hb(1024), hd(128 * 32);
  cudaMemcpy(ha.data(), dump_a, 4096 * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hb.data(), dump_b, 1024 * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hd.data(), d_out, 128 * 32 * 4, cudaMemcpyDeviceToHost);
  // expected smem image for variant 0: [row r][c][w] with 16B chunks XOR-swizzled by (c & 7)
  auto xval = [&](int c, int h, int w) -> float {
    if (h < 0 || h >= H || w < 0 || w >= W) return 0.f;
    return hx[((0 * C + c) * H + h) * W + w];
  };
  int bad = 0;
  for (int r = 0; r < 4 && variant == 0; ++r) for (int c = 0; c < 32; ++c) for (int w = 0; w < 32; ++w) {
    const int row = r * 32 + c;                       // 128-byte row index
    const int chunk = (w / 4) ^ (row & 7);
    const float got = ha[row * 32 + chunk * 4 + (w & 3)];
    const float want = xval(c, p.cy + r, p.cx + w);
    if (got != want && bad++ < 8) printf("A mismatch r%d c%d w%d got %g want %g\n", r, c, w, got, want);
  }
  printf("A layout check: %d mismatches\n", bad);
  printf("A raw first 3 rows (first 12 floats):\n");
  for (int row = 0; row < 3; ++row) { for (int i = 0; i < 12; ++i) printf(" %7.3f", ha[row * 32 + i]); printf("\n"); }
  printf(" row 32:"); for (int i = 0; i < 12; ++i) printf(" %7.3f", ha[32 * 32 + i]); printf("\n");
  if (do_mma) {
    int badd = 0; double maxerr = 0;
    for (int m = 0; m < 128; ++m) for (int j = 0; j < 32; ++j) {
      const int r = m / 32, w = m % 32;
      double ref = 0;
      for (int c = 0; c < 32; ++c) ref += (double)xval(c, p.cy + r, p.cx + w) * hw[j * C + c];
      const double err = fabs(ref - hd[m * 32 + j]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-3 * (1 + fabs(ref)) && badd++ < 8) printf("D mismatch m%d j%d got %g want %g\n", m, j, hd[m * 32 + j], ref);
    }
    printf("MMA check: %d mismatches, max abs err %g\n", badd, maxerr);
    for (int mm = 0; mm < 128; mm += 32) { printf("  D[%d][0..7] =", mm); for (int j = 0; j < 8; ++j) printf(" %g", hd[mm * 32 + j]); printf("\n"); }
    { int badb = 0; for (int j = 0; j < 32; ++j) for (int c = 0; c < 32; ++c) { int chunk = (c / 4) ^ (j & 7); if (hb[j * 32 + chunk * 4 + (c & 3)] != hw[j * C + c]) badb++; } printf("  B layout check: %d mismatches\n", badb); }
    for (int g = 0; g < 4; ++g) {
      int zeros = 0, ok = 0;
      for (int m = g * 32; m < g * 32 + 32; ++m) for (int j = 0; j < 32; ++j) {
        const int r = m / 32, w = m % 32;
        double ref = 0;
        for (int c = 0; c < 32; ++c) ref += (double)xval(c, p.cy + r, p.cx + w) * hw[j * C + c];
        if (hd[m * 32 + j] == 0.f) zeros++;
        if (fabs(ref - hd[m * 32 + j]) <= 1e-3 * (1 + fabs(ref))) ok++;
      }
      printf("  m-group %d: ok %d zeros %d of 1024\n", g, ok, zeros);
    }
    // which source row does D row 32 look like?  try every (r, w)
    for (int mm = 32; mm <= 96; mm += 32) {
      for (int r = -1; r < 6; ++r) for (int w = 0; w < 32; ++w) {
        int match = 0;
        for (int j = 0; j < 32; ++j) {
          double ref = 0;
          for (int c = 0; c < 32; ++c) ref += (double)xval(c, p.cy + r, p.cx + w) * hw[j * C + c];
          if (fabs(ref - hd[mm * 32 + j]) <= 1e-3 * (1 + fabs(ref))) match++;
        }
        if (match == 32) printf("  D row %d equals source (row %d, w %d)\n", mm, r, w);
      }
    }
  }
  return 0;
}
