// Microbenchmark: shared-memory wavefronts per LDS.128 / LDS.64 for the warp access patterns a
// register-tiled SGEMM can use on sm_100a.  Build: nvcc -arch=sm_100a -O3 -o smem_probe smem_probe.cu
// Prints cycles per warp-level load at saturation (8 warps per SM) for each pattern.
#include <cstdio>
#include <cuda_runtime.h>

template <int VEC>
__global__ void probe(const int* __restrict__ lane_off, float* out, long long* cycles, int iters) {
  __shared__ __align__(16) float smem[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) smem[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int off = lane_off[lane];  // in floats
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  const float* p = smem + off;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const float* q = p + ((u * 64) & 2047);
      if (VEC == 4) {
        float4 v;
        unsigned a = (unsigned)__cvta_generic_to_shared(q);
        asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
        acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w;
      } else if (VEC == 2) {
        float2 v;
        unsigned a = (unsigned)__cvta_generic_to_shared(q);
        asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
        acc0 += v.x; acc1 += v.y;
      } else {
        acc0 += *reinterpret_cast<const volatile float*>(q);
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}

int main() {
  struct Pat { const char* name; int vec; int (*f)(int); };
  Pat pats[] = {
      {"v4 lane&7 (8 distinct, quarters identical)  [A frag, 8x4 lanes]", 4, [](int l) { return (l & 7) * 4; }},
      {"v4 lane>>3 (4 distinct, quarter-uniform)     [B frag, 8x4 lanes]", 4, [](int l) { return (l >> 3) * 4; }},
      {"v4 lane>>2 (8 distinct, 2 per quarter)       [A frag, 4x8 swapped]", 4, [](int l) { return (l >> 2) * 4; }},
      {"v4 lane&3 (4 distinct, 4 per quarter)        [B frag, 4x8 swapped]", 4, [](int l) { return (l & 3) * 4; }},
      {"v4 lane&15 (16 distinct, halves identical)", 4, [](int l) { return (l & 15) * 4; }},
      {"v4 lane>>1 (16 distinct, pairs)", 4, [](int l) { return (l >> 1) * 4; }},
      {"v4 all distinct", 4, [](int l) { return l * 4; }},
      {"v4 all same", 4, [](int l) { return 0; }},
      {"v4 ((l&1)|((l>>3)&2))... 2x2 blocks: (l&1)*4 + ((l>>4)&1)*8", 4, [](int l) { return (l & 1) * 4 + ((l >> 4) & 1) * 8; }},
      {"v2 lane&15 (16 distinct float2, halves identical)", 2, [](int l) { return (l & 15) * 2; }},
      {"v2 all distinct", 2, [](int l) { return l * 2; }},
      {"v2 lane>>1", 2, [](int l) { return (l >> 1) * 2; }},
      {"v1 all distinct", 1, [](int l) { return l; }},
      {"v1 lane&7", 1, [](int l) { return l & 7; }},
  };
  int* d_off;
  float* d_out;
  long long* d_cyc;
  cudaMalloc(&d_off, 32 * sizeof(int));
  cudaMalloc(&d_out, 148 * 256 * sizeof(float));
  cudaMalloc(&d_cyc, sizeof(long long));
  const int iters = 2000;
  for (auto& p : pats) {
    int h[32];
    for (int l = 0; l < 32; ++l) h[l] = p.f(l);
    cudaMemcpy(d_off, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) {
      if (p.vec == 4) probe<4><<<148, 256>>>(d_off, d_out, d_cyc, iters);
      else if (p.vec == 2) probe<2><<<148, 256>>>(d_off, d_out, d_cyc, iters);
      else probe<1><<<148, 256>>>(d_off, d_out, d_cyc, iters);
      cudaDeviceSynchronize();
    }
    long long c = 0;
    cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    // 8 warps per SM each issue iters*16 loads; the SM's LSU serves them all
    const double per_load = (double)c / (iters * 16.0) / 8.0;
    printf("%-72s cycles per warp-load at the SM: %.2f\n", p.name, per_load);
  }
  return 0;
}
