// FP32 SIMT contraction core shared by convolution (fprop / dgrad / wgrad) and GEMM.
//
// One CTA of 256 threads computes a 128 (I) x 16*TN (J) output tile, where I is the dimension
// that is contiguous in the output tensor (output pixels for fprop/dgrad, the flattened C*R*S
// filter index for wgrad, the column index for a row-major GEMM) and J is the other one.
// K-tiles of 8 are staged global -> registers -> shared memory with double buffering; each
// thread accumulates an 8 x TN register micro-tile with FFMA.  Warps are laid out 2 (I) x 4 (J)
// and lanes 8 (I) x 4 (J) so that every shared-memory fragment read is a 128-bit load that is
// either contiguous across the 8 I-lanes (128 B, one wavefront) or a broadcast.
//
// Loader concept (one per problem type):
//   void fetch();                         // global -> registers for the next K-tile, advances k
//   void commit(float (*As)[LDA], float (*Bs)[LDB]);   // registers -> one shared-memory stage
#pragma once
#include "ng_common.h"

namespace ng {

constexpr int CT_BK = 8;
constexpr int CT_BM = 128;
constexpr int CT_LDA = CT_BM + 4;
constexpr int CT_THREADS = 256;

template <int TN>
struct CtTile {
  static constexpr int BN = 16 * TN;
  static constexpr int LDB = BN + 4;
  // B-tile elements each thread stages per K-tile (ceil(8*BN/256))
  static constexpr int NB = (CT_BK * BN + CT_THREADS - 1) / CT_THREADS;
};

struct CtCoord {
  int i0;  // first I offset of this thread inside the CTA tile (second fragment at i0 + 32)
  int j0;  // first J offset inside the CTA tile
};

template <int TN>
__device__ __forceinline__ CtCoord ct_coord(int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  const int wi = warp & 1, wj = warp >> 1, lx = lane & 7, ly = lane >> 3;
  CtCoord c;
  c.i0 = wi * 64 + lx * 4;
  if (TN == 8) c.j0 = wj * 32 + ly * 4;
  else if (TN == 4) c.j0 = wj * 16 + ly * 4;
  else if (TN == 2) c.j0 = wj * 8 + ly * 2;
  else c.j0 = wj * 4 + ly;
  return c;
}

// J offset (inside the CTA tile) of accumulator column n
template <int TN>
__device__ __forceinline__ int ct_jcol(const CtCoord& c, int n) {
  if (TN == 8) return c.j0 + (n < 4 ? n : 12 + n);
  return c.j0 + n;
}
// I offset of accumulator row m
__device__ __forceinline__ int ct_irow(const CtCoord& c, int m) { return c.i0 + (m < 4 ? m : 28 + m); }

template <int TN>
__device__ __forceinline__ void ct_compute_stage(const float (*As)[CT_LDA], const float (*Bs)[CtTile<TN>::LDB],
                                                 const CtCoord& c, float (&acc)[8][TN]) {
#pragma unroll
  for (int k = 0; k < CT_BK; ++k) {
    float a[8], b[TN];
    *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[k][c.i0]);
    *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[k][c.i0 + 32]);
    if (TN == 8) {
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[k][c.j0]);
      *reinterpret_cast<float4*>(&b[TN == 8 ? 4 : 0]) = *reinterpret_cast<const float4*>(&Bs[k][c.j0 + 16]);
    } else if (TN == 4) {
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[k][c.j0]);
    } else if (TN == 2) {
      *reinterpret_cast<float2*>(&b[0]) = *reinterpret_cast<const float2*>(&Bs[k][c.j0]);
    } else {
      b[0] = Bs[k][c.j0];
    }
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
      for (int n = 0; n < TN; ++n) acc[m][n] = fmaf(a[m], b[n], acc[m][n]);
  }
}

// Double-buffered main loop: global loads of K-tile t+1 are in flight while tile t is computed.
template <int TN, class Loader>
__device__ __forceinline__ void ct_mainloop(Loader& ld, int num_kt, const CtCoord& c, float (&acc)[8][TN]) {
  NG_SHARED16 float As[2][CT_BK][CT_LDA];
  NG_SHARED16 float Bs[2][CT_BK][CtTile<TN>::LDB];
#pragma unroll
  for (int m = 0; m < 8; ++m)
#pragma unroll
    for (int n = 0; n < TN; ++n) acc[m][n] = 0.f;
  if (num_kt <= 0) return;
  ld.fetch();
  ld.commit(As[0], Bs[0]);
  __syncthreads();
  int cur = 0;
  for (int kt = 0; kt < num_kt; ++kt) {
    const bool more = (kt + 1 < num_kt);
    if (more) ld.fetch();
    ct_compute_stage<TN>(As[cur], Bs[cur], c, acc);
    if (more) ld.commit(As[cur ^ 1], Bs[cur ^ 1]);
    __syncthreads();
    cur ^= 1;
  }
}

// Position inside a reduction index that factors as (slow, fast) with fast in [0, period):
// advanced by a constant step per K-tile without divisions.
struct CtKPos {
  int slow, fast;
  __device__ __forceinline__ void set(int kk, int period) { slow = kk / period; fast = kk - slow * period; }
  __device__ __forceinline__ void advance(int step, int period) {
    fast += step;
    while (fast >= period) { fast -= period; ++slow; }
  }
};

}  // namespace ng
