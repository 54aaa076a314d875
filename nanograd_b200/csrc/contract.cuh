// FP32 SIMT contraction core shared by convolution (fprop / dgrad / wgrad) and GEMM.
//
// One CTA of 256 threads (8 warps laid out WI x WJ) computes a BM x BN output tile with
// BM = 64*WI and BN = 4*TN*WJ, where I is the dimension that is contiguous in the output tensor
// (output pixels for fprop/dgrad, the flattened C*R*S filter index for wgrad, the column index
// of a row-major GEMM) and J is the other one.  K-tiles of 8 are staged global -> registers ->
// shared memory with double buffering; every thread accumulates an 8 x TN register micro-tile
// with FFMA (8 x 8 for all large problems).
//
// Lane layout.  Measured on B200 (tools/smem_probe.cu): a warp-wide LDS.128 costs 2 LSU cycles
// when every aligned group of 4 lanes reads at most 2 distinct 16-byte words, and 4 cycles
// otherwise.  The 32 lanes are therefore arranged 8 (I) x 4 (J) with each aligned 4-lane group
// a 2 x 2 block: lx = b0 | b2<<1 | b4<<2, ly = b1 | b3<<1.  Every A- and B-fragment read is then
// a 2-cycle LDS.128: 8 LSU cycles per 64 FFMA-issue cycles per warp for the 8 x 8 micro-tile.
//
// Loader concept (one per problem type):
//   void fetch();                                      // global -> registers, advances one K-tile
//   void commit(float* As, float* Bs);                 // registers -> one shared-memory stage
//   As is [8][LDA] (k-major, I contiguous), Bs is [8][LDB].
#pragma once
#include "ng_common.h"

namespace ng {

constexpr int CT_BK = 8;
constexpr int CT_THREADS = 256;

template <int WI_, int WJ_, int TN_>
struct CtTile {
  static_assert(WI_ * WJ_ == 8, "8 warps per CTA");
  static constexpr int WI = WI_, WJ = WJ_, TN = TN_;
  static constexpr int BM = 64 * WI;
  static constexpr int BN = 4 * TN * WJ;
  static constexpr int LDA = BM + 4;
  static constexpr int LDB = BN + 4;
  static constexpr int EA = BM / 32;                                       // A elements / thread / K-tile
  static constexpr int EB = (CT_BK * BN + CT_THREADS - 1) / CT_THREADS;    // B elements / thread / K-tile
  static constexpr int STAGE_A = CT_BK * LDA;
  static constexpr int STAGE_B = CT_BK * LDB;
  static constexpr int TILE_FLOATS = 2 * STAGE_A + 2 * STAGE_B;
};

struct CtCoord {
  int i0;  // first I offset of this thread inside the CTA tile (second fragment at i0 + 32)
  int j0;  // first J offset inside the CTA tile
};

template <class T>
__device__ __forceinline__ CtCoord ct_coord(int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  const int wi = warp % T::WI, wj = warp / T::WI;
  const int lx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4);
  const int ly = ((lane >> 1) & 1) | ((lane >> 2) & 2);
  CtCoord c;
  c.i0 = wi * 64 + lx * 4;
  if (T::TN == 8) c.j0 = wj * 32 + ly * 4;
  else if (T::TN == 4) c.j0 = wj * 16 + ly * 4;
  else if (T::TN == 2) c.j0 = wj * 8 + ly * 2;
  else c.j0 = wj * 4 + ly;
  return c;
}

// J offset (inside the CTA tile) of accumulator column n
template <class T>
__device__ __forceinline__ int ct_jcol(const CtCoord& c, int n) {
  if (T::TN == 8) return c.j0 + (n < 4 ? n : 12 + n);
  return c.j0 + n;
}

template <class T>
__device__ __forceinline__ void ct_compute_stage(const float* __restrict__ As, const float* __restrict__ Bs,
                                                 const CtCoord& c, float (&acc)[8][T::TN]) {
#pragma unroll
  for (int k = 0; k < CT_BK; ++k) {
    float a[8], b[T::TN];
    const float* ak = As + k * T::LDA + c.i0;
    const float* bk = Bs + k * T::LDB + c.j0;
    *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(ak);
    *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(ak + 32);
    if (T::TN == 8) {
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(bk);
      *reinterpret_cast<float4*>(&b[T::TN == 8 ? 4 : 0]) = *reinterpret_cast<const float4*>(bk + 16);
    } else if (T::TN == 4) {
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(bk);
    } else if (T::TN == 2) {
      *reinterpret_cast<float2*>(&b[0]) = *reinterpret_cast<const float2*>(bk);
    } else {
      b[0] = bk[0];
    }
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
      for (int n = 0; n < T::TN; ++n) acc[m][n] = fmaf(a[m], b[n], acc[m][n]);
  }
}

template <class T>
__device__ __forceinline__ void ct_zero(float (&acc)[8][T::TN]) {
#pragma unroll
  for (int m = 0; m < 8; ++m)
#pragma unroll
    for (int n = 0; n < T::TN; ++n) acc[m][n] = 0.f;
}

// Double-buffered main loop over `num_kt` K-tiles: the global loads of K-tile t+1 are in flight
// while tile t is computed.  `smem` holds T::TILE_FLOATS floats (2 A stages then 2 B stages).
// Accumulates into acc (not zeroed here).  Ends with a barrier: smem may be reused afterwards.
template <class T, class Loader>
__device__ __forceinline__ void ct_mainloop(Loader& ld, int num_kt, const CtCoord& c, float (&acc)[8][T::TN],
                                            float* smem) {
  if (num_kt <= 0) return;
  float* As = smem;
  float* Bs = smem + 2 * T::STAGE_A;
  ld.fetch();
  ld.commit(As, Bs);
  __syncthreads();
  int cur = 0;
  for (int kt = 0; kt < num_kt; ++kt) {
    const bool more = (kt + 1 < num_kt);
    if (more) ld.fetch();
    ct_compute_stage<T>(As + cur * T::STAGE_A, Bs + cur * T::STAGE_B, c, acc);
    if (more) ld.commit(As + (cur ^ 1) * T::STAGE_A, Bs + (cur ^ 1) * T::STAGE_B);
    __syncthreads();
    cur ^= 1;
  }
}

// A-tile ownership for operands that are contiguous along I: a thread owns NP column(s) and KP
// of the 8 k rows of each.  BM >= 256: columns tid + 256*p, all 8 rows; BM == 128: column
// tid & 127 and rows (tid >> 7) + 2*r.  Element e = p*KP + r, e in [0, EA).
template <class T>
struct CtAMapI {
  static constexpr int NP = (T::BM >= 256) ? T::BM / 256 : 1;
  static constexpr int KP = (T::BM >= 256) ? 8 : 4;
  static constexpr int KSTEP = (T::BM >= 256) ? 1 : 2;
  static_assert(NP * KP == T::EA, "A ownership covers the tile");
  static __device__ __forceinline__ int col(int tid, int p) { return (T::BM >= 256) ? tid + 256 * p : (tid & 127); }
  static __device__ __forceinline__ int row0(int tid) { return (T::BM >= 256) ? 0 : (tid >> 7); }
};
// ... and for operands contiguous along K: 8 consecutive lanes walk k, e strides I by 32.
__device__ __forceinline__ void ct_map_kcontig(int tid, int e, int& kl, int& xl) {
  kl = tid & 7;
  xl = (tid >> 3) + 32 * e;
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
