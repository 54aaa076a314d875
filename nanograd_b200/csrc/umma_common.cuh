// Blackwell (sm_100a) building blocks for the tensor-core contraction kernels: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 MMA / TMEM, UMMA shared-memory and instruction descriptors,
// and host-side tensor-map encoding through the driver entry point (the library links cudart
// statically and never links libcuda directly).
//
// Bit layouts follow the PTX ISA "tcgen05 matrix descriptor / instruction descriptor" tables
// (cross-checked against cute/arch/mma_sm100_desc.hpp of the CUTLASS copy in this image).
#pragma once
#include "ng_common.h"

#ifndef NG_EMU
#include <cuda.h>

namespace ng {
namespace umma {

// ---------------------------------------------------------------- shared-memory addresses
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug (or a faulted TMA) traps instead of hanging the GPU.  With a debug
// buffer (host-mapped, NG_UMMA_DEBUG=1) the timed-out wait records who was waiting on what.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, volatile uint32_t* dbg = nullptr,
                                          uint32_t who = 0) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {
      if (dbg) {
        dbg[0] = 0xDEAD0000u | who;
        dbg[1] = parity;
        dbg[2] = smem_u32(bar);
        __threadfence_system();
      }
      __trap();
    }
  }
}

// ---------------------------------------------------------------- TMA (tiled tensor loads)
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------- thread-block clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Wait on a barrier whose arrivals come from other CTAs of the cluster (multicast tcgen05.commit).
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, volatile uint32_t* dbg = nullptr,
                                                  uint32_t who = 0) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 24)) {
      if (dbg) {
        dbg[0] = 0xDEAD0000u | who;
        dbg[1] = parity;
        dbg[2] = smem_u32(bar);
        __threadfence_system();
      }
      __trap();
    }
  }
}
// TMA box delivered to the same shared-memory offset (and signalling the barrier at the same offset) of every
// CTA of the cluster whose rank bit is set in `mask`.
__device__ __forceinline__ void tma_load_5d_multicast(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0,
                                                      int c1, int c2, int c3, int c4, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_multicast(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0,
                                                      int c1, int c2, int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "h"(mask)
      : "memory");
}

// ---------------------------------------------------------------- TMEM allocation
// Executed by one full warp.  The base address (lane 0, first column) lands in *smem_dst.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One lane of a converged warp (the same one every time for a full mask).  The MMA warp runs its
// loops warp-uniformly and only the instruction issue sits under this predicate, so descriptors
// and addresses stay in uniform registers (a single-lane branch made every operand take the
// R2UR path: ~150 cycles of issue per tcgen05.mma, more than the MMA itself).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- MMA
// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs (the FP32 words in shared memory are
// read with their 13 low mantissa bits ignored), FP32 accumulation.  One thread issues.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand in TMEM (lane = row m, 8 consecutive 32-bit columns = the K = 8 slice),
// which keeps A's shared-memory reads off the MMA critical path (measured: tools/umma_probe3.cu).
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// FP16 inputs (kind::f16, K = 16 per instruction, twice the TF32 rate), FP32 accumulation, both operands
// from shared memory.
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 32 consecutive columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
      "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
      "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mbarrier arrive once every MMA issued so far by this thread has completed (implies
// tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// The same arrive delivered to the barrier at this offset in every CTA of the cluster whose rank bit is in `mask`.
__device__ __forceinline__ void mma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive columns (thread = lane = D row).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (64 bit):
//   [ 0,14) start address >> 4      [16,30) leading-dimension byte offset >> 4
//   [32,46) stride-dimension byte offset >> 4      [46,48) version = 1 (Blackwell)
//   [49,52) base offset = 0         [61,64) swizzle: 0 none, 2 = 128B, 4 = 64B, 6 = 32B
//   1 = 128B rows swizzled in 32B atoms: the only layout for MN-major 32-bit (TF32) operands
enum { SWZ_NONE = 0, SWZ_128B_ATOM32 = 1, SWZ_128B = 2, SWZ_64B = 4, SWZ_32B = 6 };
// High word of a descriptor (constant per operand kind) and low word for a given address: the MMA
// loops advance only the low word.
__host__ __device__ __forceinline__ uint32_t smem_desc_hi(uint32_t sbo_bytes, uint32_t swizzle) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((swizzle & 7u) << 29);
}
__host__ __device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
__host__ __device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                            uint32_t sbo_bytes, uint32_t swizzle) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(swizzle & 7u) << 61;
  return d;
}
// Instruction descriptor (32 bit) for kind::tf32 with FP32 accumulation:
//   [4,6) D format = 1 (F32)   [7,10) A format = 2 (TF32)   [10,13) B format = 2 (TF32)
//   [15] A major (0 = K, 1 = MN)   [16] B major   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// kind::f16 with FP16 inputs (A/B format 0) and FP32 accumulation
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// Round-to-nearest (ties away) FP32 -> TF32.  The tensor core would otherwise truncate the 13 low
// mantissa bits (a biased, twice larger error); the staging copies round once, for free.
__device__ __forceinline__ float tf32_round(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}


// ------------------------------------------------------------------ 3xTF32 operand split
// FP32-accurate mode.  The tensor core reads an FP32 word as TF32 by ignoring its 13 low mantissa
// bits, so the stage as TMA wrote it already IS the "hi" operand: hi = trunc13(v).  The split warps
// only produce  lo = tf32(v - hi)  (exact difference, then rounded) into a separate "lo" operand
// set, and the MMA warp issues  D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (the dropped lo*lo term is
// ~2^-20 relative to one product).  128 threads, conflict-free 16-byte accesses; the generic-proxy
// writes are published to the tensor core's async proxy with fence.proxy.async before the arrive.
// lo = v - trunc13(v): exact in FP32 and two instructions.  The tensor core truncates lo itself when
// it reads it (|error| < 2^-10 |lo| < 2^-20 |v|); rounding it here with cvt.rna would cost ~6 more ALU
// instructions per element and made the split warps, not the tensor pipe, the critical path.
__device__ __forceinline__ float tf32_lo(float v) {
  const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  return v - hi;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// Explicit shared-space accesses in batches of 8 independent 16-byte loads per thread: with generic
// pointers the compiler serialised load -> convert -> store per element (one load in flight).
__device__ __forceinline__ void split_stage_hi_lo(const uint8_t* stage_raw, uint8_t* stage_lo, uint32_t bytes, int t,
                                                  int exp_flags = 0) {
  const uint32_t raw = smem_u32(stage_raw), lo = smem_u32(stage_lo);
  int n4 = (int)(bytes >> 4);
  if (exp_flags & 16) n4 = 0;
  for (int base = t; base < n4; base += 128 * 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (base + 128 * u < n4) v[u] = lds128(raw + (uint32_t)(base + 128 * u) * 16u);
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (base + 128 * u < n4)
        sts128(lo + (uint32_t)(base + 128 * u) * 16u,
               make_float4(tf32_lo(v[u].x), tf32_lo(v[u].y), tf32_lo(v[u].z), tf32_lo(v[u].w)));
  }
  if (!(exp_flags & 8)) fence_proxy_async();
}

// ---------------------------------------------------------------- host: tensor maps
// Encodes a tiled FP32 tensor map.  dims/box are innermost-first; strides_bytes[i] is the byte
// stride of dimension i+1 (dimension 0 is contiguous).  swizzle is one of the SWZ_* codes.
int encode_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle, int fp16 = 0);


// ---------------------------------------------------------------- host: tensor-core convolution
// x3 = 0: TF32 inputs (rtol 1e-2 mode).  x3 = 1: error-compensated 3xTF32 (FP32-accurate).
// Return 0 = done, -1 = shape not supported by the tensor-core path (use the SIMT kernel),
// > 0 = error (ng_last_error is set).
// *_is_nhwc: that operand is already a staged channels-last copy (to_nhwc with the same rounding).
int conv_fprop_tf32(float* y, const float* x, const float* w, const float* bias, int N, int C, int H, int W, int K,
                    int R, int S, int stride, int pad_t, int pad_l, int OH, int OW, int relu, int x3, int x_is_nhwc);
int conv_dgrad_tf32(float* dx, const float* dy, const float* w, int N, int C, int H, int W, int K, int R, int S,
                    int stride, int pad_t, int pad_l, int OH, int OW, int x3, int dy_is_nhwc);
int conv_wgrad_tf32(float* dw, const float* dy, const float* x, int N, int C, int H, int W, int K, int R, int S,
                    int stride, int pad_t, int pad_l, int OH, int OW, int x3, int x_is_nhwc, int dy_is_nhwc);
int to_nhwc(float* out, const float* in, int N, int C, int HW, int round);
// sums[c] = sum_n partial[n][c] in a fixed order (second pass of the staging kernels' channel sums)
int channel_partial_sum(float* sums, const float* partial, long long N, int C);
// FP16x3 (FP32-accurate) engine, umma_conv_f16.cu: operands staged once per tensor as scaled (hi, lo) FP16
// pairs, channels-last; `staged` holds N*C*HW*4 + 64 bytes; sums (optional) receives the per-channel sums of
// `in`.  *_is_staged: that argument already is the to_nhwc_f16 copy.  Same return convention as above.
bool f16x3_supported(int C, int K, int RS);
// known_bits (optional): device word holding the bits of max |in| (or an upper bound) when the producer knows it
int to_nhwc_f16(void* staged, const float* in, float* sums, int N, int C, int HW, const unsigned int* known_bits);
// the same staged copy of a BatchNorm2d (+ ReLU) output / input gradient that was never written (see StageSrc)
int to_nhwc_f16_bn(void* staged, const float* x, const float* gamma, const float* beta, const float* mean,
                   const float* invstd, int act, const unsigned int* absmax_bits, int N, int C, int HW);
int to_nhwc_f16_bnbwd_pool(void* staged, float* chan_sums, const float* dpooled, const float* pooled, const float* x,
                           const float* gamma, const float* beta, const float* mean, const float* invstd,
                           const float* bn_sums, int act, const unsigned int* bound_bits, int N, int C, int H, int W);
int to_nhwc_f16_bnbwd(void* staged, float* chan_sums, const float* dy, const float* x, const float* gamma,
                      const float* beta, const float* mean, const float* invstd, const float* bn_sums, int act,
                      const unsigned int* bound_bits, int N, int C, int HW);
// w_is_staged: `w` is the stage_filters_f16 copy for that pass (fprop: staged_f, dgrad: staged_d)
int conv_fprop_f16(float* y, const void* x, const float* w, const float* bias, int N, int C, int H, int W, int K, int R,
                   int S, int stride, int pad_t, int pad_l, int OH, int OW, int relu, int x_is_staged, int w_is_staged);
int conv_dgrad_f16(float* dx, const void* dy, const float* w, int N, int C, int H, int W, int K, int R, int S,
                   int stride, int pad_t, int pad_l, int OH, int OW, int dy_is_staged, int w_is_staged);
int stage_filters_f16(void* staged_f, void* staged_d, const float* w, int K, int C, int RS);
int conv_wgrad_f16(float* dw, const void* dy, const void* x, int N, int C, int H, int W, int K, int R, int S,
                   int stride, int pad_t, int pad_l, int OH, int OW, int x_is_staged, int dy_is_staged);
// to_nhwc plus sums[c] = sum over (n, p) of in[n][c][p] (the bias gradient when `in` is dy)
int to_nhwc_sum(float* out, const float* in, float* sums, int N, int C, int HW, int round);
// C[M,N] = op(A) @ op(B) (+ bias[N]) (ReLU), row-major FP32 (umma_gemm.cu)
int gemm_tf32(float* c, const float* a, const float* b, const float* bias, long long M, long long N, long long K,
              int trans_a, int trans_b, int relu, int x3);
bool gemm_supported(long long M, long long N, long long K, int trans_a, int trans_b);
bool gather_supported(int KC, int J, int RS);
bool wgrad_supported(int C, int K, int RS);

}  // namespace umma
}  // namespace ng
#endif  // NG_EMU
