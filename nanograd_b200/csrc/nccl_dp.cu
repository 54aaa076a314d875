// Data-parallel exchanges: ncclAllReduce(sum, fp32) over chunks of the flat gradient bucket on a
// dedicated communication stream (launched while the backward pass is still running, fenced against
// the compute stream by events), plus the small in-line exchanges of synchronised BatchNorm.
// New functionality — the reference has no multi-device support (nanograd/device.py:10-11 has
// no device ordinal; no nccl/mpi anywhere).  libnccl.so.2 is resolved at run time with
// dlopen so the core library loads (and every single-GPU path works) without NCCL.
#include "ng_common.h"

#ifndef NG_EMU
#include <dlfcn.h>
#endif

namespace ng {

static int g_world = 1, g_rank = 0;
static int g_sync_bn = 0;

#ifndef NG_EMU
// minimal NCCL ABI (stable across 2.x): opaque comm, 128-byte unique id, enums as ints
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
static const int kNcclFloat32 = 7;  // ncclFloat32
static const int kNcclFloat64 = 8;  // ncclFloat64
static const int kNcclSum = 0;      // ncclSum
static const int kNcclMax = 2;      // ncclMax

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
// Two communicators over the same ranks: g_comm carries the gradient bucket on the communication
// stream (overlapping the backward tail); g_comm_inline carries the small in-line exchanges of the
// compute stream (SyncBN statistics, host scalars), so a statistics exchange never queues behind a
// bucket chunk that is still waiting for its producers.
static ncclComm_t g_comm = nullptr, g_comm_inline = nullptr;
static cudaStream_t g_comm_stream = nullptr;
static cudaEvent_t g_ev_ready = nullptr, g_ev_done = nullptr;
static double* g_scalar_dev = nullptr;

static int load_nccl() {
  if (g_nccl.handle) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  // prefer an instance that is already mapped into the process (e.g. the one torch ships)
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
    if (h) break;
  }
  const char* env = getenv("NANOGRAD_B200_NCCL");
  if (!h && env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  for (const char* n : names) {
    if (h) break;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!h) return set_error("NCCL is required for multi-GPU runs but libnccl.so.2 could not be loaded: %s", dlerror());
  g_nccl.handle = h;
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.CommDestroy) {
    g_nccl.handle = nullptr;
    return set_error("libnccl is missing required symbols");
  }
  return 0;
}

#define NG_CHECK_NCCL(expr)                                                                      \
  do {                                                                                           \
    ncclResult_t _r = (expr);                                                                    \
    if (_r != 0)                                                                                 \
      return set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error"); \
  } while (0)

int dp_allreduce_inline(float* buf, int64_t count) {
  NG_REQUIRE(g_comm_inline != nullptr, "data-parallel exchange requested but the communicator is not initialised");
  NG_CHECK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat32, kNcclSum, g_comm_inline, g_stream));
  g_launches++;
  return 0;
}

int dp_allgather_inline(const float* send, float* recv, int64_t count_per_rank) {
  NG_REQUIRE(g_comm_inline != nullptr, "data-parallel exchange requested but the communicator is not initialised");
  NG_CHECK_NCCL(g_nccl.AllGather(send, recv, (size_t)count_per_rank, kNcclFloat32, g_comm_inline, g_stream));
  g_launches++;
  return 0;
}
#else
// Host emulation (tests/emu): the collectives call back into the test process, which runs them over
// torch.distributed's gloo backend on the host arrays that stand in for device memory.
//   kind 0: in-place allreduce(sum) of `count` floats at `send`
//   kind 1: allgather of `count` floats per rank from `send` into `recv`
//   kind 2 / 3: in-place allreduce sum / max of `count` doubles at `send`
typedef int (*ng_emu_collective_fn)(int kind, void* send, void* recv, int64_t count);
static ng_emu_collective_fn g_emu_collective = nullptr;

static int emu_collective(int kind, void* send, void* recv, int64_t count) {
  NG_REQUIRE(g_emu_collective != nullptr, "data-parallel exchange requested but no collective is registered");
  if (g_emu_collective(kind, send, recv, count)) return set_error("emulated collective %d failed", kind);
  return 0;
}
int dp_allreduce_inline(float* buf, int64_t count) { return emu_collective(0, buf, nullptr, count); }
int dp_allgather_inline(const float* send, float* recv, int64_t count_per_rank) {
  return emu_collective(1, (void*)send, recv, count_per_rank);
}
#endif

int dp_world() { return g_world; }
int dp_rank() { return g_rank; }
bool dp_sync_bn() { return g_sync_bn != 0 && g_world > 1; }

}  // namespace ng

using namespace ng;

extern "C" {

int ng_nccl_sync_bn(int on) {
  g_sync_bn = on ? 1 : 0;
  return 0;
}

int ng_nccl_world(int* world, int* rank) {
  if (world) *world = g_world;
  if (rank) *rank = g_rank;
  return 0;
}

#ifdef NG_EMU
int ng_emu_set_collective(void* fn, int world_size, int rank) {
  g_emu_collective = (ng_emu_collective_fn)fn;
  g_world = fn ? world_size : 1;
  g_rank = fn ? rank : 0;
  return 0;
}
int ng_nccl_unique_id(void*) { return set_error("NCCL is not available in the emulation build"); }
int ng_nccl_init(const void*, int, int) { return set_error("NCCL is not available in the emulation build"); }
int ng_nccl_allreduce_sum(uint64_t buf, int64_t count) { return count > 0 ? emu_collective(0, dptr<void>(buf), nullptr, count) : 0; }
int ng_nccl_allreduce_begin(uint64_t buf, int64_t count) { return ng_nccl_allreduce_sum(buf, count); }
int ng_nccl_allreduce_end(void) { return 0; }
int ng_nccl_host_allreduce(double* values, int count, int op) { return emu_collective(op == 1 ? 3 : 2, values, nullptr, count); }
int ng_nccl_destroy(void) {
  g_emu_collective = nullptr;
  g_world = 1;
  g_rank = 0;
  return 0;
}
#else

int ng_nccl_unique_id(void* out_256_bytes) {
  // two ids back to back: the bucket communicator and the in-line communicator
  if (load_nccl()) return 1;
  ncclUniqueId id;
  for (int i = 0; i < 2; ++i) {
    NG_CHECK_NCCL(g_nccl.GetUniqueId(&id));
    memcpy((char*)out_256_bytes + i * sizeof(id), &id, sizeof(id));
  }
  return 0;
}

int ng_nccl_init(const void* id_256_bytes, int world_size, int rank) {
  NG_ENSURE_INIT();
  if (load_nccl()) return 1;
  NG_REQUIRE(g_comm == nullptr, "ng_nccl_init: communicator already initialised");
  NG_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "ng_nccl_init: bad rank %d / world %d", rank, world_size);
  ncclUniqueId id;
  NG_CHECK_CUDA(cudaSetDevice(g_device));
  memcpy(&id, id_256_bytes, sizeof(id));
  NG_CHECK_NCCL(g_nccl.CommInitRank(&g_comm, world_size, id, rank));
  memcpy(&id, (const char*)id_256_bytes + sizeof(id), sizeof(id));
  NG_CHECK_NCCL(g_nccl.CommInitRank(&g_comm_inline, world_size, id, rank));
  NG_CHECK_CUDA(cudaStreamCreateWithFlags(&g_comm_stream, cudaStreamNonBlocking));
  NG_CHECK_CUDA(cudaEventCreateWithFlags(&g_ev_ready, cudaEventDisableTiming));
  NG_CHECK_CUDA(cudaEventCreateWithFlags(&g_ev_done, cudaEventDisableTiming));
  NG_CHECK_CUDA(cudaMalloc((void**)&g_scalar_dev, 64 * sizeof(double)));
  g_world = world_size;
  g_rank = rank;
  return 0;
}

// Launch the allreduce(sum) of buf[0..count) on the communication stream, ordered after everything the
// compute stream has been given so far (the kernels that produced `buf`).  Several chunks may be in
// flight; the compute stream does not wait until ng_nccl_allreduce_end().
int ng_nccl_allreduce_begin(uint64_t buf, int64_t count) {
  NG_ENSURE_INIT();
  NG_REQUIRE(g_comm != nullptr, "ng_nccl_allreduce_begin: communicator not initialised (call ng_nccl_init)");
  if (count <= 0) return 0;
  NG_CHECK_CUDA(cudaEventRecord(g_ev_ready, g_stream));
  NG_CHECK_CUDA(cudaStreamWaitEvent(g_comm_stream, g_ev_ready, 0));
  NG_CHECK_NCCL(g_nccl.AllReduce(dptr<const void>(buf), dptr<void>(buf), (size_t)count, kNcclFloat32, kNcclSum, g_comm,
                                 g_comm_stream));
  g_launches++;
  return 0;
}

// The compute stream waits for every chunk launched so far (before the optimizer kernel).
int ng_nccl_allreduce_end(void) {
  NG_ENSURE_INIT();
  if (!g_comm) return 0;
  NG_CHECK_CUDA(cudaEventRecord(g_ev_done, g_comm_stream));
  NG_CHECK_CUDA(cudaStreamWaitEvent(g_stream, g_ev_done, 0));
  return 0;
}

int ng_nccl_allreduce_sum(uint64_t buf, int64_t count) {
  if (ng_nccl_allreduce_begin(buf, count)) return 1;
  return ng_nccl_allreduce_end();
}

// Host scalars across ranks (barriers, max-over-ranks of device-timed durations, checksums): op 0 = sum,
// 1 = max, on up to 64 doubles.  Blocking.
int ng_nccl_host_allreduce(double* values, int count, int op) {
  NG_ENSURE_INIT();
  NG_REQUIRE(count >= 0 && count <= 64, "ng_nccl_host_allreduce: at most 64 values");
  if (g_world == 1 || count == 0) return 0;
  NG_REQUIRE(g_comm_inline != nullptr, "ng_nccl_host_allreduce: communicator not initialised");
  NG_CHECK_CUDA(cudaMemcpyAsync(g_scalar_dev, values, count * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  NG_CHECK_NCCL(g_nccl.AllReduce(g_scalar_dev, g_scalar_dev, (size_t)count, kNcclFloat64, op == 1 ? kNcclMax : kNcclSum,
                                 g_comm_inline, g_stream));
  NG_CHECK_CUDA(cudaMemcpyAsync(values, g_scalar_dev, count * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  NG_CHECK_CUDA(cudaStreamSynchronize(g_stream));
  return 0;
}

int ng_nccl_destroy(void) {
  if (g_comm) {
    cudaStreamSynchronize(g_comm_stream);
    cudaStreamSynchronize(g_stream);
    g_nccl.CommDestroy(g_comm);
    g_nccl.CommDestroy(g_comm_inline);
    g_comm = g_comm_inline = nullptr;
    g_world = 1;
    g_rank = 0;
  }
  return 0;
}
#endif

}  // extern "C"
