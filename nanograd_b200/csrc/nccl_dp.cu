// Data-parallel gradient exchange: one ncclAllReduce(sum, fp32) over the flat gradient bucket
// per step, on a dedicated communication stream fenced against the compute stream by events.
// New functionality — the reference has no multi-device support (nanograd/device.py:10-11 has
// no device ordinal; no nccl/mpi anywhere).  libnccl.so.2 is resolved at run time with
// dlopen so the core library loads (and every single-GPU path works) without NCCL.
#include "ng_common.h"

#ifndef NG_EMU
#include <dlfcn.h>
#endif

namespace ng {

#ifndef NG_EMU
// minimal NCCL ABI (stable across 2.x): opaque comm, 128-byte unique id, enums as ints
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
static const int kNcclFloat32 = 7;  // ncclFloat32
static const int kNcclSum = 0;      // ncclSum

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
static cudaStream_t g_comm_stream = nullptr;
static cudaEvent_t g_ev_ready = nullptr, g_ev_done = nullptr;
static int g_world = 1, g_rank = 0;

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
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
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
#endif

}  // namespace ng

using namespace ng;

extern "C" {

#ifdef NG_EMU
int ng_nccl_unique_id(void*) { return set_error("NCCL is not available in the emulation build"); }
int ng_nccl_init(const void*, int, int) { return set_error("NCCL is not available in the emulation build"); }
int ng_nccl_allreduce_sum(uint64_t, int64_t) { return set_error("NCCL is not available in the emulation build"); }
int ng_nccl_destroy(void) { return 0; }
#else

int ng_nccl_unique_id(void* out_128_bytes) {
  if (load_nccl()) return 1;
  ncclUniqueId id;
  NG_CHECK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out_128_bytes, &id, sizeof(id));
  return 0;
}

int ng_nccl_init(const void* id_128_bytes, int world_size, int rank) {
  NG_ENSURE_INIT();
  if (load_nccl()) return 1;
  NG_REQUIRE(g_comm == nullptr, "ng_nccl_init: communicator already initialised");
  NG_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "ng_nccl_init: bad rank %d / world %d", rank, world_size);
  ncclUniqueId id;
  memcpy(&id, id_128_bytes, sizeof(id));
  NG_CHECK_CUDA(cudaSetDevice(g_device));
  NG_CHECK_NCCL(g_nccl.CommInitRank(&g_comm, world_size, id, rank));
  NG_CHECK_CUDA(cudaStreamCreateWithFlags(&g_comm_stream, cudaStreamNonBlocking));
  NG_CHECK_CUDA(cudaEventCreateWithFlags(&g_ev_ready, cudaEventDisableTiming));
  NG_CHECK_CUDA(cudaEventCreateWithFlags(&g_ev_done, cudaEventDisableTiming));
  g_world = world_size;
  g_rank = rank;
  return 0;
}

int ng_nccl_allreduce_sum(uint64_t buf, int64_t count) {
  NG_ENSURE_INIT();
  NG_REQUIRE(g_comm != nullptr, "ng_nccl_allreduce_sum: communicator not initialised (call ng_nccl_init)");
  if (count <= 0) return 0;
  // comm stream waits for the producers of `buf` on the compute stream ...
  NG_CHECK_CUDA(cudaEventRecord(g_ev_ready, g_stream));
  NG_CHECK_CUDA(cudaStreamWaitEvent(g_comm_stream, g_ev_ready, 0));
  NG_CHECK_NCCL(g_nccl.AllReduce(dptr<const void>(buf), dptr<void>(buf), (size_t)count, kNcclFloat32, kNcclSum, g_comm,
                                 g_comm_stream));
  // ... and the compute stream waits for the reduced values before the optimizer kernel
  NG_CHECK_CUDA(cudaEventRecord(g_ev_done, g_comm_stream));
  NG_CHECK_CUDA(cudaStreamWaitEvent(g_stream, g_ev_done, 0));
  return 0;
}

int ng_nccl_destroy(void) {
  if (g_comm) {
    cudaStreamSynchronize(g_comm_stream);
    g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
  }
  return 0;
}
#endif

}  // extern "C"
