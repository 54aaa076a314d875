// Runtime: device selection, the single compute stream, a stream-ordered caching allocator,
// host<->device copies, CUDA-event timers and two measurement helpers.
//
// Replaces the PyOpenCL context/queue bootstrap and buffer management of the reference
// (nanograd/tensor.py:14-30,153-182; nanograd/nn/buffer.py:14-26).
#include "ng_common.h"
#include <stdlib.h>

#include <chrono>
#include <condition_variable>
#include <string>
#include <thread>
#include <vector>

#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace ng {

cudaStream_t g_stream = nullptr;
int g_device = -1;
int g_sm_count = 0;
uint64_t g_launches = 0;

static thread_local char t_err[1024] = "";
static char g_err[1024] = "";

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
  memcpy(g_err, t_err, sizeof(g_err));
  return 1;
}

int ensure_init() {
  if (g_device < 0) return ng_init(0);
  return 0;
}

// ---- caching allocator -------------------------------------------------------------------------
// All kernels run on one in-order stream, so a block returned to the pool can be handed to
// the next request immediately: any kernel still reading it is ordered before any kernel the
// new owner will launch.  Training repeats the same shapes every step → after the first step
// no cudaMalloc is issued.
struct Block { uint64_t size; uint64_t owner; };  // owner 0 = the global pool, else the id of a captured graph
struct Pool {
  std::mutex mu;
  std::map<uint64_t, std::vector<void*>> free_by_size;
  std::unordered_map<void*, Block> size_of;  // every block we ever allocated and still own
  uint64_t live = 0, pooled = 0, n_allocs = 0;
};
static Pool& pool() {
  static Pool* p = new Pool();  // leaked on purpose: outlive Python's teardown order
  return *p;
}

// ---- captured steps (CUDA graphs) ----------------------------------------------------------------
// A training step whose cost is host dispatch (the README MNIST CNN: ~30 launches of microseconds each)
// is captured once into a CUDA graph and replayed with one cudaGraphLaunch.  The graph bakes device
// addresses in, so everything allocated while capturing comes from — and returns to — a pool private
// to that graph: replays can never collide with memory the regular pool hands out later.
struct GraphRec {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  std::map<uint64_t, std::vector<void*>> free_by_size;  // private free list (reused within the capture)
  uint64_t kernel_nodes = 0, bytes = 0;
};
static std::unordered_map<uint64_t, GraphRec*>* g_graphs = nullptr;
static uint64_t g_capturing = 0;  // id of the graph being captured, 0 = none
static uint64_t g_next_graph = 1;

static uint64_t round_size(uint64_t b) {
  if (b == 0) b = 1;
#ifdef NG_EMU_EXACT_ALLOC
  return b;   // AddressSanitizer build of the host emulation (tools/kernel_memcheck.py): no slack behind a tensor
#else
  const uint64_t q = (b < (1u << 20)) ? 512 : (1u << 16);
  return (b + q - 1) / q * q;
#endif
}

static void release_pooled_locked(Pool& P) {
  for (auto& kv : P.free_by_size) {
    for (void* p : kv.second) {
      cudaFree(p);
      P.size_of.erase(p);
    }
  }
  P.free_by_size.clear();
  P.pooled = 0;
}

int pool_alloc(void** out, uint64_t bytes) {
  Pool& P = pool();
  const uint64_t sz = round_size(bytes);
  std::lock_guard<std::mutex> lk(P.mu);
  GraphRec* rec = g_capturing ? (*g_graphs)[g_capturing] : nullptr;
  auto& free_list = rec ? rec->free_by_size : P.free_by_size;
  auto it = free_list.find(sz);
  if (it != free_list.end() && !it->second.empty()) {
    *out = it->second.back();
    it->second.pop_back();
    if (!rec) P.pooled -= sz;
    P.live += sz;
#ifdef NG_EMU
    ng_emu_poison_nan(*out, sz);   // host emulation only (tools/kernel_initcheck.py): recycled blocks look unwritten
#endif
    return 0;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, sz);
  if (e != cudaSuccess && !rec) {
    cudaGetLastError();
    cudaStreamSynchronize(g_stream);
    release_pooled_locked(P);
    e = cudaMalloc(&p, sz);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error("device out of memory: %llu bytes requested (%llu live, pool emptied): %s",
                     (unsigned long long)sz, (unsigned long long)P.live, cudaGetErrorString(e));
  }
  P.n_allocs++;
  P.size_of[p] = Block{sz, g_capturing};
  if (rec) rec->bytes += sz;
  P.live += sz;
#ifdef NG_EMU
  ng_emu_poison_nan(p, sz);
#endif
  *out = p;
  return 0;
}

int pool_free(void* p) {
  if (!p) return 0;
  Pool& P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  auto it = P.size_of.find(p);
  if (it == P.size_of.end()) return set_error("ng_free: pointer %p is not owned by the pool", p);
  const uint64_t sz = it->second.size;
  P.live -= sz;
  if (it->second.owner) {
    auto g = g_graphs ? g_graphs->find(it->second.owner) : decltype(g_graphs->end()){};
    if (g_graphs && g != g_graphs->end()) {
      g->second->free_by_size[sz].push_back(p);  // stays with its graph: replays still use the address
      return 0;
    }
    it->second.owner = 0;  // the graph is gone: the block joins the global pool
  }
  P.free_by_size[sz].push_back(p);
  P.pooled += sz;
  return 0;
}

// ---- FFMA peak microkernel -----------------------------------------------------------------------
__global__ void ffma_peak_kernel(float* out, int iters, float a, float b) {
  float r0 = threadIdx.x, r1 = r0 + 1.f, r2 = r0 + 2.f, r3 = r0 + 3.f;
  float r4 = r0 + 4.f, r5 = r0 + 5.f, r6 = r0 + 6.f, r7 = r0 + 7.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      r0 = fmaf(r0, a, b); r1 = fmaf(r1, a, b); r2 = fmaf(r2, a, b); r3 = fmaf(r3, a, b);
      r4 = fmaf(r4, a, b); r5 = fmaf(r5, a, b); r6 = fmaf(r6, a, b); r7 = fmaf(r7, a, b);
    }
  }
  const float s = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
  if (s == 123.456f) out[0] = s;  // keeps the chain alive, practically never true
}

}  // namespace ng

using namespace ng;

extern "C" {

int ng_init(int device_ordinal) {
  if (g_device >= 0) {
    NG_REQUIRE(g_device == device_ordinal || device_ordinal < 0,
               "ng_init: already initialised on device %d (requested %d)", g_device, device_ordinal);
    return 0;
  }
  if (device_ordinal < 0) device_ordinal = 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return set_error("ng_init: no CUDA device available (%s); nanograd_b200 has no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  NG_REQUIRE(device_ordinal < n, "ng_init: device %d out of range (%d devices)", device_ordinal, n);
  NG_CHECK_CUDA(cudaSetDevice(device_ordinal));
  cudaDeviceProp prop;
  NG_CHECK_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
  g_sm_count = prop.multiProcessorCount;
  {
    // NANOGRAD_B200_SM_RESERVE=k: size every grid for k fewer SMs.  The persistent tensor-core kernels run one CTA per
    // SM; under data parallelism an SM held by an NCCL kernel makes one of their CTAs wait for a whole persistent
    // loop, so nanograd_b200.dist reserves as many SMs as it lets NCCL use.
    const char* e = getenv("NANOGRAD_B200_SM_RESERVE");
    const int k = e ? atoi(e) : 0;
    if (k > 0 && k < g_sm_count / 2) g_sm_count -= k;
  }
  NG_CHECK_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  g_device = device_ordinal;
  return 0;
}

const char* ng_last_error(void) { return g_err; }

int ng_device_info(int* sm_count, int* cc_major, int* cc_minor, uint64_t* total_mem_bytes) {
  NG_ENSURE_INIT();
  cudaDeviceProp prop;
  NG_CHECK_CUDA(cudaGetDeviceProperties(&prop, g_device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (total_mem_bytes) *total_mem_bytes = (uint64_t)prop.totalGlobalMem;
  return 0;
}

int ng_sync(void) {
  NG_ENSURE_INIT();
  NG_CHECK_CUDA(cudaStreamSynchronize(g_stream));
  return 0;
}

int ng_launch_count(uint64_t* out) {
  *out = g_launches;
  return 0;
}

int ng_malloc(uint64_t* out_ptr, uint64_t bytes) {
  NG_ENSURE_INIT();
  void* p = nullptr;
  int r = pool_alloc(&p, bytes);
  if (r) return r;
  *out_ptr = (uint64_t)(uintptr_t)p;
  return 0;
}

int ng_free(uint64_t ptr) { return pool_free(dptr<void>(ptr)); }

#ifndef NG_EMU
int ng_graph_begin(void) {
  NG_ENSURE_INIT();
  NG_REQUIRE(g_capturing == 0, "ng_graph_begin: a capture is already in progress");
  if (!g_graphs) g_graphs = new std::unordered_map<uint64_t, GraphRec*>();
  const uint64_t id = g_next_graph++;
  {
    std::lock_guard<std::mutex> lk(pool().mu);
    (*g_graphs)[id] = new GraphRec();
    g_capturing = id;
  }
  // relaxed mode: the private pool may call cudaMalloc while capturing
  cudaError_t e = cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) {
    std::lock_guard<std::mutex> lk(pool().mu);
    g_capturing = 0;
    delete (*g_graphs)[id];
    g_graphs->erase(id);
    return set_error("cudaStreamBeginCapture failed: %s", cudaGetErrorString(e));
  }
  return 0;
}

int ng_graph_end(uint64_t* out_handle, uint64_t* out_kernel_nodes, uint64_t* out_private_bytes) {
  NG_REQUIRE(g_capturing != 0, "ng_graph_end: no capture in progress");
  const uint64_t id = g_capturing;
  GraphRec* rec = (*g_graphs)[id];
  cudaError_t e = cudaStreamEndCapture(g_stream, &rec->graph);
  {
    std::lock_guard<std::mutex> lk(pool().mu);
    g_capturing = 0;
  }
  if (e != cudaSuccess || !rec->graph) {
    cudaGetLastError();
    return set_error("cudaStreamEndCapture failed: %s (a captured step must not synchronise, copy from pageable "
                     "host memory or read results back)", cudaGetErrorString(e));
  }
  size_t n = 0;
  NG_CHECK_CUDA(cudaGraphGetNodes(rec->graph, nullptr, &n));
  std::vector<cudaGraphNode_t> nodes(n);
  if (n) NG_CHECK_CUDA(cudaGraphGetNodes(rec->graph, nodes.data(), &n));
  for (size_t i = 0; i < n; ++i) {
    cudaGraphNodeType t;
    if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) rec->kernel_nodes++;
  }
  NG_CHECK_CUDA(cudaGraphInstantiate(&rec->exec, rec->graph, 0));
  *out_handle = id;
  if (out_kernel_nodes) *out_kernel_nodes = rec->kernel_nodes;
  if (out_private_bytes) *out_private_bytes = rec->bytes;
  return 0;
}

int ng_graph_launch(uint64_t handle) {
  NG_REQUIRE(g_graphs && g_graphs->count(handle), "ng_graph_launch: unknown graph %llu", (unsigned long long)handle);
  NG_REQUIRE(g_capturing == 0, "ng_graph_launch: cannot replay while capturing");
  GraphRec* rec = (*g_graphs)[handle];
  NG_CHECK_CUDA(cudaGraphLaunch(rec->exec, g_stream));
  g_launches += rec->kernel_nodes;
  return 0;
}

int ng_graph_destroy(uint64_t handle) {
  if (!g_graphs || !g_graphs->count(handle)) return 0;
  NG_CHECK_CUDA(cudaStreamSynchronize(g_stream));
  GraphRec* rec = (*g_graphs)[handle];
  if (rec->exec) cudaGraphExecDestroy(rec->exec);
  if (rec->graph) cudaGraphDestroy(rec->graph);
  Pool& P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  for (auto& kv : rec->free_by_size)
    for (void* p : kv.second) {
      cudaFree(p);
      P.size_of.erase(p);
    }
  // blocks of this graph that are still held by live tensors join the global pool when they are freed
  g_graphs->erase(handle);
  delete rec;
  return 0;
}

int ng_graph_capturing(int* out) {
  *out = g_capturing != 0;
  return 0;
}
#else
int ng_graph_begin(void) { return set_error("CUDA graphs are not available in the emulation build"); }
int ng_graph_end(uint64_t*, uint64_t*, uint64_t*) { return set_error("CUDA graphs are not available in the emulation build"); }
int ng_graph_launch(uint64_t) { return set_error("CUDA graphs are not available in the emulation build"); }
int ng_graph_destroy(uint64_t) { return 0; }
int ng_graph_capturing(int* out) { *out = 0; return 0; }
#endif

int ng_mem_stats(uint64_t* live_bytes, uint64_t* pooled_bytes, uint64_t* n_device_allocs) {
  Pool& P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  if (live_bytes) *live_bytes = P.live;
  if (pooled_bytes) *pooled_bytes = P.pooled;
  if (n_device_allocs) *n_device_allocs = P.n_allocs;
  return 0;
}

int ng_pool_release(void) {
  NG_ENSURE_INIT();
  NG_CHECK_CUDA(cudaStreamSynchronize(g_stream));
  Pool& P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  release_pooled_locked(P);
  return 0;
}

int ng_memset_zero(uint64_t ptr, uint64_t bytes) {
  NG_ENSURE_INIT();
  if (bytes == 0) return 0;
  NG_CHECK_CUDA(cudaMemsetAsync(dptr<void>(ptr), 0, bytes, g_stream));
  return 0;
}

int ng_h2d(uint64_t dst, const void* src, uint64_t bytes) {
  NG_ENSURE_INIT();
  if (bytes == 0) return 0;
  // Pageable sources are staged by the driver before the call returns; pinned sources
  // (ng_pinned_alloc) are copied asynchronously on the compute stream.
  NG_CHECK_CUDA(cudaMemcpyAsync(dptr<void>(dst), src, bytes, cudaMemcpyHostToDevice, g_stream));
  return 0;
}

int ng_d2h(void* dst, uint64_t src, uint64_t bytes) {
  NG_ENSURE_INIT();
  if (bytes) NG_CHECK_CUDA(cudaMemcpyAsync(dst, dptr<void>(src), bytes, cudaMemcpyDeviceToHost, g_stream));
  NG_CHECK_CUDA(cudaStreamSynchronize(g_stream));
  return 0;
}

// Device-to-host read of a result whose producer was enqueued before `ready_event` (an ng_event_* handle recorded
// right after it): the copy runs on a separate read-back stream that waits for that event only, so reading a
// step's loss does not drain the backward pass and the optimizer step already queued behind it.
static cudaStream_t g_read_stream = nullptr;

int ng_d2h_after(void* dst, uint64_t src, uint64_t bytes, uint64_t ready_event) {
  NG_ENSURE_INIT();
  NG_REQUIRE(ready_event != 0, "ng_d2h_after: null event");
  if (!g_read_stream) NG_CHECK_CUDA(cudaStreamCreateWithFlags(&g_read_stream, cudaStreamNonBlocking));
  NG_CHECK_CUDA(cudaStreamWaitEvent(g_read_stream, reinterpret_cast<cudaEvent_t>(static_cast<uintptr_t>(ready_event)), 0));
  if (bytes) NG_CHECK_CUDA(cudaMemcpyAsync(dst, dptr<void>(src), bytes, cudaMemcpyDeviceToHost, g_read_stream));
  NG_CHECK_CUDA(cudaStreamSynchronize(g_read_stream));
  return 0;
}

int ng_d2d(uint64_t dst, uint64_t src, uint64_t bytes) {
  NG_ENSURE_INIT();
  if (bytes == 0) return 0;
  NG_CHECK_CUDA(cudaMemcpyAsync(dptr<void>(dst), dptr<void>(src), bytes, cudaMemcpyDeviceToDevice, g_stream));
  return 0;
}

// ---- input pipeline: host->device prefetch on a copy stream ----------------------------------------
// The reference's loop builds every batch on the host and copies it synchronously before the step
// (tests/helpers.py:296-299, README.md:107-109).  Here batch i+1 is copied on a second stream while
// step i computes: ng_h2d_prefetch(slot, ...) first records what the compute stream has been given
// so far (the last reader of a double-buffered destination is step i-1) and makes the copy wait for
// that, then copies and records the slot's "done" event; ng_prefetch_wait(slot) makes the compute
// stream wait for it before the first kernel that reads the batch.
static cudaStream_t g_copy_stream = nullptr;
static cudaEvent_t g_consumed = nullptr;
static cudaEvent_t g_copy_done[4] = {nullptr, nullptr, nullptr, nullptr};

int ng_h2d_prefetch(int slot, uint64_t dst, const void* pinned_src, uint64_t bytes) {
  NG_ENSURE_INIT();
  NG_REQUIRE(slot >= 0 && slot < 4, "ng_h2d_prefetch: slot must be in [0, 4)");
  if (!g_copy_stream) {
    NG_CHECK_CUDA(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
    NG_CHECK_CUDA(cudaEventCreateWithFlags(&g_consumed, cudaEventDisableTiming));
  }
  if (!g_copy_done[slot]) NG_CHECK_CUDA(cudaEventCreateWithFlags(&g_copy_done[slot], cudaEventDisableTiming));
  NG_CHECK_CUDA(cudaEventRecord(g_consumed, g_stream));
  NG_CHECK_CUDA(cudaStreamWaitEvent(g_copy_stream, g_consumed, 0));
  if (bytes) NG_CHECK_CUDA(cudaMemcpyAsync(dptr<void>(dst), pinned_src, bytes, cudaMemcpyHostToDevice, g_copy_stream));
  NG_CHECK_CUDA(cudaEventRecord(g_copy_done[slot], g_copy_stream));
  return 0;
}

int ng_prefetch_wait(int slot) {
  NG_ENSURE_INIT();
  NG_REQUIRE(slot >= 0 && slot < 4 && g_copy_done[slot], "ng_prefetch_wait: nothing was prefetched into slot %d", slot);
  NG_CHECK_CUDA(cudaStreamWaitEvent(g_stream, g_copy_done[slot], 0));
  return 0;
}

// dst[i][:] = src[idx[i]][:] on the host (rows of row_bytes bytes).  Called through ctypes, i.e.
// without Python's GIL: the loader's worker thread gathers the next batch into pinned memory
// while the training thread keeps enqueuing kernels.
int ng_host_gather_rows(void* dst, const void* src, const int64_t* idx, int64_t n, uint64_t row_bytes,
                        int64_t n_src_rows) {
  for (int64_t i = 0; i < n; ++i) {
    if (idx[i] < 0 || idx[i] >= n_src_rows) return set_error("ng_host_gather_rows: index %lld out of range", (long long)idx[i]);
    memcpy((char*)dst + (uint64_t)i * row_bytes, (const char*)src + (uint64_t)idx[i] * row_bytes, row_bytes);
  }
  return 0;
}

// host-side wait for the slot's copy: the pinned source may be refilled afterwards
int ng_prefetch_sync(int slot) {
  NG_REQUIRE(slot >= 0 && slot < 4, "ng_prefetch_sync: slot must be in [0, 4)");
  if (g_copy_done[slot]) NG_CHECK_CUDA(cudaEventSynchronize(g_copy_done[slot]));
  return 0;
}

// ---- batch loader: a library-owned host thread gathers and prefetches batches ------------------------
// Python threads would trade the GIL with the training thread on every device call (measured: 0.3-0.7 ms
// per step); this thread never touches the interpreter.  Per step t it waits until batch t-2 was handed
// to the caller (its slot is free) and that slot's previous DMA is over, gathers rows plan[t] of the
// host dataset into pinned staging memory, enqueues the copy on the copy stream (ordered after the
// compute work enqueued so far, i.e. after the last reader of that device slot) and publishes it.
// pinned staging buffers are recycled across loaders (cudaHostAlloc / cudaFreeHost cost milliseconds)
static std::mutex g_pin_mu;
static std::vector<std::pair<uint64_t, char*>> g_pin_free;
static char* pinned_get(uint64_t bytes) {
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (size_t i = 0; i < g_pin_free.size(); ++i)
      if (g_pin_free[i].first == bytes) {
        char* p = g_pin_free[i].second;
        g_pin_free.erase(g_pin_free.begin() + i);
        return p;
      }
  }
  char* p = nullptr;
  if (cudaHostAlloc((void**)&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
static void pinned_put(char* p, uint64_t bytes) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pin_mu);
  if (g_pin_free.size() < 16) g_pin_free.emplace_back(bytes, p);
  else cudaFreeHost(p);
}

struct NgLoader {
  const char* X = nullptr;
  const char* Y = nullptr;
  int64_t n_rows = 0;
  uint64_t row_bytes = 0;
  int batch = 0, steps = 0;
  std::vector<int64_t> plan;    // concatenated index lists
  std::vector<int64_t> offset;  // steps + 1 offsets into plan
  char* x_pin[2] = {nullptr, nullptr};
  char* y_pin[2] = {nullptr, nullptr};
  uint64_t x_dev[2] = {0, 0}, y_dev[2] = {0, 0};
  cudaEvent_t done[2] = {nullptr, nullptr};
  cudaEvent_t consumed_ev = nullptr;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  int produced = 0;   // batches published
  int taken = 0;      // batches handed to the caller
  bool stop = false;
  std::string err;
};

static void loader_main(NgLoader* L) {
  cudaSetDevice(g_device);
  const bool dbg = getenv("NG_LOADER_DEBUG") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (int t = 0; t < L->steps; ++t) {
    const double t_a = now();
    {
      std::unique_lock<std::mutex> lk(L->mu);
      // slot t&1 was last used by batch t-2, whose loop body is over once the caller asked for batch t-1
      // (taken >= t): only then does the compute stream hold every kernel that reads it
      L->cv.wait(lk, [&] { return L->stop || t < 2 || t <= L->taken; });
      if (L->stop) return;
    }
    const int k = t & 1;
    const double t_b = now();
    cudaError_t e = cudaEventSynchronize(L->done[k]);   // pinned slot no longer read by its previous DMA
    const double t_c = now();
    const int64_t* idx = L->plan.data() + L->offset[t];
    const int64_t n = L->offset[t + 1] - L->offset[t];
    for (int64_t i = 0; i < n && e == cudaSuccess; ++i) {
      memcpy(L->x_pin[k] + (uint64_t)i * L->row_bytes, L->X + (uint64_t)idx[i] * L->row_bytes, L->row_bytes);
      if (L->Y) memcpy(L->y_pin[k] + i * 4, L->Y + idx[i] * 4, 4);
    }
    const double t_d = now();
    if (e == cudaSuccess) e = cudaEventRecord(L->consumed_ev, g_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g_copy_stream, L->consumed_ev, 0);
    if (e == cudaSuccess && n)
      e = cudaMemcpyAsync(dptr<void>(L->x_dev[k]), L->x_pin[k], (uint64_t)n * L->row_bytes, cudaMemcpyHostToDevice, g_copy_stream);
    if (e == cudaSuccess && n && L->Y)
      e = cudaMemcpyAsync(dptr<void>(L->y_dev[k]), L->y_pin[k], (uint64_t)n * 4, cudaMemcpyHostToDevice, g_copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(L->done[k], g_copy_stream);
    if (dbg) fprintf(stderr, "[loader] t=%d gate %.3f evsync %.3f gather %.3f enqueue %.3f ms\n", t, t_b - t_a, t_c - t_b,
                     t_d - t_c, now() - t_d);
    {
      std::lock_guard<std::mutex> lk(L->mu);
      if (e != cudaSuccess) { L->err = cudaGetErrorString(e); L->stop = true; }
      L->produced = t + 1;
    }
    L->cv.notify_all();
    if (e != cudaSuccess) return;
  }
}

int ng_loader_create(uint64_t* out_handle, const void* X, const void* Y, int64_t n_rows, uint64_t row_bytes,
                     int batch, int steps, const int64_t* plan, const int64_t* offsets, uint64_t x_dev0,
                     uint64_t x_dev1, uint64_t y_dev0, uint64_t y_dev1) {
  NG_ENSURE_INIT();
  NG_REQUIRE(X && n_rows > 0 && row_bytes > 0 && batch > 0 && steps >= 0, "ng_loader_create: bad arguments");
  for (int64_t i = 0; i < offsets[steps]; ++i)
    NG_REQUIRE(plan[i] >= 0 && plan[i] < n_rows, "ng_loader_create: index %lld out of range", (long long)plan[i]);
  for (int t = 0; t < steps; ++t)
    NG_REQUIRE(offsets[t + 1] - offsets[t] <= batch, "ng_loader_create: step %d has more than %d rows", t, batch);
  if (!g_copy_stream) {
    NG_CHECK_CUDA(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
    NG_CHECK_CUDA(cudaEventCreateWithFlags(&g_consumed, cudaEventDisableTiming));
  }
  NgLoader* L = new NgLoader();
  L->X = (const char*)X; L->Y = (const char*)Y;
  L->n_rows = n_rows; L->row_bytes = row_bytes; L->batch = batch; L->steps = steps;
  L->plan.assign(plan, plan + offsets[steps]);
  L->offset.assign(offsets, offsets + steps + 1);
  L->x_dev[0] = x_dev0; L->x_dev[1] = x_dev1; L->y_dev[0] = y_dev0; L->y_dev[1] = y_dev1;
  for (int k = 0; k < 2; ++k) {
    L->x_pin[k] = pinned_get((uint64_t)batch * row_bytes);
    if (Y) L->y_pin[k] = pinned_get((uint64_t)batch * 4);
    NG_REQUIRE(L->x_pin[k] && (!Y || L->y_pin[k]), "ng_loader_create: pinned host allocation failed");
    NG_CHECK_CUDA(cudaEventCreateWithFlags(&L->done[k], cudaEventDisableTiming));
    NG_CHECK_CUDA(cudaEventRecord(L->done[k], g_copy_stream));
  }
  NG_CHECK_CUDA(cudaEventCreateWithFlags(&L->consumed_ev, cudaEventDisableTiming));
  L->th = std::thread(loader_main, L);
  *out_handle = (uint64_t)(uintptr_t)L;
  return 0;
}

// Hands batch number `taken` to the caller: marks the previous batch consumed, waits for the loader
// thread to publish this one, makes the compute stream wait for its copy.  slot -> which device buffers.
int ng_loader_next(uint64_t handle, int* out_slot, int64_t* out_rows) {
  NgLoader* L = (NgLoader*)(uintptr_t)handle;
  NG_REQUIRE(L != nullptr, "ng_loader_next: null loader");
  int t;
  const bool dbg = getenv("NG_LOADER_DEBUG") != nullptr;
  const auto c0 = std::chrono::steady_clock::now();
  {
    std::unique_lock<std::mutex> lk(L->mu);
    t = L->taken;
    NG_REQUIRE(t < L->steps, "ng_loader_next: all %d batches were already taken", L->steps);
    L->taken = t + 1;   // batch t-1 is done with: its slot may be refilled with batch t+1
    L->cv.notify_all();
    L->cv.wait(lk, [&] { return L->stop || L->produced > t; });
    if (L->produced <= t) return set_error("ng_loader_next: loader thread failed: %s", L->err.c_str());
  }
  const int k = t & 1;
  const auto c1 = std::chrono::steady_clock::now();
  NG_CHECK_CUDA(cudaStreamWaitEvent(g_stream, L->done[k], 0));
  if (dbg) fprintf(stderr, "[next] t=%d wait %.3f streamwait %.3f ms\n", t, std::chrono::duration<double, std::milli>(c1 - c0).count(),
                   std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c1).count());
  *out_slot = k;
  *out_rows = L->offset[t + 1] - L->offset[t];
  return 0;
}

int ng_loader_destroy(uint64_t handle) {
  NgLoader* L = (NgLoader*)(uintptr_t)handle;
  if (!L) return 0;
  {
    std::lock_guard<std::mutex> lk(L->mu);
    L->stop = true;
  }
  L->cv.notify_all();
  if (L->th.joinable()) L->th.join();
  cudaStreamSynchronize(g_copy_stream);
  for (int k = 0; k < 2; ++k) {
    pinned_put(L->x_pin[k], (uint64_t)L->batch * L->row_bytes);
    pinned_put(L->y_pin[k], (uint64_t)L->batch * 4);
    if (L->done[k]) cudaEventDestroy(L->done[k]);
  }
  if (L->consumed_ev) cudaEventDestroy(L->consumed_ev);
  delete L;
  return 0;
}

int ng_pinned_alloc(void** out, uint64_t bytes) {
  NG_ENSURE_INIT();
  NG_CHECK_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}

int ng_pinned_free(void* p) {
  if (p) NG_CHECK_CUDA(cudaFreeHost(p));
  return 0;
}

int ng_event_create(uint64_t* out_ev) {
  NG_ENSURE_INIT();
  cudaEvent_t ev;
  NG_CHECK_CUDA(cudaEventCreate(&ev));
  *out_ev = (uint64_t)(uintptr_t)ev;
  return 0;
}
int ng_event_record(uint64_t ev) {
  NG_CHECK_CUDA(cudaEventRecord((cudaEvent_t)(uintptr_t)ev, g_stream));
  return 0;
}
int ng_event_sync(uint64_t ev) {
  NG_CHECK_CUDA(cudaEventSynchronize((cudaEvent_t)(uintptr_t)ev));
  return 0;
}
int ng_event_elapsed_ms(uint64_t a, uint64_t b, float* out_ms) {
  NG_CHECK_CUDA(cudaEventElapsedTime(out_ms, (cudaEvent_t)(uintptr_t)a, (cudaEvent_t)(uintptr_t)b));
  return 0;
}
int ng_event_destroy(uint64_t ev) {
  NG_CHECK_CUDA(cudaEventDestroy((cudaEvent_t)(uintptr_t)ev));
  return 0;
}

int ng_flush_l2(void) {
  NG_ENSURE_INIT();
  static void* scratch = nullptr;
  const uint64_t bytes = 256ull << 20;  // 2x the 126 MB L2
  if (!scratch) NG_CHECK_CUDA(cudaMalloc(&scratch, bytes));
  static int v = 0;
  NG_CHECK_CUDA(cudaMemsetAsync(scratch, (v++) & 0xff, bytes, g_stream));
  return 0;
}

int ng_measure_ffma_peak(int iters, float* out_tflops) {
  NG_ENSURE_INIT();
  if (iters <= 0) iters = 4096;
  float* d = nullptr;
  int r = pool_alloc((void**)&d, 64);
  if (r) return r;
  const int blocks = g_sm_count * 8, threads = 256;
  cudaEvent_t e0, e1;
  NG_CHECK_CUDA(cudaEventCreate(&e0));
  NG_CHECK_CUDA(cudaEventCreate(&e1));
  float best_ms = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {  // rep 0 is the warm-up
    NG_CHECK_CUDA(cudaEventRecord(e0, g_stream));
    NG_LAUNCH(ffma_peak_kernel, blocks, threads, 0, d, iters, 0.999f, 0.001f);
    NG_CHECK_LAUNCH();
    NG_CHECK_CUDA(cudaEventRecord(e1, g_stream));
    NG_CHECK_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    NG_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  pool_free(d);
  const double flops = 2.0 * 8 * 16 * (double)iters * threads * (double)blocks;
  *out_tflops = (float)(flops / (best_ms * 1e-3) / 1e12);
  return 0;
}

}  // extern "C"

// ---- per-kernel-class timing (bench.py roofline) ---------------------------------------------
// When enabled, the contraction launchers bracket each launch with a CUDA-event pair on the
// compute stream; ng_prof_read sums elapsed times per class.  Recording a timestamp does not
// serialise the stream.
namespace ng {

struct ProfRec { cudaEvent_t e0, e1; int cls; double flops, bytes; };
static bool g_prof_on = false;
static std::vector<ProfRec>* g_prof = nullptr;
static std::vector<cudaEvent_t>* g_prof_free = nullptr;
static const size_t kProfMax = 1u << 16;

static cudaEvent_t prof_event() {
  if (!g_prof_free) g_prof_free = new std::vector<cudaEvent_t>();
  if (!g_prof_free->empty()) {
    cudaEvent_t e = g_prof_free->back();
    g_prof_free->pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

int prof_begin(int cls, double flops, double bytes) {
  if (!g_prof_on) return -1;
  if (!g_prof) g_prof = new std::vector<ProfRec>();
  if (g_prof->size() >= kProfMax) return -1;
  ProfRec r;
  r.e0 = prof_event();
  r.e1 = prof_event();
  r.cls = cls; r.flops = flops; r.bytes = bytes;
  cudaEventRecord(r.e0, g_stream);
  g_prof->push_back(r);
  return (int)g_prof->size() - 1;
}

void prof_end(int token) {
  if (token < 0 || !g_prof) return;
  cudaEventRecord((*g_prof)[token].e1, g_stream);
}

}  // namespace ng

extern "C" {

int ng_prof_enable(int on) {
  g_prof_on = on != 0;
  return 0;
}

int ng_prof_reset(void) {
  if (g_prof) {
    if (!g_prof_free) g_prof_free = new std::vector<cudaEvent_t>();
    for (auto& r : *g_prof) { g_prof_free->push_back(r.e0); g_prof_free->push_back(r.e1); }
    g_prof->clear();
  }
  return 0;
}

int ng_prof_read(int cls, double* total_ms, int64_t* launches, double* flops, double* bytes) {
  NG_ENSURE_INIT();
  NG_CHECK_CUDA(cudaStreamSynchronize(g_stream));
  double ms = 0, fl = 0, by = 0;
  int64_t n = 0;
  if (g_prof) {
    for (auto& r : *g_prof) {
      if (r.cls != cls) continue;
      float t = 0.f;
      NG_CHECK_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
      ms += t; fl += r.flops; by += r.bytes; n++;
    }
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = n;
  if (flops) *flops = fl;
  if (bytes) *bytes = by;
  return 0;
}

}  // extern "C"
