// TEST INFRASTRUCTURE ONLY — see ng_emu.h.  Runs one CUDA block at a time on a pool of host
// threads (one per CUDA thread) with real barriers for __syncthreads and warp shuffles.
#include "ng_emu.h"

thread_local uint3 threadIdx;
thread_local uint3 blockIdx;
thread_local dim3 blockDim;
thread_local dim3 gridDim;

namespace ng_emu {

struct BlockCtx {
  int nthreads = 0;
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<float> xchg;
  std::vector<unsigned char> dyn;
};
static BlockCtx g_ctx;

struct Pool {
  int T;
  std::vector<std::thread> workers;
  std::barrier<> start_bar, done_bar;
  const std::function<void()>* body = nullptr;
  uint3 bidx{0, 0, 0};
  dim3 bdim, gdim;
  bool quit = false;
  explicit Pool(int t) : T(t), start_bar(t + 1), done_bar(t + 1) {
    for (int i = 0; i < T; ++i) workers.emplace_back([this, i]() { run(i); });
  }
  void run(int tid) {
    for (;;) {
      start_bar.arrive_and_wait();
      if (quit) return;
      threadIdx = uint3{(unsigned)tid, 0, 0};
      blockIdx = bidx;
      blockDim = bdim;
      gridDim = gdim;
      (*body)();
      g_ctx.block_bar->arrive_and_drop();
      g_ctx.warp_bar[tid >> 5]->arrive_and_drop();
      done_bar.arrive_and_wait();
    }
  }
};

static std::map<int, Pool*>& pools() {
  static std::map<int, Pool*>* p = new std::map<int, Pool*>();
  return *p;
}

void sync_block() { g_ctx.block_bar->arrive_and_wait(); }

float shfl(float v, int src_lane) {
  const int tid = (int)threadIdx.x, warp = tid >> 5;
  g_ctx.xchg[tid] = v;
  g_ctx.warp_bar[warp]->arrive_and_wait();
  const float r = g_ctx.xchg[warp * 32 + (src_lane & 31)];
  g_ctx.warp_bar[warp]->arrive_and_wait();
  return r;
}

void* dyn_smem_ptr() { return g_ctx.dyn.data(); }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  if (block.y != 1 || block.z != 1 || block.x % 32 != 0 || block.x == 0 || block.x > 1024) {
    fprintf(stderr, "ng_emu: unsupported block shape (%u,%u,%u)\n", block.x, block.y, block.z);
    abort();
  }
  // the limits a real launch has (a violation is cudaErrorInvalidConfiguration on the B200, silently fine on a host loop)
  if (grid.x == 0 || grid.y == 0 || grid.z == 0 || grid.y > 65535u || grid.z > 65535u || grid.x > 2147483647u ||
      smem > 227u * 1024u) {
    fprintf(stderr, "ng_emu: launch configuration a GPU would reject: grid (%u,%u,%u), %zu bytes of dynamic shared memory\n",
            grid.x, grid.y, grid.z, smem);
    abort();
  }
  const int T = (int)block.x;
  Pool*& pool = pools()[T];
  if (!pool) pool = new Pool(T);
  g_ctx.nthreads = T;
  g_ctx.xchg.assign(T, 0.f);
  g_ctx.dyn.assign(smem + 16, 0);
  pool->body = &body;
  pool->bdim = block;
  pool->gdim = grid;
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        g_ctx.block_bar = std::make_unique<std::barrier<>>(T);
        g_ctx.warp_bar.clear();
        for (int w = 0; w < T / 32; ++w) g_ctx.warp_bar.push_back(std::make_unique<std::barrier<>>(32));
        pool->bidx = uint3{x, y, z};
        pool->start_bar.arrive_and_wait();
        pool->done_bar.arrive_and_wait();
      }
}

}  // namespace ng_emu
