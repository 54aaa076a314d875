// TEST INFRASTRUCTURE ONLY — see ng_emu.h.  Runs one CUDA block at a time: its threads are cooperative contexts on
// the calling host thread (default) or one host thread each (NG_EMU_THREADS=1), with real barriers for __syncthreads
// and the warp shuffles.
#include "ng_emu.h"

#include <sys/mman.h>
#include <ucontext.h>

thread_local uint3 threadIdx;
thread_local uint3 blockIdx;
thread_local dim3 blockDim;
thread_local dim3 gridDim;

namespace ng_emu {

// Two executors for a block of T CUDA threads:
//   fibers (default)  the T threads are cooperative contexts (ucontext) on the calling host thread; __syncthreads and
//                     the warp exchanges yield to the next context, so a barrier costs T user-level switches
//                     instead of T futex sleeps / wake-ups (the thread pool below spends 80 % of its time there);
//   threads           NG_EMU_THREADS=1: one host thread per CUDA thread with std::barrier — the first implementation,
//                     kept as an independent cross-check (really concurrent threads expose races fibers cannot).
// Blocks run one after the other in both (static __shared__ arrays are shared by whatever runs concurrently).

// ------------------------------------------------------------------------------------------ fibers
struct Bar {
  int expected = 0, arrived = 0;
  unsigned gen = 0;
};
// Context switch.  x86-64: six callee-saved registers and the stack pointer (glibc's swapcontext also saves the
// signal mask — two system calls per switch, which dominated blocks of 256 threads x a few barriers); elsewhere ucontext.
#if defined(__x86_64__)
#define NG_EMU_ASM_SWITCH 1
extern "C" void ng_emu_ctx_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl ng_emu_ctx_switch
.type ng_emu_ctx_switch,@function
ng_emu_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size ng_emu_ctx_switch,.-ng_emu_ctx_switch
)");
struct Ctx { void* sp = nullptr; };
static inline void ctx_switch(Ctx& from, Ctx& to) { ng_emu_ctx_switch(&from.sp, to.sp); }
#else
#define NG_EMU_ASM_SWITCH 0
struct Ctx { ucontext_t uc; };
static inline void ctx_switch(Ctx& from, Ctx& to) { swapcontext(&from.uc, &to.uc); }
#endif

struct Fiber {
  Ctx ctx;
  bool done = false;
};
struct FiberBlock {
  int T = 0, cur = 0;
  std::vector<Fiber> fibers;
  std::vector<void*> stacks;
  Ctx sched;
  Bar block;
  std::vector<Bar> warp;
  const std::function<void()>* body = nullptr;
  dim3 bdim, gdim;
  uint3 bidx{0, 0, 0};
};
static FiberBlock g_fb;
static bool g_use_fibers = true;
static const size_t kStackBytes = 256 * 1024;

static void fiber_yield() { ctx_switch(g_fb.fibers[g_fb.cur].ctx, g_fb.sched); }

static void bar_wait(Bar& b) {
  const unsigned my = b.gen;
  if (++b.arrived >= b.expected) {
    b.arrived = 0;
    ++b.gen;
    return;
  }
  do fiber_yield(); while (b.gen == my);
}
static void bar_drop(Bar& b) {
  --b.expected;
  if (b.expected > 0 && b.arrived >= b.expected) {
    b.arrived = 0;
    ++b.gen;
  }
}

static void fiber_entry() {
  const int tid = g_fb.cur;
  (*g_fb.body)();
  g_fb.fibers[tid].done = true;
  bar_drop(g_fb.block);
  bar_drop(g_fb.warp[tid >> 5]);
#if NG_EMU_ASM_SWITCH
  ctx_switch(g_fb.fibers[tid].ctx, g_fb.sched);   // never resumed
  abort();
#endif
  // ucontext: returning resumes uc_link = the scheduler
}

static void fiber_prepare(Fiber& f, void* stack) {
  f.done = false;
#if NG_EMU_ASM_SWITCH
  // frame that ng_emu_ctx_switch pops: r15 r14 r13 r12 rbx rbp, then `ret` into fiber_entry with the stack aligned
  // as after a call (rsp = 8 mod 16) and a null return address above it
  uintptr_t top = ((uintptr_t)stack + kStackBytes) & ~(uintptr_t)15;
  void** sp = (void**)top;
  *--sp = nullptr;                 // fiber_entry's "return address" (it never returns)
  *--sp = (void*)&fiber_entry;
  for (int i = 0; i < 6; ++i) *--sp = nullptr;
  f.ctx.sp = sp;
#else
  getcontext(&f.ctx.uc);
  f.ctx.uc.uc_stack.ss_sp = stack;
  f.ctx.uc.uc_stack.ss_size = kStackBytes;
  f.ctx.uc.uc_link = &g_fb.sched.uc;
  makecontext(&f.ctx.uc, fiber_entry, 0);
#endif
}

static void run_block_fibers(int T) {
  FiberBlock& F = g_fb;
  if ((int)F.stacks.size() < T) {
    const size_t have = F.stacks.size();
    F.stacks.resize(T, nullptr);
    for (int i = (int)have; i < T; ++i) {
      F.stacks[i] = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (F.stacks[i] == MAP_FAILED) { perror("ng_emu: mmap of a fiber stack"); abort(); }
    }
  }
  if ((int)F.fibers.size() < T) F.fibers.resize(T);
  F.T = T;
  F.block = Bar{T, 0, 0};
  F.warp.assign(T / 32, Bar{32, 0, 0});
  for (int i = 0; i < T; ++i) fiber_prepare(F.fibers[i], F.stacks[i]);
  int live = T;
  while (live > 0) {
    for (int i = 0; i < T; ++i) {
      Fiber& f = F.fibers[i];
      if (f.done) continue;
      F.cur = i;
      threadIdx = uint3{(unsigned)i, 0, 0};
      blockIdx = F.bidx;
      blockDim = F.bdim;
      gridDim = F.gdim;
      ctx_switch(F.sched, f.ctx);
      if (f.done) --live;
    }
  }
}

// ------------------------------------------------------------------------------------------ host threads
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

// ------------------------------------------------------------------------------------------ device intrinsics
void sync_block() {
  if (g_use_fibers) bar_wait(g_fb.block);
  else g_ctx.block_bar->arrive_and_wait();
}

float shfl(float v, int src_lane) {
  const int tid = (int)threadIdx.x, warp = tid >> 5;
  g_ctx.xchg[tid] = v;
  if (g_use_fibers) bar_wait(g_fb.warp[warp]); else g_ctx.warp_bar[warp]->arrive_and_wait();
  const float r = g_ctx.xchg[warp * 32 + (src_lane & 31)];
  if (g_use_fibers) bar_wait(g_fb.warp[warp]); else g_ctx.warp_bar[warp]->arrive_and_wait();
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
  static const bool threads_mode = [] { const char* e = getenv("NG_EMU_THREADS"); return e && e[0] == '1'; }();
  g_use_fibers = !threads_mode;
  const int T = (int)block.x;
  g_ctx.nthreads = T;
  g_ctx.xchg.assign(T, 0.f);
  g_ctx.dyn.assign(smem + 16, 0);
  if (g_use_fibers) {
    // callers of the C ABI are serialised by the library (one compute stream), but not necessarily on one host
    // thread (the loader thread launches nothing): one block executor at a time
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    g_fb.body = &body;
    g_fb.bdim = block;
    g_fb.gdim = grid;
    const uint3 t0 = threadIdx, b0 = blockIdx;
    const dim3 bd0 = blockDim, gd0 = gridDim;
    for (unsigned z = 0; z < grid.z; ++z)
      for (unsigned y = 0; y < grid.y; ++y)
        for (unsigned x = 0; x < grid.x; ++x) {
          g_fb.bidx = uint3{x, y, z};
          run_block_fibers(T);
        }
    threadIdx = t0; blockIdx = b0; blockDim = bd0; gridDim = gd0;
    return;
  }
  Pool*& pool = pools()[T];
  if (!pool) pool = new Pool(T);
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
