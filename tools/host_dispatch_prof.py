"""Host-side dispatch cost of one training step, measured WITHOUT a GPU.

Generates a "null" build of the C ABI (every entry point of nanograd_b200/backend.py:_PROTOTYPES returns 0
immediately; ng_malloc hands out fake addresses, ng_conv2d_tc_plan answers like the sm_100a library does for
the FP16x3 / 3xTF32 engines) into the git-ignored tools/bin/, binds it, and runs BASELINE config 2 (VGG / BN /
Adam, batch 512) or config 0 (README MNIST CNN) through the product's Python path.  What is left is exactly the
Python + ctypes time the host needs to dispatch a step (`host_ms_per_step` of the bench line): the number that
has to stay below the GPU's step time for the end-to-end rate to equal the device-resident rate.

    python tools/host_dispatch_prof.py [--model vgg|mnist] [--steps 200] [--profile]

Diagnostic only: nothing in the package, the tests or bench.py uses the null library.
"""
import argparse
import cProfile
import io
import os
import pstats
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SPECIAL = r"""
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
static uint64_t g_next = 0x100000000ull, g_launch = 0, g_handle = 1, g_mallocs = 0, g_frees = 0;
const char* ng_last_error(void) { return "null library"; }
int ng_malloc(uint64_t* out, uint64_t n) { *out = g_next; g_next += (n + 511) & ~511ull; g_mallocs++; return 0; }
int ng_free(uint64_t p) { (void)p; g_frees++; return 0; }
/* live_bytes slot = allocations not yet freed (a count, not bytes) */
int ng_launch_count(uint64_t* out) { *out = g_launch; return 0; }
int ng_device_info(int* sm, int* mj, int* mn, uint64_t* mem) { *sm = 148; *mj = 10; *mn = 0; *mem = 180ull << 30; return 0; }
int ng_mem_stats(uint64_t* a, uint64_t* b, uint64_t* c) { *a = g_mallocs - g_frees; *b = 0; *c = g_mallocs; return 0; }
int ng_event_create(uint64_t* h) { *h = g_handle++; return 0; }
int ng_event_elapsed_ms(uint64_t a, uint64_t b, float* ms) { (void)a; (void)b; *ms = 1.0f; return 0; }
int ng_pinned_alloc(void** p, uint64_t n) { *p = malloc(n); return 0; }
int ng_pinned_free(void* p) { free(p); return 0; }
int ng_d2h(void* dst, uint64_t src, uint64_t n) { (void)src; memset(dst, 0, n); return 0; }
int ng_d2h_after(void* dst, uint64_t src, uint64_t n, uint64_t ev) { (void)src; (void)ev; memset(dst, 0, n); return 0; }
int ng_graph_capturing(int* f) { *f = 0; return 0; }
int ng_measure_ffma_peak(int n, float* out) { (void)n; *out = 1.0f; return 0; }
int ng_nccl_world(int* r, int* w) { *r = 0; *w = 1; return 0; }
/* conv_engine() + ng_conv2d_tc_plan() of csrc/conv_gemm.cu, gather/wgrad_supported of csrc/umma_conv.cu
   (flags: TF32 2, TF32X3 4, TC_FORCE 8, F16X3 64) */
int ng_conv2d_tc_plan(int N, int C, int H, int W, int K, int R, int S, int stride, int OH, int OW, int flags, int* out) {
  (void)H; (void)W;
  double flops = 2.0 * N * K * C * R * S * (double)OH * OW;
  int engine = 0, RS = R * S, f = 0, d = 0, w = 0;
  if (stride == 1) {
    if (flags & 2) engine = 1;
    else if ((flags & (4 | 64)) && ((flags & 8) || flops >= 2.0e8)) {
      if ((flags & 64) && C % 64 == 0 && K % 64 == 0 && RS <= 64) engine = 2;
      else engine = (flags & 4) ? 1 : 0;
    }
  }
  if (engine == 2) f = d = w = 1;
  else if (engine == 1) {
    f = (C % 4 == 0 && C >= 8 && RS <= 64 && K >= 8);
    d = (K % 4 == 0 && K >= 8 && RS <= 64 && C >= 8);
    w = (C % 32 == 0 && K % 32 == 0 && RS <= 64);
  }
  out[0] = f; out[1] = d; out[2] = w; out[3] = engine;
  return 0;
}
"""


_GENERIC = []  # entry points with a generated counting stub, in g_cnt order


def build_null_library():
    from nanograd_b200 import backend as B
    out_dir = os.path.join(ROOT, "tools", "bin")
    os.makedirs(out_dir, exist_ok=True)
    src, lib = os.path.join(out_dir, "ng_null.c"), os.path.join(out_dir, "libng_null.so")
    special = {ln.split("(")[0].split()[-1].lstrip("*") for ln in SPECIAL.splitlines() if ln.startswith(("int ng_", "const char* ng_"))}
    del _GENERIC[:]
    _GENERIC.extend(n for n in B.EXPORTED_SYMBOLS if n not in special)
    body = [SPECIAL, "static uint64_t g_cnt[%d];" % len(_GENERIC)]
    for i, name in enumerate(_GENERIC):
        # unprototyped definition: the System V x86-64 callee may ignore whatever the caller passed
        body.append("int %s() { g_cnt[%d]++; g_launch++; return 0; }" % (name, i))
    body.append("int ng_null_counts(uint64_t* out) { memcpy(out, g_cnt, sizeof(g_cnt)); return %d; }" % len(_GENERIC))
    with open(src, "w") as f:
        f.write("\n".join(body) + "\n")
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-w", "-std=gnu89", "-o", lib, src], check=True)
    return lib


def call_counts(handle):
    """{entry point: calls so far} of the generated stubs of the bound null library."""
    import ctypes
    buf = (ctypes.c_uint64 * len(_GENERIC))()
    fn = handle.ng_null_counts
    fn.restype, fn.argtypes = ctypes.c_int, [ctypes.POINTER(ctypes.c_uint64)]
    fn(buf)
    return {n: int(buf[i]) for i, n in enumerate(_GENERIC)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="vgg", choices=["vgg", "mnist", "attention"])
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batch", type=int, default=0, help="0 = the BASELINE batch (512 / 128)")
    ap.add_argument("--math", default="fp32")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--world", type=int, default=1, help="> 1: this process plays rank 0 of a data-parallel group whose "
                    "collectives are the null library's no-ops (host logic of the gradient arena / chunked allreduce)")
    ap.add_argument("--sync-bn", action="store_true")
    ap.add_argument("--reference", action="store_true", help="drive the step through the unmodified reference's own "
                    "Tensor / Module / Optimizer (oracle/_ref or /root/reference) bound with nanograd_b200.integrate")
    ap.add_argument("--json", action="store_true", help="one JSON line: calls per step by entry point, allocation balance")
    args = ap.parse_args()

    import numpy as np
    from nanograd_b200 import backend as B
    handle = B.load_library(build_null_library())
    B.init(0)
    B.set_math_mode(args.math)
    if args.world > 1:
        import tempfile
        from nanograd_b200 import dist
        os.environ.update(WORLD_SIZE=str(args.world), RANK="0", LOCAL_RANK="0", MASTER_PORT="0",
                          NANOGRAD_B200_RDZV_DIR=tempfile.mkdtemp(prefix="ng_null_rdzv_"))
        dist.init(use_device=True, sync_bn=args.sync_bn)
    if args.reference:
        import helpers
        ref = helpers.load_reference()
        if ref is None:
            raise SystemExit("the reference is not present (run tools/make_oracle_ref.sh where /root/reference exists)")
        import nanograd_b200.integrate as integrate
        integrate.install(ref["package"], library=B.library_path())
        nn, optim, Device, Tensor = ref["module"], ref["optim"], ref["device"].Device, ref["tensor"].Tensor
    else:
        import nanograd_b200.nn.module as nn
        import nanograd_b200.optim.optimizer as optim
        from nanograd_b200.device import Device
        from nanograd_b200.tensor import Tensor

    np.random.seed(0)
    if args.model == "vgg":
        import bench
        batch = args.batch or 512
        model = bench.build_model(nn)
        model.gpu()
        opt = optim.Adam(model.parameters(), lr=1e-3)
        X = np.zeros((batch, 3, 32, 32), np.float32)
    elif args.model == "attention":
        # BASELINE configs[3] as bench.attention_extra composes it: seq 512, d 256, MSE loss, AdamW
        seq, d = args.batch or 512, 256

        class Attn(nn.Module):
            def __init__(self):
                super().__init__()
                mk = lambda: Tensor((np.random.randn(d, d) / np.sqrt(d)).astype(np.float32), requires_grad=True,
                                    is_parameter=True)
                self.wq, self.wk, self.wv, self.wo = mk(), mk(), mk(), mk()

            def forward(self, x):
                q, k, v = x @ self.wq, x @ self.wk, x @ self.wv
                s = (q @ k.T) * (1.0 / np.sqrt(d))
                p = s.log_softmax().exp()
                return (p @ v) @ self.wo

        batch = seq
        model = Attn()
        model.gpu()
        opt = optim.AdamW(model.parameters(), lr=1e-3)
        X = np.zeros((seq, d), np.float32)
    else:
        import helpers
        batch = args.batch or 128
        model = helpers.build_mnist_cnn(nn)
        model.gpu()
        opt = optim.SGD(model.parameters(), lr=0.01, momentum=0)
        X = np.zeros((batch, 784), np.float32)
    Y = np.zeros((batch, 256) if args.model == "attention" else batch, np.float32)
    crit = nn.MSELoss() if args.model == "attention" else nn.NLLLoss()
    xb, yb = Tensor(X, device=Device.GPU), Tensor(Y, device=Device.GPU)

    def step():
        if args.model == "vgg":
            return bench.train_step(model, opt, crit, xb, yb)   # the step bench.py times
        out = model(xb)
        opt.zero_grad()
        loss = crit(out, yb)
        loss.backward()
        opt.step()
        return loss

    import gc
    for _ in range(10):
        loss = step()
    del loss
    gc.collect()
    gc.disable()
    # steady state: one step's calls by entry point, and the allocations a step leaves behind (must be none)
    live0, c0 = B.mem_stats(), call_counts(handle)
    loss = step()
    float(loss.data.numpy()[0])
    del loss
    live1, c1 = B.mem_stats(), call_counts(handle)
    per_step = {n: c1[n] - c0[n] for n in c1 if c1[n] != c0[n]}
    l0 = B.launch_count()
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        best = min(best, (time.perf_counter() - t0) / args.steps)
    calls = (B.launch_count() - l0) / (5 * args.steps)
    live2 = B.mem_stats()
    if args.json:
        import json
        print(json.dumps({"model": args.model, "batch": batch, "math": args.math, "world": args.world, "reference_classes": bool(args.reference), "host_ms_per_step": round(best * 1e3, 4),
                          "library_calls_per_step": calls, "calls": per_step,
                          "allocs_per_step": live1["device_allocs"] - live0["device_allocs"],
                          "live_growth_one_step": live1["live_bytes"] - live0["live_bytes"],
                          "live_growth_timed_steps": live2["live_bytes"] - live1["live_bytes"]}))
    else:
        print("%s: host dispatch %.3f ms/step, %.1f library calls/step (%.1f us per call), %d allocations/step, "
              "live allocations +%d over %d steps" %
              (args.model, best * 1e3, calls, best * 1e6 / max(calls, 1), live1["device_allocs"] - live0["device_allocs"],
               live2["live_bytes"] - live1["live_bytes"], 5 * args.steps))
    if args.profile:
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(args.steps):
            step()
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(args.top)
        print(s.getvalue())


if __name__ == "__main__":
    main()
