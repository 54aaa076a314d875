#!/usr/bin/env python
"""bench.py — train images/sec of the CIFAR-shaped VGG/BatchNorm CNN on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" is one full training step (forward, NLL loss, backward, Adam) of BASELINE config 2 —
[Conv3x3(pad 1)-BatchNorm2d-ReLU]x2-MaxPool for 64/128/256 channels, Linear(4096,256)-ReLU-
Linear(256,10), log_softmax + NLLLoss, Adam(lr 1e-3) — on a batch of 512 synthetic 3x32x32 images
per GPU.  With N GPUs (launched by torchrun, one rank per GPU) every rank trains its own shard of
the global batch and the gradients are summed with NCCL (chunks of the flat gradient arena on a
communication stream while the backward pass is still running).  Default: weak scaling, 512 images
per GPU (N = 8 is BASELINE config 4, global batch 4096), per-shard BatchNorm statistics.
--scaling strong: the global batch stays 4096 (4096/2048/1024/512 images per GPU
at N = 1/2/4/8); --sync-bn: BatchNorm statistics over the global batch (the reference's
single-process semantics), at the price of two small exchanges per BatchNorm layer and direction.

Printed JSON line (rank 0):
  value        images/s with inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e          the same step through the public API with the step's batch copied from pinned
               host memory and the loss read back to the host, inside the timed region
  roofline     dominant kernel class (conv fprop/dgrad/wgrad or GEMM), timed with CUDA events
               on the compute stream during the timed steps; FP32 SIMT path => the bound is the
               FFMA pipe, peak = FFMA-saturating microkernel measured in this run
  cpu_baseline the NumPy oracle (port of the reference's CPU path) timed on this box's host
               cores on a bounded sample of the same workload (rank 0, N = 1 only)
  conv2d_sweep BASELINE config 1 (Conv2d 3x3, 32x32, BS 256, C=K in 16..256): TFLOP/s per pass
  mnist_cnn    BASELINE config 0 (README MNIST CNN, BS 128, SGD) images/s on the device

--impl reference times the UNMODIFIED reference (PABannier/nanograd pip-installed into the
git-ignored oracle/_ref by tools/make_oracle_ref.sh, which travels to the GPU box; /root/reference
in the build container; the oracle port only if neither exists) through its own public API —
its Tensor / nn.Sequential / Adam / NLLLoss on Device.CPU — on all host cores, for the same model,
optimizer and loss.  Each of the W + K steps is one full training step on a bounded sample of
the workload's batch (64 of the 512 images: the NumPy path runs ~10 images/s, a 512-image step
costs ~30-60 s and, through the lru_cache on Tensor.build_graph_topology, ~20 GB of host memory
that is never released); images/s does not depend on the sample size to within a few percent.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PER_GPU_BATCH = 512
STRONG_GLOBAL_BATCH = 4096
CPU_SAMPLE_BATCH = 64
LR = 1e-3
WORKLOAD = "cifar_vgg_bn_adam_bs512_per_gpu (BASELINE configs[2]; x8 GPUs = configs[4], global batch 4096)"


# ----------------------------------------------------------------------------- model / data
def build_model(nn):
    layers, c = [], 3
    for w in (64, 128, 256):
        layers += [nn.Conv2d(c, w, 3, 1, 1), nn.BatchNorm2d(w), nn.ReLU(),
                   nn.Conv2d(w, w, 3, 1, 1), nn.BatchNorm2d(w), nn.ReLU(), nn.MaxPool2d(2)]
        c = w
    layers += [nn.Flatten(), nn.Linear(256 * 4 * 4, 256), nn.ReLU(), nn.Linear(256, 10)]
    return nn.Sequential(*layers)


def synthetic_batch(batch, seed):
    rng = np.random.RandomState(seed)
    return rng.randn(batch, 3, 32, 32).astype(np.float32), rng.randint(0, 10, batch).astype(np.float32)


def train_step(model, opt, criterion, xb, yb):
    out = model(xb).log_softmax()
    opt.zero_grad()
    loss = criterion(out, yb)
    loss.backward()
    opt.step()
    return loss


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index, self.lines, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arms
def _import_reference():
    """The unmodified reference as an importable package: oracle/_ref (pip-installed copy, present on
    the GPU box) or /root/reference (build container).  Returns (modules, where) or (None, why)."""
    import importlib
    import warnings
    for root in (os.path.join(ROOT, "oracle", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(root, "nanograd", "tensor.py")):
            sys.path.insert(0, os.path.join(ROOT, "tests", "_shim"))   # 4-line graphviz stand-in
            sys.path.insert(0, root)
            sys.dont_write_bytecode = True
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                mods = {k: importlib.import_module(f"nanograd.{k}") for k in
                        ("tensor", "nn.module", "optim.optimizer", "device")}
            return mods, root
    return None, "neither oracle/_ref nor /root/reference holds the reference"


def reference_images_per_sec(steps, warmup, batch=CPU_SAMPLE_BATCH):
    """The real reference, one continuous run: model and Adam built once, W + K training steps
    (so the NumPy path's drift to float64 after its first optimizer step is part of what is timed,
    as it is for any user of the reference).  The lru_cache of Tensor.build_graph_topology is cleared
    between steps — it would otherwise keep every step's graph and activations alive."""
    import warnings
    warnings.simplefilter("ignore")      # the reference's Pow.backward takes log of negative bases (ops_cpu.py:251)
    np.seterr(all="ignore")
    mods, where = _import_reference()
    if mods is None:
        return None, where
    T, nn, optim = mods["tensor"], mods["nn.module"], mods["optim.optimizer"]
    np.random.seed(0)
    model = build_model(nn)
    opt = optim.Adam(model.parameters(), lr=LR)
    crit = nn.NLLLoss()
    rng = np.random.RandomState(0)
    times = []
    for i in range(warmup + steps):
        X = rng.randn(batch, 3, 32, 32).astype(np.float32)
        Y = rng.randint(0, 10, batch).astype(np.float32)
        t0 = time.perf_counter()
        loss = train_step(model, opt, crit, T.Tensor(X), T.Tensor(Y))
        float(loss.data[0])
        dt = time.perf_counter() - t0
        cache_clear = getattr(T.Tensor.build_graph_topology, "cache_clear", None)
        if cache_clear:
            cache_clear()
        if i >= warmup:
            times.append(dt)
    return (batch / float(np.mean(times)), float(np.mean(times))), where


def cpu_oracle_images_per_sec(steps, warmup, batch=CPU_SAMPLE_BATCH):
    """Fallback when the reference itself is not importable: the oracle port of its NumPy path (this
    repo's host mirror on Device.CPU with oracle/ registered: the reference's graph, primitive by
    primitive), one continuous run like reference_images_per_sec."""
    import nanograd_b200.nn.module as nn
    import nanograd_b200.optim.optimizer as optim
    from nanograd_b200.tensor import Tensor
    import oracle
    oracle.register_host()
    np.random.seed(0)
    model = build_model(nn)
    opt = optim.Adam(model.parameters(), lr=LR)
    crit = nn.NLLLoss()
    rng = np.random.RandomState(0)
    times = []
    for i in range(warmup + steps):
        X = rng.randn(batch, 3, 32, 32).astype(np.float32)
        Y = rng.randint(0, 10, batch).astype(np.float32)
        t0 = time.perf_counter()
        loss = train_step(model, opt, crit, Tensor(X), Tensor(Y))
        float(loss.data[0])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return batch / float(np.mean(times)), float(np.mean(times))


def cpu_arm(steps, warmup):
    """(images/s, s per step, kind, description of what was run)."""
    res, where = reference_images_per_sec(steps, warmup)
    if res is not None:
        return res[0], res[1], "reference", f"unmodified PABannier/nanograd 1.0.4 imported from {os.path.relpath(where, ROOT)}"
    ips, sec = cpu_oracle_images_per_sec(steps, warmup)
    return ips, sec, "port", f"oracle port ({where})"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def workload_config(world, n_params=2198602, math_mode="fp32", batch=None, args=None):
    """The `config` object of the JSON line — identical for both arms (the reference arm runs the same
    workload definition; what it samples per step is said in cpu_baseline.sample)."""
    scaling = getattr(args, "scaling", "weak")
    if batch is None:
        batch = STRONG_GLOBAL_BATCH // world if scaling == "strong" else PER_GPU_BATCH
    sync_bn = bool(getattr(args, "sync_bn", False)) and world > 1
    return {"workload": WORKLOAD, "global_batch": batch * world, "per_gpu_batch": batch,
            "image": "3x32x32", "optimizer": "Adam(lr=1e-3)", "loss": "log_softmax + NLLLoss",
            "params": n_params, "parallelism": f"dp{world}" if world > 1 else "single",
            "batchnorm": ("global-batch statistics (SyncBN: the single-process reference step at the global batch)"
                          if sync_bn else "per-shard statistics" if world > 1 else "batch statistics"),
            "l2": "no flush: one step touches ~1.9 GB of activations, far larger than the 126 MB L2",
            "math": MATH_TEXT[math_mode]}


MATH_TEXT = {"fp32": "FP32-accurate: tcgen05 tensor-core conv with error-compensated FP16x3 operand pairs (channel "
                     "counts multiples of 64) or the 3xTF32 split (other supported shapes), FP32 SIMT (FFMA) "
                     "elsewhere (first conv layer, small GEMMs)",
             "fp32_tf32x3": "FP32-accurate: tcgen05 tensor-core conv with 3xTF32 error compensation where the shape is "
                            "supported, FP32 SIMT (FFMA) elsewhere",
             "f16x3": "FP32-accurate FP16x3 / 3xTF32 tcgen05 conv forced for every supported shape",
             "tf32x3": "FP32-accurate 3xTF32 tcgen05 conv forced for every supported shape",
             "simt": "FP32 SIMT (FFMA) implicit-GEMM conv / GEMM only",
             "tf32": "TF32 tcgen05 tensor-core conv (FP32 accumulate), FP32 SIMT elsewhere"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    ips, sec, kind, what = cpu_arm(steps, warmup)
    sample = (f"{what}; {warmup} warm-up + {steps} timed training steps, each on {CPU_SAMPLE_BATCH} of the workload's "
              f"{PER_GPU_BATCH} images (same model, Adam, log_softmax + NLLLoss), {sec:.2f} s per step, float32 inputs "
              f"(the NumPy path computes in float64 from its first optimizer step on)")
    line = {
        "impl": "reference", "metric": "train_images_per_sec", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.gpus, args=args),
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- device arm
def prof_read(lib, cls):
    ms, n, fl, by = ctypes.c_double(0), ctypes.c_int64(0), ctypes.c_double(0), ctypes.c_double(0)
    lib.ng_prof_read(cls, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(fl), ctypes.byref(by))
    return {"ms": ms.value, "launches": n.value, "flops": fl.value, "bytes": by.value}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_class):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(kernel_class)
        except Exception:
            return None
    return None


def conv2d_sweep(B, lib, ffma_peak):
    """BASELINE config 1: per-pass TFLOP/s of the FP32 SIMT kernels (fraction of the measured FFMA
    peak), of the FP32-accurate 3xTF32 tensor-core kernels (useful FLOPs: one per FP32 MAC) and of the
    opt-in TF32 tensor-core kernels (staging transposes included in the time)."""
    from nanograd_b200.nn.buffer import DeviceBuffer
    rows = []
    rng = np.random.RandomState(0)
    N, H, W, R, S = 256, 32, 32, 3, 3
    OH, OW = H - R + 1, W - S + 1
    for ch in (16, 32, 64, 128, 256):
        C = K = ch
        x = DeviceBuffer((N, C, H, W), hostbuf=rng.randn(N, C, H, W).astype(np.float32))
        w = DeviceBuffer((K, C, R, S), hostbuf=(rng.randn(K, C, R, S) / np.sqrt(9 * C)).astype(np.float32))
        dy = DeviceBuffer((N, K, OH, OW), hostbuf=rng.randn(N, K, OH, OW).astype(np.float32))
        y, dx, dw = DeviceBuffer((N, K, OH, OW)), DeviceBuffer((N, C, H, W)), DeviceBuffer((K, C, R, S))
        dims = (N, C, H, W, K, R, S, 1, 0, 0, OH, OW)
        calls = {"fprop": lambda f: lib.ng_conv2d_fprop(y.ptr, x.ptr, w.ptr, 0, *dims, f),
                 "dgrad": lambda f: lib.ng_conv2d_dgrad(dx.ptr, dy.ptr, w.ptr, *dims, f),
                 "wgrad": lambda f: lib.ng_conv2d_wgrad(dw.ptr, dy.ptr, x.ptr, *dims, f)}
        flops = 2.0 * N * K * C * R * S * OH * OW
        row = {"C=K": ch}
        iters = 20 if ch <= 64 else 5
        modes = [("", 0), ("tf32x3_", B.F_TF32X3 | B.F_TC_FORCE), ("tf32_", B.F_TF32)]
        if ch % 64 == 0:
            modes.insert(2, ("f16x3_", B.F_F16X3 | B.F_TC_FORCE))
        for mode, flag in modes:
            total = 0.0
            for name, fn in calls.items():
                for _ in range(3):
                    fn(flag)
                e0, e1 = B.Event().record(), B.Event()
                for _ in range(iters):
                    fn(flag)
                e1.record()
                e1.synchronize()
                ms = e0.elapsed_ms(e1) / iters
                total += ms
                row[f"{mode}{name}_tflops"] = round(flops / (ms * 1e-3) / 1e12, 2)
            row[f"{mode}fwd_bwd_tflops"] = round(3 * flops / (total * 1e-3) / 1e12, 2)
        row["frac_of_ffma_peak"] = round(row["fwd_bwd_tflops"] / ffma_peak, 3) if ffma_peak else None
        if ch % 64 == 0:
            # the layer-level cost of the FP16x3 engine, as a training step pays it: x, dy and the filters are staged
            # ONCE (timed as `staging`) and shared by the three passes (x: fprop + wgrad, dy: dgrad + wgrad)
            flag = B.F_F16X3 | B.F_TC_FORCE
            xs, dys = DeviceBuffer((x.size + 16,)), DeviceBuffer((dy.size + 16,))
            wf, wd = DeviceBuffer((w.size + 16,)), DeviceBuffer((w.size + 16,))
            staged = {"staging": lambda: (lib.ng_conv2d_stage_f16(xs.ptr, x.ptr, 0, 0, N, C, H * W),
                                          lib.ng_conv2d_stage_f16(dys.ptr, dy.ptr, 0, 0, N, K, OH * OW),
                                          lib.ng_conv2d_stage_filters_f16(wf.ptr, wd.ptr, w.ptr, K, C, R * S)),
                      "fprop": lambda: lib.ng_conv2d_fprop(y.ptr, xs.ptr, wf.ptr, 0, *dims, flag | B.F_X_NHWC | B.F_W_STAGED),
                      "dgrad": lambda: lib.ng_conv2d_dgrad(dx.ptr, dys.ptr, wd.ptr, *dims, flag | B.F_DY_NHWC | B.F_W_STAGED),
                      "wgrad": lambda: lib.ng_conv2d_wgrad(dw.ptr, dys.ptr, xs.ptr, *dims, flag | B.F_X_NHWC | B.F_DY_NHWC)}
            total = 0.0
            for name, fn in staged.items():
                for _ in range(3):
                    fn()
                e0, e1 = B.Event().record(), B.Event()
                for _ in range(iters):
                    fn()
                e1.record()
                e1.synchronize()
                ms = e0.elapsed_ms(e1) / iters
                total += ms
                if name == "staging":
                    row["f16x3_layer_staging_ms"] = round(ms, 4)
                else:
                    row[f"f16x3_layer_{name}_tflops"] = round(flops / (ms * 1e-3) / 1e12, 2)
            row["f16x3_layer_fwd_bwd_tflops"] = round(3 * flops / (total * 1e-3) / 1e12, 2)
            del xs, dys, wf, wd
        rows.append(row)
        del x, w, dy, y, dx, dw
    return rows


def mnist_cnn_images_per_sec(B, steps=200, warmup=10):
    """BASELINE config 0: the README MNIST CNN, batch 128, SGD(lr 0.01), NLLLoss.
    eager   — the reference's loop shape: every primitive dispatched from Python (device-resident batch);
    graphed — nanograd_b200.graph.GraphedStep: the same step captured once, then per step one H2D copy of
              a NEW host batch from pinned staging plus one cudaGraphLaunch (this is the end-to-end figure:
              host batches in, loss read back at the end)."""
    import helpers
    import nanograd_b200.nn.module as nn
    import nanograd_b200.optim.optimizer as optim
    from nanograd_b200.device import Device
    from nanograd_b200.graph import GraphedStep
    from nanograd_b200.tensor import Tensor
    import gc

    def setup():
        np.random.seed(0)
        model = helpers.build_mnist_cnn(nn)
        model.gpu()
        opt = optim.SGD(model.parameters(), lr=0.01, momentum=0)
        crit = nn.NLLLoss()

        def step(xb, yb):
            out = model(xb)
            opt.zero_grad()
            loss = crit(out, yb)
            loss.backward()
            opt.step()
            return loss
        return step

    rng = np.random.RandomState(0)
    X = rng.rand(8, 128, 784).astype(np.float32)
    Y = rng.randint(0, 10, (8, 128)).astype(np.float32)

    # ---- eager
    step = setup()
    xb, yb = Tensor(X[0], device=Device.GPU), Tensor(Y[0], device=Device.GPU)
    for _ in range(warmup):
        step(xb, yb)
    B.sync()
    gc.collect()  # garbage of the previous measurements would otherwise be collected inside the timed loop
    launches0 = B.launch_count()
    e0, e1 = B.Event().record(), B.Event()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step(xb, yb)
    host_ms = (time.perf_counter() - t0) * 1e3 / steps
    e1.record()
    e1.synchronize()
    eager_ms = e0.elapsed_ms(e1) / steps
    eager_launches = (B.launch_count() - launches0) / steps
    eager_loss = float(loss.numpy()[0])

    # ---- graphed
    fast = GraphedStep(setup(), X[0], Y[0], warmup=3)
    for i in range(warmup):
        fast(X[i % 8], Y[i % 8])
    B.sync()
    gc.collect()
    e0, e1 = B.Event().record(), B.Event()
    t0 = time.perf_counter()
    for i in range(steps):
        out = fast(X[i % 8], Y[i % 8])
    ghost_ms = (time.perf_counter() - t0) * 1e3 / steps
    e1.record()
    e1.synchronize()
    graph_ms = max(e0.elapsed_ms(e1), (time.perf_counter() - t0) * 1e3) / steps
    graph_loss = float(out.numpy()[0])
    res = {"workload": "README MNIST CNN, BS 128, SGD (BASELINE configs[0])",
           "images_per_sec": round(128 / (graph_ms * 1e-3), 1), "ms_per_step": round(graph_ms, 4),
           "mode": "captured step (CUDA graph): per step one pinned H2D of a new host batch + one cudaGraphLaunch",
           "kernels_per_step": fast.kernel_nodes, "host_ms_per_step": round(ghost_ms, 4), "loss": graph_loss,
           "eager": {"images_per_sec": round(128 / (eager_ms * 1e-3), 1), "ms_per_step": round(eager_ms, 4),
                     "launches_per_step": eager_launches, "host_ms_per_step": round(host_ms, 4),
                     "host_us_per_launch": round(host_ms * 1e3 / max(eager_launches, 1), 2), "loss": eager_loss}}
    fast.close()
    return res


def per_gpu_batch(args, world):
    if args.scaling == "strong":
        if STRONG_GLOBAL_BATCH % world:
            raise SystemExit(f"--scaling strong needs a world size that divides {STRONG_GLOBAL_BATCH}")
        return STRONG_GLOBAL_BATCH // world
    return PER_GPU_BATCH


def measure_training(args, B, dist, lib, math_mode, with_clocks):
    """W warm-up + K timed training steps of the workload in one math mode (backend.set_math_mode).
    Returns the device-resident throughput, the end-to-end throughput through the public API with
    host batches, the per-class kernel timings and (fp32 only) the measured FFMA peak."""
    import nanograd_b200.nn.module as nn
    import nanograd_b200.optim.optimizer as optim
    from nanograd_b200.device import Device
    from nanograd_b200.tensor import Tensor

    B.set_math_mode(math_mode)
    rank, world = dist.rank(), dist.world_size()
    batch = per_gpu_batch(args, world)
    # identical parameters on every rank: host-side NumPy init under the same seed
    np.random.seed(0)
    model = build_model(nn)
    model.gpu()
    opt = optim.Adam(model.parameters(), lr=LR)
    crit = nn.NLLLoss()
    n_params = sum(int(np.prod(p.shape)) for p in model.parameters())

    # this rank's shard of the global synthetic batch
    Xh, Yh = synthetic_batch(batch, seed=1000 + rank)
    x_res, y_res = Tensor(Xh, device=Device.GPU), Tensor(Yh, device=Device.GPU)

    W, K = max(args.warmup, 3), max(args.steps, 1)

    # ---- value: inputs resident in HBM -------------------------------------------------------
    # the clock sampler starts before the warm-up steps (nvidia-smi needs ~0.1 s to produce its first
    # line; the timed region of a short run would be over by then) and stops after the timed steps
    clocks = ClockSampler(dist.local_rank()) if with_clocks else None
    if clocks:
        clocks.__enter__()
    for _ in range(W):
        loss = train_step(model, opt, crit, x_res, y_res)
    float(loss.numpy()[0])
    ffma_peak = ctypes.c_float(0)
    lib.ng_measure_ffma_peak(4096, ctypes.byref(ffma_peak))
    import gc
    gc.collect()
    gc.disable()  # a generation-2 collection inside a 10-step timed loop costs more than a step
    dist.barrier()
    B.sync()
    launches0 = B.launch_count()
    e0, e1 = B.Event().record(), B.Event()
    h0 = time.perf_counter()
    for _ in range(K):
        loss = train_step(model, opt, crit, x_res, y_res)
    host_ms_per_step = (time.perf_counter() - h0) * 1e3 / K   # Python + ctypes + launch time (no sync inside)
    e1.record()
    e1.synchronize()
    B.sync()
    dist.barrier()
    ms_total = dist.all_reduce_max(e0.elapsed_ms(e1))
    launches = B.launch_count() - launches0
    loss_value = float(loss.numpy()[0])
    # per-class kernel timings: the same K steps once more with the library's CUDA-event brackets enabled (about
    # 45 bracketed calls per step: recording them costs ~0.15 ms per step, so they stay out of the timed region
    # above; each class's own time is what the roofline objects use)
    lib.ng_prof_reset()
    lib.ng_prof_enable(1)
    p0 = B.Event().record()
    for _ in range(K):
        loss = train_step(model, opt, crit, x_res, y_res)
    p1 = B.Event().record()
    p1.synchronize()
    lib.ng_prof_enable(0)
    profiled_ms_per_step = p0.elapsed_ms(p1) / K
    if clocks:
        clocks.__exit__(None, None, None)
    prof = {name: prof_read(lib, cls) for name, cls in
            (("conv_fprop", 0), ("conv_dgrad", 1), ("conv_wgrad", 2), ("gemm", 3),
             ("staging", 4), ("batchnorm", 5), ("pooling", 6), ("optimizer", 7))}
    lib.ng_prof_reset()
    ms_per_step = ms_total / K
    value = batch * world / (ms_per_step * 1e-3)

    # ---- e2e: the public input pipeline (nanograd_b200.data.BatchLoader) -----------------------
    # Every step draws a random mini-batch from a host-resident synthetic dataset (the reference's loop,
    # tests/helpers.py:294-299), gathers it into pinned staging memory, copies it host->device on the
    # copy stream (overlapping the previous step) and reads the loss back to the host (blocking).
    from nanograd_b200.data import BatchLoader
    rng = np.random.RandomState(2000 + rank)
    n_data = 4 * batch
    Xd = rng.randn(n_data, 3, 32, 32).astype(np.float32)
    Yd = rng.randint(0, 10, n_data).astype(np.float32)
    # one loader for 2 untimed + K timed steps (its pinned staging and thread exist before the clock starts)
    e0, e1, t0 = None, B.Event(), 0.0
    for i, (xb, yb) in enumerate(BatchLoader(Xd, Yd, batch, 2 + K, device=Device.GPU)):
        if i == 2:
            dist.barrier()
            B.sync()
            e0 = B.Event().record()
            t0 = time.perf_counter()
        loss = train_step(model, opt, crit, xb, yb)
        float(loss.numpy()[0])                  # D2H of the step's result (blocking)
    e1.record()
    e1.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dist.barrier()
    gc.enable()
    e2e_ms = dist.all_reduce_max(max(e0.elapsed_ms(e1), wall_ms)) / K
    e2e_value = batch * world / (e2e_ms * 1e-3)
    checksum = dist.replica_checksum(list(model.parameters())) if world > 1 else None
    del loss, x_res, y_res, model, opt
    B.set_math_mode("fp32")
    return {"value": value, "ms_per_step": ms_per_step, "e2e_value": e2e_value, "e2e_ms": e2e_ms,
            "h2d": int(Xh.nbytes + Yh.nbytes), "launches": int(launches), "loss": loss_value, "prof": prof,
            "profiled_ms_per_step": profiled_ms_per_step, "host_ms_per_step": host_ms_per_step,
            "ffma_peak": float(ffma_peak.value), "clocks": clocks.summary() if clocks else None,
            "n_params": n_params, "K": K, "W": W, "batch": batch, "replica_checksum": checksum}


KERNEL_NAMES = {
    "simt": {"conv_fprop": "conv_gather_kernel<CtTile<2,4,8>,TAPMAJOR> (fprop)",
             "conv_dgrad": "conv_gather_kernel<CtTile<2,4,8>,TAPMAJOR> (dgrad)",
             "conv_wgrad": "conv_wgrad_kernel<CtTile<2,4,8>> + splitk_reduce_kernel", "gemm": "gemm_kernel"},
    "tf32": {"conv_fprop": "nchw_to_nhwc + filter_relayout + umma_conv_gather_kernel (tcgen05, fprop)",
             "conv_dgrad": "nchw_to_nhwc + filter_relayout + umma_conv_gather_kernel (tcgen05, dgrad)",
             "conv_wgrad": "2x nchw_to_nhwc + umma_conv_wgrad_kernel<false> (tcgen05) + wgrad_reduce_kernel",
             "gemm": "gemm_kernel (FP32 SIMT)"},
    "tf32x3": {"conv_fprop": "nchw_to_nhwc + filter_relayout + umma_conv_gather_kernel (tcgen05 3xTF32, fprop)",
               "conv_dgrad": "nchw_to_nhwc + filter_relayout + umma_conv_gather_kernel (tcgen05 3xTF32, dgrad)",
               "conv_wgrad": "2x nchw_to_nhwc + umma_conv_wgrad_kernel<true> (tcgen05 3xTF32) + wgrad_reduce_kernel",
               "gemm": "gemm_kernel (FP32 SIMT)"},
    "fp32": {"conv_fprop": "umma_conv_gather_f16_kernel (tcgen05 FP16x3, fprop; x and both filter layouts staged once per layer; the "
                           "class also holds the first layer's conv_fprop_smallc_kernel and the filter staging launches)",
             "conv_dgrad": "umma_conv_gather_f16_kernel (tcgen05 FP16x3, dgrad; dy staged once per layer)",
             "conv_wgrad": "umma_conv_wgrad_f16_kernel (tcgen05 FP16x3) + wgrad_reduce_scaled_kernel",
             "gemm": "umma_gemm_kernel (3xTF32) / gemm_small_kernel (FP32 SIMT, N = 10 head)"},
}
KERNEL_NAMES["fp32_tf32x3"] = KERNEL_NAMES["tf32x3"]
KERNEL_NAMES["f16x3"] = KERNEL_NAMES["fp32"]


def pick_peak(peaks, clocks):
    """Burst or sustained tensor peak?  MEASURED_PEAKS.json holds the dense bf16 cuBLAS rate measured
    best-of-10 (burst: SM clock at its maximum) and back-to-back for 4 s (sustained: the clock has
    dropped to ~1.3 GHz under the power cap).  The figure that applies is the one whose clock matches
    the SM clock this run actually held during the timed region (nvidia-smi samples)."""
    burst = float(peaks.get("bf16_tflops", 0.0))
    sustained = float(peaks.get("bf16_tflops_sustained", burst))
    sm, mx = (clocks or {}).get("sm_mhz"), (clocks or {}).get("sm_max_mhz")
    if sm and mx and sm < 0.85 * mx:
        return sustained, "sustained", burst, sustained
    return burst, "burst", burst, sustained


def roofline_object(m, math_mode, peaks, peak_src, clocks=None):
    """Roofline of the dominant contraction class.
    simt: the FFMA pipe bounds it (peak = FFMA-saturating microkernel measured in this run).
    tf32: the tensor pipe bounds it; dense TF32 peak = 1/2 of the measured dense BF16 GEMM rate of
          MEASURED_PEAKS.json — the burst figure when the SM clock held its maximum over the timed
          region, the sustained one when it did not (pick_peak).
    fp32 / tf32x3 (FP32-accurate 3xTF32 on tensor cores): three TF32 MMAs per useful FP32 MAC, so the
          peak for useful FLOPs is the TF32 peak / 3.  Both fractions and the raw tensor-pipe fraction
          (MMA FLOPs issued / dense TF32 peak) are printed next to `frac`."""
    prof, K = m["prof"], m["K"]
    contraction = {k: v for k, v in prof.items() if k in ("conv_fprop", "conv_dgrad", "conv_wgrad", "gemm")}
    kernel_ms = sum(v["ms"] for v in contraction.values())
    # The dominant KERNEL: on the tensor-core engines fprop and dgrad are the same kernel function (the implicit-GEMM
    # "gather" kernel, 10 of its launches per step), so the two event-bracketed classes are one roofline candidate;
    # wgrad and GEMM are their own kernels.  (SIMT mode: one conv_gather_kernel serves both as well.)
    merged = {"conv_gather": {k: prof["conv_fprop"][k] + prof["conv_dgrad"][k] for k in ("ms", "launches", "flops", "bytes")},
              "conv_wgrad": prof["conv_wgrad"], "gemm": prof["gemm"]}
    dominant = max(merged, key=lambda k: merged[k]["ms"])
    d = merged[dominant]
    achieved = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
    bf16, which, burst, sustained = pick_peak(peaks, clocks)
    extra = {}
    if math_mode == "simt":
        bound, peak = "fp32_ffma", m["ffma_peak"]
        src = ("FFMA-saturating microkernel measured in this run (nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4); "
               "no FP32 tensor-core MMA exists, so the tensor peak of %s does not bound this mode" % peak_src)
    else:
        mma_per_mac = 1.0 if math_mode == "tf32" else 3.0
        # dense rate of the MMA kind the mode issues: kind::f16 (the measured bf16 GEMM rate) for FP16x3,
        # kind::tf32 (half of it) for TF32 / 3xTF32
        f16 = math_mode in ("fp32", "f16x3")
        dense_frac = 1.0 if f16 else 0.5
        bound, peak = "tensor", dense_frac * bf16 / mma_per_mac
        src = ("%s = dense %s (%sthe measured %s dense bf16 GEMM rate in %s, %.1f TFLOP/s: the SM clock held "
               "%s of %s MHz over the timed region)%s" % (
                   "dense TF32 peak" if mma_per_mac == 1 else
                   "useful-FLOP peak of the FP32-accurate %s scheme" % ("FP16x3" if f16 else "3xTF32"),
                   "FP16" if f16 else "TF32", "" if f16 else "1/2 of ",
                   which, peak_src, bf16, (clocks or {}).get("sm_mhz"), (clocks or {}).get("sm_max_mhz"),
                   "" if mma_per_mac == 1 else " / 3 MMAs per FP32 MAC; for scale, the measured FFMA peak of the SIMT "
                   "alternative is %.1f TFLOP/s" % m["ffma_peak"]))
        extra = {"frac_vs_burst_peak": round(achieved / (dense_frac * burst / mma_per_mac), 4) if burst else None,
                 "frac_vs_sustained_peak": round(achieved / (dense_frac * sustained / mma_per_mac), 4) if sustained else None,
                 "tensor_pipe_frac_of_dense": round(achieved * mma_per_mac / (dense_frac * bf16), 4) if bf16 else None,
                 "useful_frac_of_dense_tf32": round(achieved / (0.5 * bf16), 4) if bf16 else None,
                 "peak_kind": which}
    out = {
        "kernel": (KERNEL_NAMES[math_mode]["conv_fprop"] + " | " + KERNEL_NAMES[math_mode]["conv_dgrad"]
                   if dominant == "conv_gather" else KERNEL_NAMES[math_mode][dominant]),
        "dominant": dominant + (" = the conv_fprop + conv_dgrad classes of per_class (one kernel function)"
                                if dominant == "conv_gather" else ""),
        "bound": bound, "achieved": round(achieved, 2),
        "peak": round(peak, 2), "unit": "TFLOP/s", "frac": round(achieved / peak, 4) if peak else None,
        "traffic": ncu_traffic(math_mode + ":" + ("conv_fprop" if dominant == "conv_gather" else dominant)), "peak_source": src,
        "avg_launch_ms": round(d["ms"] / max(d["launches"], 1), 4), "launches": d["launches"],
        "class_flops_per_step": d["flops"] / K, "class_ms_per_step": round(d["ms"] / K, 4),
        "share_of_step": round(d["ms"] / (m["ms_per_step"] * K), 4),
        "contraction_share_of_step": round(kernel_ms / (m["ms_per_step"] * K), 4),
        "class_timing": "CUDA events around each library call of the class, K steps run right after the timed region "
                        "(%.4f ms/step with the brackets recording, %.4f without)" % (
                            m.get("profiled_ms_per_step", 0.0), m["ms_per_step"]),
        "per_class": {k: {"ms_per_step": round(v["ms"] / K, 4), "launches_per_step": v["launches"] / K,
                          "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 and v["flops"] > 0 else None,
                          "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 and v["bytes"] > 0 else None}
                      for k, v in prof.items() if v["launches"]},
        "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_peak_source": peak_src,
    }
    out.update(extra)
    # the memory-bound classes of the step against the measured HBM copy bandwidth: algorithmic bytes (DESIGN.md
    # 3: what a pass must read and write once) / CUDA-event time of the calls of the class
    hbm = float(peaks.get("hbm_gbs") or 0.0)
    mem = {}
    for k in ("staging", "batchnorm", "pooling", "optimizer"):
        v = prof.get(k)
        if not v or not v["launches"] or v["ms"] <= 0:
            continue
        gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9
        mem[k] = {"bound": "hbm", "achieved": round(gbs, 1), "peak": hbm, "unit": "GB/s",
                  "frac": round(gbs / hbm, 4) if hbm else None, "ms_per_step": round(v["ms"] / K, 4),
                  "calls_per_step": v["launches"] / K, "algorithmic_bytes_per_step": v["bytes"] / K,
                  "share_of_step": round(v["ms"] / (m["ms_per_step"] * K), 4)}
    if mem:
        out["memory_bound"] = mem
        out["memory_bound_note"] = ("staging = channels-last FP16-pair copies of the tensor-core operands with the "
                                    "BatchNorm normalise / dx passes fused in where the tensor has no other reader "
                                    "(8 B/elt, 12 B/elt for the fused backward); batchnorm = statistics 4 B/elt, "
                                    "normalise 8, backward sums 8, dx 12; pooling 4(1+1/p^2) fwd, 4(2+2/p^2) bwd; "
                                    "Adam 28 B/param")
    out["per_class"] = {k: v for k, v in out["per_class"].items()
                        if k in ("conv_fprop", "conv_dgrad", "conv_wgrad", "gemm")}
    return out


def attention_extra(B, lib, peaks):
    """BASELINE configs[3]: one self-attention block (matmul + softmax heavy), seq 512, d 256, MSE loss, AdamW,
    composed from the reference's primitives like tests/cases.py:case_attention (examples/attention.py is a stub
    in the reference).  Reports steps/s and the GEMM / softmax classes of the step."""
    import nanograd_b200.nn.module as nn
    import nanograd_b200.optim.optimizer as optim
    from nanograd_b200.device import Device
    from nanograd_b200.tensor import Tensor
    seq, d = 512, 256
    np.random.seed(0)

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

    model = Attn()
    model.gpu()
    opt = optim.AdamW(model.parameters(), lr=1e-3)
    crit = nn.MSELoss()
    x = Tensor(np.random.randn(seq, d).astype(np.float32), device=Device.GPU)
    t = Tensor(np.random.randn(seq, d).astype(np.float32), device=Device.GPU)

    def step():
        out = model(x)
        opt.zero_grad()
        loss = crit(out, t)
        loss.backward()
        opt.step()
        return loss

    for _ in range(5):
        loss = step()
    float(loss.numpy()[0])
    lib.ng_prof_reset()
    lib.ng_prof_enable(1)
    K = 30
    l0 = B.launch_count()
    e0, e1 = B.Event().record(), B.Event()
    t0 = time.perf_counter()
    for _ in range(K):
        loss = step()
    host_ms = (time.perf_counter() - t0) * 1e3 / K
    e1.record()
    e1.synchronize()
    lib.ng_prof_enable(0)
    ms = e0.elapsed_ms(e1) / K
    gemm = prof_read(lib, 3)
    lib.ng_prof_reset()
    # forward GEMMs: 3 projections (2 s d^2 each), q k^T (2 s^2 d), p v (2 s^2 d), out projection (2 s d^2);
    # the backward doubles each (two products per matmul)
    flops = 3.0 * (4 * 2.0 * seq * d * d + 2 * 2.0 * seq * seq * d)
    return {"workload": "self-attention block, seq 512, d 256, MSE loss, AdamW (BASELINE configs[3])",
            "steps_per_sec": round(1e3 / ms, 1), "ms_per_step": round(ms, 4), "host_ms_per_step": round(host_ms, 4),
            "launches_per_step": (B.launch_count() - l0) / K, "loss": float(loss.numpy()[0]),
            "gemm": {"calls_per_step": gemm["launches"] / K, "ms_per_step": round(gemm["ms"] / K, 4),
                     "tflops": round(gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12, 2) if gemm["ms"] > 0 else None,
                     "algorithmic_gflop_per_step": round(flops / 1e9, 3)},
            "note": "1.6 GFLOP and a few MB per step in 15 GEMM calls of 67-134 MFLOP (3xTF32 tcgen05 GEMM with "
                    "split-K) plus ~20 element-wise / softmax launches: the step is launch- and host-dispatch-bound "
                    "(compare host_ms_per_step), not roofline-bound"}


def run_device_arm(args):
    from nanograd_b200 import backend as B
    from nanograd_b200 import dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun: python -m torch.distributed.run --nnodes=1 "
                             f"--nproc-per-node {args.gpus} --master-addr 127.0.0.1 bench.py --gpus {args.gpus}")
    dist.init(use_device=True, sync_bn=args.sync_bn)
    rank, world = dist.rank(), dist.world_size()
    lib = B.lib()
    B.init(dist.local_rank())
    peaks, peak_src = measured_peaks()

    main_mode = args.math
    m = measure_training(args, B, dist, lib, main_mode, with_clocks=True)
    others = {}
    if args.math == "fp32" and not args.no_tf32:
        others["simt_mode"] = ("simt", measure_training(args, B, dist, lib, "simt", with_clocks=False))
        others["tf32_mode"] = ("tf32", measure_training(args, B, dist, lib, "tf32", with_clocks=False))

    if rank != 0:
        dist.shutdown()
        return

    K, W = m["K"], m["W"]
    notes = {"simt_mode": "same workload restricted to the FP32 SIMT (FFMA) kernels (NANOGRAD_B200_MATH=simt)",
             "tf32_mode": "same workload with the opt-in plain-TF32 tensor-core convolutions (NANOGRAD_B200_MATH=tf32; "
                          "parity bar rtol 1e-2); reduced precision, hence not the headline value"}
    line = {
        "metric": "train_images_per_sec", "value": round(m["value"], 1), "unit": "images/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": round(m["ms_per_step"], 4), "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None, "dtype": "tf32" if main_mode == "tf32" else "f32", "data": "synthetic",
        "config": workload_config(world, m["n_params"], main_mode, m["batch"], args),
        "e2e": {"value": round(m["e2e_value"], 1), "unit": "images/s", "ms_per_step": round(m["e2e_ms"], 4),
                "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": 4},
        "gpu_launches": m["launches"], "launches_per_step": m["launches"] / K, "loss": m["loss"],
        "host_ms_per_step": round(m["host_ms_per_step"], 4),
        "replica_checksum": m["replica_checksum"],
        "clocks": m["clocks"], "roofline": roofline_object(m, main_mode, peaks, peak_src, m["clocks"]),
    }
    for key, (mode, o) in others.items():
        line[key] = {
            "note": notes[key], "value": round(o["value"], 1), "unit": "images/s", "ms_per_step": round(o["ms_per_step"], 4),
            "e2e": {"value": round(o["e2e_value"], 1), "unit": "images/s", "ms_per_step": round(o["e2e_ms"], 4),
                    "h2d_bytes_per_step": o["h2d"], "d2h_bytes_per_step": 4},
            "gpu_launches": o["launches"], "loss": o["loss"], "roofline": roofline_object(o, mode, peaks, peak_src, m["clocks"])}

    if world == 1 and not args.no_extras:
        ffma = m["ffma_peak"]

        def extra(key, fn):
            # the headline measurement is complete at this point: a failing side measurement must not cost the line
            try:
                line[key] = fn()
            except Exception as e:  # noqa: BLE001 - reported in the line, never swallowed silently
                line[key] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

        extra("conv2d_sweep", lambda: {
            "workload": "Conv2d 3x3 s1 p0, x(256,C,32,32), BASELINE configs[1]",
            "columns": "<engine>_<pass>_tflops: one C-ABI call per pass on NCHW operands, every call "
                       "staging its own tensor-core operands; f16x3_layer_*: operands staged once per "
                       "layer and shared by the passes, as a training step does (staging time included "
                       "in f16x3_layer_fwd_bwd_tflops)",
            "ffma_peak_tflops": round(ffma, 2),
            "rows": conv2d_sweep(B, lib, ffma)})
        extra("mnist_cnn", lambda: mnist_cnn_images_per_sec(B))
        extra("attention", lambda: attention_extra(B, lib, peaks))

        def cpu_baseline():
            t0 = time.perf_counter()
            cpu_ips, cpu_sec, kind, what = cpu_arm(steps=3, warmup=1)
            return {"value": round(cpu_ips, 2), "unit": "images/s", "cores": host_threads(), "kind": kind,
                    "sample": f"{what}; 3 training steps (+1 warm-up) on {CPU_SAMPLE_BATCH} of the workload's {PER_GPU_BATCH} "
                              f"images, same model/optimizer, {cpu_sec:.2f} s per step, {time.perf_counter() - t0:.1f} s total"}

        extra("cpu_baseline", cpu_baseline)
    print(json.dumps(line), flush=True)
    dist.shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="device", choices=["device", "reference"])
    ap.add_argument("--math", default="fp32", choices=["fp32", "simt", "tf32x3", "tf32", "fp32_tf32x3", "f16x3"],
                    help="math mode of the headline measurement (default fp32: FP32-accurate, engine picked per call)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 512 images per GPU; strong: global batch 4096 split over the GPUs (SURVEY 8-d cfg4)")
    ap.add_argument("--sync-bn", action="store_true",
                    help="BatchNorm statistics over the global batch (SyncBN) instead of per shard")
    ap.add_argument("--no-tf32", action="store_true", help="skip the secondary simt-mode and tf32-mode measurements")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the conv2d sweep / MNIST CNN / CPU baseline side measurements (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_device_arm(args)


if __name__ == "__main__":
    main()
