"""Compare the tcgen05 TF32 convolution kernels with the FP32 SIMT kernels on the GPU (both go
through the C ABI) for a list of shapes, and time them.  python tools/umma_check.py [--perf]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanograd_b200 import backend as B  # noqa: E402
from nanograd_b200.nn.buffer import DeviceBuffer  # noqa: E402

TF32 = B.F_TF32


def time_call(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    e0, e1 = B.Event(), B.Event()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_ms(e1) / iters


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run(N, C, H, W, K, R, S, pad, perf, passes, pad_h=None):
    lib = B.lib()
    rng = np.random.RandomState(0)
    pad_h = pad if pad_h is None else pad_h
    OH, OW = H + 2 * pad_h - R + 1, W + 2 * pad - S + 1
    x = DeviceBuffer((N, C, H, W), hostbuf=rng.randn(N, C, H, W).astype(np.float32))
    w = DeviceBuffer((K, C, R, S), hostbuf=(rng.randn(K, C, R, S) / np.sqrt(R * S * C)).astype(np.float32))
    b = DeviceBuffer((K,), hostbuf=rng.randn(K).astype(np.float32))
    dy = DeviceBuffer((N, K, OH, OW), hostbuf=rng.randn(N, K, OH, OW).astype(np.float32))
    dims = (N, C, H, W, K, R, S, 1, pad_h, pad, OH, OW)
    flops = 2.0 * N * K * C * R * S * OH * OW
    row = {"shape": f"N{N} C{C} {H}x{W} K{K} {R}x{S} p{pad}"}
    calls = {
        "fprop": ((N, K, OH, OW), lambda o, f: lib.ng_conv2d_fprop(o.ptr, x.ptr, w.ptr, b.ptr, *dims, f)),
        "dgrad": ((N, C, H, W), lambda o, f: lib.ng_conv2d_dgrad(o.ptr, dy.ptr, w.ptr, *dims, f)),
        "wgrad": ((K, C, R, S), lambda o, f: lib.ng_conv2d_wgrad(o.ptr, dy.ptr, x.ptr, *dims, f)),
    }
    for name in passes:
        shape, fn = calls[name]
        o_ref, o_tc = DeviceBuffer(shape), DeviceBuffer(shape)
        lib.ng_memset_zero(o_tc.ptr, int(np.prod(shape)) * 4)
        fn(o_ref, 0)
        fn(o_tc, TF32)
        a, r = o_tc.numpy(), o_ref.numpy()
        row[name + "_relerr"] = round(relerr(a, r), 6)
        if row[name + "_relerr"] > 1e-2:
            bad = np.argwhere(np.abs(a - r) > 1e-2 * np.abs(r).max())
            row[name + "_bad"] = {"count": int(len(bad)), "first": bad[:6].tolist(), "of": int(a.size),
                                  "tc": [float(a[tuple(i)]) for i in bad[:4]], "ref": [float(r[tuple(i)]) for i in bad[:4]]}
        if perf:
            t_ref = time_call(lambda: fn(o_ref, 0))
            t_tc = time_call(lambda: fn(o_tc, TF32))
            row[name + "_simt_tflops"] = round(flops / t_ref / 1e9, 1)
            row[name + "_tf32_tflops"] = round(flops / t_tc / 1e9, 1)
            row[name + "_tf32_ms"] = round(t_tc, 4)
    print(json.dumps(row), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--perf", action="store_true")
    ap.add_argument("--passes", default="fprop,dgrad,wgrad")
    ap.add_argument("--set", default="small")
    ap.add_argument("--x3", action="store_true", help="check the FP32-accurate 3xTF32 mode instead of TF32")
    ap.add_argument("--f16", action="store_true", help="check the FP32-accurate FP16x3 mode (channels % 64 == 0)")
    args = ap.parse_args()
    global TF32
    if args.x3:
        TF32 = B.F_TF32X3 | B.F_TC_FORCE
    if args.f16:
        TF32 = B.F_F16X3 | B.F_TC_FORCE
    B.init()
    passes = args.passes.split(",")
    small = [(2, 32, 32, 32, 32, 3, 3, 1), (2, 64, 32, 32, 64, 3, 3, 1), (4, 64, 16, 16, 128, 3, 3, 1),
             (4, 128, 8, 8, 256, 3, 3, 1), (3, 16, 32, 32, 48, 3, 3, 0), (2, 8, 28, 28, 16, 3, 3, 0),
             (2, 64, 64, 64, 64, 3, 3, 1), (2, 32, 32, 32, 32, 1, 1, 0), (5, 40, 12, 12, 24, 5, 5, 2)]
    odd = [(3, 64, 9, 13, 320, 3, 3, 1), (7, 96, 5, 5, 40, 3, 3, 1), (2, 32, 4, 4, 32, 3, 3, 1),
           (130, 32, 2, 2, 64, 1, 1, 0), (2, 256, 20, 36, 32, 3, 3, 0), (1, 32, 33, 65, 32, 3, 3, 2)]
    vgg = [(512, 64, 32, 32, 64, 3, 3, 1), (512, 64, 16, 16, 128, 3, 3, 1), (512, 128, 16, 16, 128, 3, 3, 1),
           (512, 128, 8, 8, 256, 3, 3, 1), (512, 256, 8, 8, 256, 3, 3, 1)]
    sweep = [(256, c, 32, 32, c, 3, 3, 0) for c in (16, 32, 64, 128, 256)]
    f16 = [(2, 64, 32, 32, 64, 3, 3, 1), (4, 64, 16, 16, 128, 3, 3, 1), (4, 128, 8, 8, 256, 3, 3, 1),
           (2, 64, 64, 64, 64, 3, 3, 1), (3, 64, 9, 13, 320, 3, 3, 1), (2, 256, 20, 36, 64, 3, 3, 0),
           (130, 64, 2, 2, 64, 1, 1, 0), (1, 64, 33, 65, 192, 3, 3, 2), (5, 128, 12, 12, 64, 5, 5, 2),
           (2, 192, 7, 7, 128, 3, 3, 1), (9, 64, 1, 37, 64, 1, 5, 2)]
    w64 = [(16, 64, 16, 16, 64, 3, 3, 1), (16, 64, 8, 8, 128, 3, 3, 1), (16, 128, 8, 8, 128, 3, 3, 1),
           (16, 32, 16, 16, 32, 3, 3, 1), (16, 32, 8, 8, 64, 3, 3, 1), (16, 64, 8, 8, 64, 3, 3, 1)]
    for shp in {"small": small, "vgg": vgg, "sweep": sweep, "odd": odd, "f16": f16, "w64": w64}[args.set]:
        run(*shp, perf=args.perf, passes=passes)
    if args.set == "odd":   # conv1d-shaped problems (H = R = 1), as Conv1d issues them
        run(5, 32, 1, 100, 48, 1, 3, 0, perf=args.perf, passes=passes, pad_h=0)
        run(9, 64, 1, 37, 64, 1, 5, 2, perf=args.perf, passes=passes, pad_h=0)


if __name__ == "__main__":
    main()
