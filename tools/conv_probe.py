"""Time the contraction kernels on BASELINE config 1 shapes (Conv2d 3x3, 32x32, BS 256) and a few
GEMMs with CUDA events; prints TFLOP/s per pass.  Usage: python tools/conv_probe.py [--sizes 16,64,256]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanograd_b200 import backend as B  # noqa: E402
from nanograd_b200.nn.buffer import DeviceBuffer  # noqa: E402


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="16,32,64,128,256")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--hw", type=int, default=32)
    ap.add_argument("--pad", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    B.init()
    lib = B.lib()
    import ctypes
    peak = ctypes.c_float(0)
    lib.ng_measure_ffma_peak(4096, ctypes.byref(peak))
    print(json.dumps({"ffma_peak_tflops": round(peak.value, 2), **B.device_info()}))
    N, H, W = args.batch, args.hw, args.hw
    R = S = 3
    pad = args.pad
    OH, OW = H + 2 * pad - R + 1, W + 2 * pad - S + 1
    rng = np.random.RandomState(0)
    for ch in [int(s) for s in args.sizes.split(",")]:
        C = K = ch
        x = DeviceBuffer((N, C, H, W), hostbuf=rng.randn(N, C, H, W).astype(np.float32))
        w = DeviceBuffer((K, C, R, S), hostbuf=(rng.randn(K, C, R, S) / np.sqrt(9 * C)).astype(np.float32))
        dy = DeviceBuffer((N, K, OH, OW), hostbuf=rng.randn(N, K, OH, OW).astype(np.float32))
        y = DeviceBuffer((N, K, OH, OW))
        dx = DeviceBuffer((N, C, H, W))
        dw = DeviceBuffer((K, C, R, S))
        flops = 2.0 * N * K * C * R * S * OH * OW
        dims = (N, C, H, W, K, R, S, 1, pad, pad, OH, OW, 0)
        t_f = time_call(lambda: lib.ng_conv2d_fprop(y.ptr, x.ptr, w.ptr, 0, *dims), args.iters)
        t_d = time_call(lambda: lib.ng_conv2d_dgrad(dx.ptr, dy.ptr, w.ptr, *dims), args.iters)
        t_w = time_call(lambda: lib.ng_conv2d_wgrad(dw.ptr, dy.ptr, x.ptr, *dims), args.iters)
        tf = lambda ms: round(flops / (ms * 1e-3) / 1e12, 2)
        print(json.dumps({"C=K": ch, "fprop_ms": round(t_f, 4), "dgrad_ms": round(t_d, 4), "wgrad_ms": round(t_w, 4),
                          "fprop_tflops": tf(t_f), "dgrad_tflops": tf(t_d), "wgrad_tflops": tf(t_w),
                          "fwd_bwd_tflops": round(3 * flops / ((t_f + t_d + t_w) * 1e-3) / 1e12, 2)}), flush=True)
        del x, w, dy, y, dx, dw
    for (M, Nn, Kk) in [(512, 256, 4096), (4096, 4096, 4096), (512, 10, 256)]:
        a = DeviceBuffer((M, Kk), hostbuf=rng.randn(M, Kk).astype(np.float32))
        b = DeviceBuffer((Nn, Kk), hostbuf=rng.randn(Nn, Kk).astype(np.float32))
        c = DeviceBuffer((M, Nn))
        t = time_call(lambda: lib.ng_gemm(c.ptr, a.ptr, b.ptr, M, Nn, Kk, 0, 1, 1, 0, 0), args.iters)
        print(json.dumps({"gemm_nt": [M, Nn, Kk], "ms": round(t, 4), "tflops": round(2.0 * M * Nn * Kk / (t * 1e-3) / 1e12, 2)}),
              flush=True)


if __name__ == "__main__":
    main()
