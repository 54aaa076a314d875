"""Tensor-core GEMM (tcgen05) vs the FP32 SIMT GEMM through ng_gemm, all four transpose combinations."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanograd_b200 import backend as B
from nanograd_b200.nn.buffer import DeviceBuffer
B.init(); lib = B.lib()
def t(fn, iters=20):
    for _ in range(3): fn()
    e0, e1 = B.Event().record(), B.Event()
    for _ in range(iters): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_ms(e1) / iters
perf = "--perf" in sys.argv
shapes = [(512, 256, 256), (512, 512, 256), (512, 256, 512), (256, 256, 512), (512, 256, 4096), (512, 4096, 256),
          (256, 4096, 512), (128, 64, 96), (37 * 4, 96, 132), (512, 10, 256),
          (512, 512, 512), (4096, 4096, 4096)] if perf else \
         [(16, 32, 2048), (16, 2048, 32), (32, 2048, 16), (16, 10, 32), (16, 32, 1024), (16, 1024, 32), (32, 1024, 16), (512, 256, 4096), (512, 4096, 256), (256, 4096, 512), (128, 64, 96), (148, 96, 132), (512, 10, 256), (64, 32, 32), (300, 160, 1000)]
rng = np.random.RandomState(0)
for (M, N, K) in shapes:
    for ta in (0, 1):
        for tb in (0, 1):
            a = rng.randn(*((K, M) if ta else (M, K))).astype(np.float32)
            b = rng.randn(*((N, K) if tb else (K, N))).astype(np.float32)
            bias = rng.randn(N).astype(np.float32)
            ad, bd, biasd = DeviceBuffer(a.shape, hostbuf=a), DeviceBuffer(b.shape, hostbuf=b), DeviceBuffer((N,), hostbuf=bias)
            row = {"MNK": [M, N, K], "ta": ta, "tb": tb}
            ref = DeviceBuffer((M, N)); lib.ng_gemm(ref.ptr, ad.ptr, bd.ptr, M, N, K, ta, tb, 1, biasd.ptr, 1)
            r = ref.numpy()
            for name, flag in (("x3", B.F_TF32X3 | B.F_TC_FORCE), ("tf32", B.F_TF32)):
                out = DeviceBuffer((M, N)); lib.ng_memset_zero(out.ptr, M * N * 4)
                lib.ng_gemm(out.ptr, ad.ptr, bd.ptr, M, N, K, ta, tb, 1, biasd.ptr, 1 | flag)
                row[name + "_err"] = round(float(np.abs(out.numpy() - r).max() / max(np.abs(r).max(), 1e-30)), 7)
                if perf:
                    ms = t(lambda: lib.ng_gemm(out.ptr, ad.ptr, bd.ptr, M, N, K, ta, tb, 1, biasd.ptr, 1 | flag))
                    row[name + "_tflops"] = round(2.0 * M * N * K / ms / 1e9, 1)
            if perf:
                ms = t(lambda: lib.ng_gemm(ref.ptr, ad.ptr, bd.ptr, M, N, K, ta, tb, 1, biasd.ptr, 1))
                row["simt_tflops"] = round(2.0 * M * N * K / ms / 1e9, 1)
            print(json.dumps(row), flush=True)
