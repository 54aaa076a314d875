"""Run one convolution pass of one shape a few times (profiling target for ncu).
usage: one_conv.py N C H W K R S pad pass[fprop|dgrad|wgrad] mode[simt|tf32x3|tf32] [iters]"""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from nanograd_b200 import backend as B
from nanograd_b200.nn.buffer import DeviceBuffer

N, C, H, W, K, R, S, pad = map(int, sys.argv[1:9])
which, mode = sys.argv[9], sys.argv[10]
iters = int(sys.argv[11]) if len(sys.argv) > 11 else 3
B.init()
lib = B.lib()
flag = {"simt": 0, "tf32x3": B.F_TF32X3 | B.F_TC_FORCE, "tf32": B.F_TF32}[mode]
OH, OW = H + 2 * pad - R + 1, W + 2 * pad - S + 1
rng = np.random.RandomState(0)
x = DeviceBuffer((N, C, H, W), hostbuf=rng.randn(N, C, H, W).astype(np.float32))
w = DeviceBuffer((K, C, R, S), hostbuf=rng.randn(K, C, R, S).astype(np.float32))
b = DeviceBuffer((K,), hostbuf=rng.randn(K).astype(np.float32))
dy = DeviceBuffer((N, K, OH, OW), hostbuf=rng.randn(N, K, OH, OW).astype(np.float32))
dims = (N, C, H, W, K, R, S, 1, pad, pad, OH, OW)
out = DeviceBuffer({"fprop": (N, K, OH, OW), "dgrad": (N, C, H, W), "wgrad": (K, C, R, S)}[which])
for _ in range(iters):
    if which == "fprop":
        lib.ng_conv2d_fprop(out.ptr, x.ptr, w.ptr, b.ptr, *dims, flag)
    elif which == "dgrad":
        lib.ng_conv2d_dgrad(out.ptr, dy.ptr, w.ptr, *dims, flag)
    else:
        lib.ng_conv2d_wgrad(out.ptr, dy.ptr, x.ptr, *dims, flag)
lib.ng_sync()
print("ok", float(np.abs(out.numpy()).max()))
