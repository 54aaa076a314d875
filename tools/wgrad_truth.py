"""Accuracy of the wgrad paths against a float64 NumPy reference at a BASELINE-size layer."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanograd_b200 import backend as B
from nanograd_b200.nn.buffer import DeviceBuffer
B.init()
lib = B.lib()
N, C, H, W, K, R, S, pad = 512, 64, 32, 32, 64, 3, 3, 1
rng = np.random.RandomState(0)
x = rng.randn(N, C, H, W).astype(np.float32)
dy = rng.randn(N, K, H, W).astype(np.float32)
t0 = time.time()
xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (1, 1), (1, 1)))
dy64 = dy.astype(np.float64)
ref = np.zeros((K, C, R, S))
for r in range(R):
    for s in range(S):
        ref[:, :, r, s] = np.tensordot(dy64, xp[:, :, r:r + H, s:s + W], axes=([0, 2, 3], [0, 2, 3]))
print("fp64 reference in %.1f s" % (time.time() - t0))
xd, dyd = DeviceBuffer(x.shape, hostbuf=x), DeviceBuffer(dy.shape, hostbuf=dy)
dims = (N, C, H, W, K, R, S, 1, pad, pad, H, W)
for name, flag in (("fp32 simt", 0), ("tf32", B.F_TF32), ("tf32x3", B.F_TF32X3 | B.F_TC_FORCE)):
    dw = DeviceBuffer((K, C, R, S))
    lib.ng_conv2d_wgrad(dw.ptr, dyd.ptr, xd.ptr, *dims, flag)
    got = dw.numpy().astype(np.float64)
    print(name, "max|err|/max|ref| = %.3e" % (np.abs(got - ref).max() / np.abs(ref).max()))
