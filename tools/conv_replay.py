"""Record the (dy, x, w) operands of every conv backward of the vgg_w64 test net, then replay fprop / dgrad /
wgrad through the C ABI in every engine and compare with a float64 evaluation (diagnostic)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import helpers as H  # noqa: E402
from nanograd_b200 import backend as B  # noqa: E402
from nanograd_b200.device import Device  # noqa: E402
from nanograd_b200.nn import ops_cuda  # noqa: E402
from nanograd_b200.nn.buffer import DeviceBuffer  # noqa: E402


def truth(x, w, dy, pad):
    n, c, h, wd = x.shape
    k, _, r, s = w.shape
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    dyp = np.pad(dy.astype(np.float64), ((0, 0), (0, 0), (r - 1 - pad, r - 1 - pad), (s - 1 - pad, s - 1 - pad)))
    w64 = w.astype(np.float64)
    y = np.zeros(dy.shape)
    dx = np.zeros(x.shape)
    dw = np.zeros(w.shape)
    oh, ow = dy.shape[2:]
    for i in range(r):
        for j in range(s):
            y += np.einsum("nchw,kc->nkhw", xp[:, :, i:i + oh, j:j + ow], w64[:, :, i, j])
            dx += np.einsum("nkhw,kc->nchw", dyp[:, :, r - 1 - i:r - 1 - i + h, s - 1 - j:s - 1 - j + wd], w64[:, :, i, j])
            dw[:, :, i, j] = np.einsum("nkhw,nchw->kc", dy.astype(np.float64), xp[:, :, i:i + oh, j:j + ow])
    return y, dx, dw


def main():
    H.use_cuda_library()
    B.set_math_mode("simt")
    rec = []
    orig = ops_cuda._conv_backward_shared

    def spy(ctx, grad_output, input, weight, stride, padding, want_db=False):
        rec.append((grad_output.numpy().copy(), input.numpy().copy(), weight.numpy().copy(), padding))
        return orig(ctx, grad_output, input, weight, stride, padding, want_db=want_db)

    ops_cuda._conv_backward_shared = spy
    fn, kwargs = cases.all_cases()[sys.argv[1] if len(sys.argv) > 1 else "vgg_w64_sgd"]
    kwargs = dict(kwargs, steps=1)
    fn(H.device_framework(Device.GPU), **kwargs)
    ops_cuda._conv_backward_shared = orig
    lib = B.lib()
    for dy, x, w, padding in rec:
        n, c, h, wd = x.shape
        k, _, r, s = w.shape
        if c % 64 or k % 64:
            continue
        pad = padding[0]
        oh, ow = dy.shape[2:]
        ty, tdx, tdw = truth(x, w, dy, pad)
        xd, wd_, dyd = DeviceBuffer(x.shape, hostbuf=x), DeviceBuffer(w.shape, hostbuf=w), DeviceBuffer(dy.shape, hostbuf=dy)
        dims = (n, c, h, wd, k, r, s, 1, pad, pad, oh, ow)
        print(f"conv {c}->{k} {h}x{wd} n{n}: max|x| {np.abs(x).max():.2e} max|dy| {np.abs(dy).max():.2e} "
              f"min nonzero |dy| {np.abs(dy[dy != 0]).min():.2e} zeros in dy {(dy == 0).mean():.3f} x {(x == 0).mean():.3f}")
        for name, flag in (("simt", 0), ("tf32x3", B.F_TF32X3 | B.F_TC_FORCE), ("f16x3", B.F_F16X3 | B.F_TC_FORCE)):
            yo, dxo, dwo = DeviceBuffer(dy.shape), DeviceBuffer(x.shape), DeviceBuffer(w.shape)
            lib.ng_conv2d_fprop(yo.ptr, xd.ptr, wd_.ptr, 0, *dims, flag)
            lib.ng_conv2d_dgrad(dxo.ptr, dyd.ptr, wd_.ptr, *dims, flag)
            lib.ng_conv2d_wgrad(dwo.ptr, dyd.ptr, xd.ptr, *dims, flag)
            e = [float(np.abs(a.numpy() - t).max() / np.abs(t).max()) for a, t in ((yo, ty), (dxo, tdx), (dwo, tdw))]
            print(f"   {name:7s} fprop {e[0]:.1e} dgrad {e[1]:.1e} wgrad {e[2]:.1e}", flush=True)


if __name__ == "__main__":
    main()
