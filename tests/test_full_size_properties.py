"""Parity at BASELINE.json's full sizes (config 2: batch 512, the five tensor-core conv layers and the
3-channel first layer of the VGG; config 2's Linear(4096, 256); 2.2 M-parameter optimizer steps) through
size-independent properties, since the NumPy oracle needs minutes per pass at these sizes:

  * adjointness  <fprop(x, w), dy> = <x, dgrad(dy, w)> = <w, wgrad(dy, x)>   (one identity ties the three
    kernels of a layer together; evaluated in float64 on the host from the device results, with a probe dy
    correlated with y and a tolerance relative to the inner product itself; the test also checks that a
    dropped filter tap fails it).  Elementwise checks at the same sizes: tests/test_full_size_elementwise.py
  * linearity    fprop(a x1 + b x2) = a fprop(x1) + b fprop(x2)
  * pooling      max >= avg elementwise; both backward passes conserve the gradient sum; avg-pool backward
                 of ones is exactly 1 / window
  * BatchNorm    training output has per-channel mean beta and variance gamma^2 var / (var + eps)
  * optimizers   SGD with momentum 0 is bit-exact p - lr g; Adam's first step moves every parameter with a
                 non-zero gradient by lr (bias-corrected m / sqrt(v) = sign(g)) to 1e-4

FP32 bar: rtol 1e-4 of the quantity's scale, in the default (FP32-accurate tensor-core) math mode and with
the SIMT kernels alone."""
import numpy as np
import pytest

import helpers as H

LAYERS = [(512, 3, 32, 64), (512, 64, 32, 64), (512, 64, 16, 128), (512, 128, 16, 128), (512, 128, 8, 256),
          (512, 256, 8, 256)]   # N, C, H = W, K — 3x3, stride 1, pad 1


def _conv_calls(lib, B, mode):
    from nanograd_b200.nn.buffer import DeviceBuffer
    flags = {"default": B.F_TF32X3, "simt": 0}[mode]

    def run(which, N, C, HW, K, a, b):
        dims = (N, C, HW, HW, K, 3, 3, 1, 1, 1, HW, HW)
        shape = {"fprop": (N, K, HW, HW), "dgrad": (N, C, HW, HW), "wgrad": (K, C, 3, 3)}[which]
        out = DeviceBuffer(shape)
        if which == "fprop":
            lib.ng_conv2d_fprop(out.ptr, a.ptr, b.ptr, 0, *dims, flags)
        elif which == "dgrad":
            lib.ng_conv2d_dgrad(out.ptr, a.ptr, b.ptr, *dims, flags)
        else:
            lib.ng_conv2d_wgrad(out.ptr, a.ptr, b.ptr, *dims, flags)
        return out
    return run


def _dot(a, b):
    return float(np.dot(a.reshape(-1).astype(np.float64), b.reshape(-1).astype(np.float64)))


def _adjoint_tolerance(lhs):
    """The identity is checked relative to ITS OWN magnitude.  The probe dy is correlated with y
    (dy = y + noise), so <y, dy> ~ |y|^2 is of the order of the norm product and a kernel that drops one
    of the nine filter taps changes an inner product by ~10 % — 1000x the bar.  (With independent random
    x, w, dy the inner product is ~|y||dy|/sqrt(n) and a norm-product tolerance accepts anything.)"""
    return H.RTOL_FP32 * abs(lhs)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["default", "simt"])
@pytest.mark.parametrize("layer", LAYERS, ids=lambda s: "x".join(map(str, s)))
def test_conv_passes_are_mutually_adjoint_at_full_size(layer, mode):
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    H.use_cuda_library()
    lib = B.lib()
    run = _conv_calls(lib, B, mode)
    N, C, HW, K = layer
    rng = np.random.RandomState(C + K)
    x = rng.randn(N, C, HW, HW).astype(np.float32)
    w = (rng.randn(K, C, 3, 3) / np.sqrt(9 * C)).astype(np.float32)
    xd, wd = DeviceBuffer(x.shape, hostbuf=x), DeviceBuffer(w.shape, hostbuf=w)
    y = run("fprop", N, C, HW, K, xd, wd).numpy()
    dy = (y + 0.5 * rng.randn(*y.shape)).astype(np.float32)      # correlated probe: <y, dy> ~ |y|^2
    dyd = DeviceBuffer(dy.shape, hostbuf=dy)
    dx = run("dgrad", N, C, HW, K, dyd, wd).numpy()
    dw = run("wgrad", N, C, HW, K, dyd, xd).numpy()
    lhs, mid, rhs = _dot(y, dy), _dot(x, dx), _dot(w, dw)
    assert lhs > 0.3 * float(np.linalg.norm(y.astype(np.float64)) * np.linalg.norm(dy.astype(np.float64)))
    assert abs(lhs - mid) <= _adjoint_tolerance(lhs), (lhs, mid)
    assert abs(lhs - rhs) <= _adjoint_tolerance(lhs), (lhs, rhs)
    # the test has teeth: a dgrad / wgrad that loses ONE filter tap must fail it
    w_cut = w.copy()
    w_cut[:, :, 0, 2] = 0.0
    dx_cut = run("dgrad", N, C, HW, K, dyd, DeviceBuffer(w.shape, hostbuf=w_cut)).numpy()
    assert abs(lhs - _dot(x, dx_cut)) > 100 * _adjoint_tolerance(lhs), "the identity cannot see a dropped tap"
    dw_cut = dw.copy()
    dw_cut[:, :, 2, 0] = 0.0
    assert abs(lhs - _dot(w, dw_cut)) > 100 * _adjoint_tolerance(lhs), "the identity cannot see a dropped tap"


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["default", "simt"])
def test_conv_fprop_is_linear_at_full_size(mode):
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    H.use_cuda_library()
    lib = B.lib()
    run = _conv_calls(lib, B, mode)
    N, C, HW, K = 512, 64, 32, 64
    rng = np.random.RandomState(1)
    x1 = rng.randn(N, C, HW, HW).astype(np.float32)
    x2 = rng.randn(N, C, HW, HW).astype(np.float32)
    w = (rng.randn(K, C, 3, 3) / np.sqrt(9 * C)).astype(np.float32)
    wd = DeviceBuffer(w.shape, hostbuf=w)
    f = lambda a: run("fprop", N, C, HW, K, DeviceBuffer(a.shape, hostbuf=a), wd).numpy()
    combo = f((0.5 * x1 - 2.0 * x2).astype(np.float32))
    H.assert_close(combo, 0.5 * f(x1).astype(np.float64) - 2.0 * f(x2).astype(np.float64), H.RTOL_FP32, "linearity")


@pytest.mark.gpu
def test_linear_layer_is_adjoint_at_full_size():
    """config 2's Linear(4096, 256) on a 512-row batch: <x W^T, dy> = <x, dy W> = <W, dy^T x> through ng_gemm
    in the transpose combinations nn.Linear's forward and backward use."""
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    H.use_cuda_library()
    lib = B.lib()
    rng = np.random.RandomState(2)
    x = rng.randn(512, 4096).astype(np.float32)
    w = (rng.randn(256, 4096) / 64).astype(np.float32)
    dy = ((x.astype(np.float64) @ w.astype(np.float64).T) + 0.5 * rng.randn(512, 256)).astype(np.float32)  # ~ y
    xd, wd, dyd = DeviceBuffer(x.shape, hostbuf=x), DeviceBuffer(w.shape, hostbuf=w), DeviceBuffer(dy.shape, hostbuf=dy)
    for flags in (B.F_TF32X3, 0):
        y, dx, dw = DeviceBuffer((512, 256)), DeviceBuffer((512, 4096)), DeviceBuffer((256, 4096))
        lib.ng_gemm(y.ptr, xd.ptr, wd.ptr, 512, 256, 4096, 0, 1, 1, 0, flags)     # x @ W.T
        lib.ng_gemm(dx.ptr, dyd.ptr, wd.ptr, 512, 4096, 256, 0, 0, 1, 0, flags)   # dy @ W
        lib.ng_gemm(dw.ptr, dyd.ptr, xd.ptr, 256, 4096, 512, 1, 0, 1, 0, flags)   # dy.T @ x
        lhs, mid, rhs = _dot(y.numpy(), dy), _dot(x, dx.numpy()), _dot(w, dw.numpy())
        assert lhs > 0.3 * float(np.linalg.norm(y.numpy().astype(np.float64)) * np.linalg.norm(dy.astype(np.float64)))
        assert abs(lhs - mid) <= _adjoint_tolerance(lhs) and abs(lhs - rhs) <= _adjoint_tolerance(lhs), (lhs, mid, rhs)


@pytest.mark.gpu
def test_pooling_properties_at_full_size():
    from nanograd_b200.device import Device
    from nanograd_b200.tensor import Tensor
    H.use_cuda_library()
    rng = np.random.RandomState(3)
    x_np = np.maximum(rng.randn(512, 64, 32, 32), 0).astype(np.float32)   # post-ReLU: ties at 0 are common
    up = rng.randn(512, 64, 16, 16).astype(np.float32)
    res = {}
    for kind in ("max", "avg"):
        x = Tensor(x_np.copy(), requires_grad=True, device=Device.GPU)
        y = x.max_pool2d((2, 2)) if kind == "max" else x.avg_pool2d((2, 2))
        (y * Tensor(up.copy(), device=Device.GPU)).sum().backward()
        res[kind] = (y.numpy(), x.grad.numpy())
    assert np.all(res["max"][0] >= res["avg"][0])
    windows = x_np.reshape(512, 64, 16, 2, 16, 2)
    H.assert_exact(res["max"][0], windows.max(axis=(3, 5)), "max-pool forward")
    for kind in ("max", "avg"):
        total = float(res[kind][1].astype(np.float64).sum())
        assert abs(total - float(up.astype(np.float64).sum())) <= 1e-4 * float(np.abs(up).astype(np.float64).sum()), kind
    H.assert_exact(res["avg"][1], np.repeat(np.repeat(up * np.float32(0.25), 2, axis=2), 2, axis=3), "avg-pool backward")


@pytest.mark.gpu
def test_batchnorm_training_statistics_at_full_size():
    import nanograd_b200.nn.module as nn
    from nanograd_b200.device import Device
    from nanograd_b200.tensor import Tensor
    H.use_cuda_library()
    rng = np.random.RandomState(4)
    x_np = (rng.randn(512, 64, 32, 32) * rng.uniform(0.5, 3, (1, 64, 1, 1)) + rng.uniform(-2, 2, (1, 64, 1, 1))).astype(np.float32)
    bn = nn.BatchNorm2d(64)
    gamma = rng.uniform(0.5, 2, 64).astype(np.float32)
    beta = rng.uniform(-1, 1, 64).astype(np.float32)
    bn.weight = Tensor(gamma, requires_grad=True, is_parameter=True)
    bn.bias = Tensor(beta, requires_grad=True, is_parameter=True)
    bn.train()
    bn.gpu()
    y = bn(Tensor(x_np, device=Device.GPU)).numpy().astype(np.float64)
    var = x_np.astype(np.float64).var(axis=(0, 2, 3))
    H.assert_close(y.mean(axis=(0, 2, 3)), beta, H.RTOL_FP32, "BN output mean")
    H.assert_close(y.var(axis=(0, 2, 3)), gamma.astype(np.float64) ** 2 * var / (var + 1e-5), H.RTOL_FP32, "BN output variance")
    n = x_np.size // 64
    H.assert_close(bn.running_var.numpy(), 0.9 + 0.1 * var * n / (n - 1), H.RTOL_FP32, "running variance (unbiased)")


@pytest.mark.gpu
def test_optimizer_steps_at_full_size():
    import nanograd_b200.optim.optimizer as optim
    from nanograd_b200.device import Device
    from nanograd_b200.tensor import Tensor
    H.use_cuda_library()
    rng = np.random.RandomState(5)
    shapes = [(256, 4096), (256, 256, 3, 3), (128, 128, 3, 3), (64, 3, 3, 3), (256,), (10, 256), (10,)]
    p_np = [rng.randn(*s).astype(np.float32) for s in shapes]
    g_np = [rng.randn(*s).astype(np.float32) for s in shapes]
    g_np[1][:, :128] = 0.0   # a block of exactly-zero gradients

    def fresh():
        ps = [Tensor(p.copy(), requires_grad=True, is_parameter=True, device=Device.GPU) for p in p_np]
        for p, g in zip(ps, g_np):
            p.grad = Tensor(g.copy(), device=Device.GPU)
        return ps

    ps = fresh()
    optim.SGD(ps, lr=0.01, momentum=0).step()
    for p, p0, g in zip(ps, p_np, g_np):
        H.assert_exact(p.numpy(), p0 - np.float32(0.01) * g, "SGD(momentum=0) step")
    ps = fresh()
    optim.Adam(ps, lr=1e-3).step()
    for p, p0, g in zip(ps, p_np, g_np):
        want = p0.astype(np.float64) - 1e-3 * g.astype(np.float64) / (np.abs(g.astype(np.float64)) + 1e-8)
        np.testing.assert_allclose(p.numpy(), want, rtol=0, atol=1e-4 * 1e-3 + 1e-7 * np.abs(p0).max())
