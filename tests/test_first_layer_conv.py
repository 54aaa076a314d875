"""The image-facing convolution (few input channels, 3-wide filters, stride 1): the direct kernels
conv_wgrad_smallc_kernel and conv_fprop_smallc_kernel (csrc/conv_gemm.cu) through the C ABI against
the NumPy oracle (oracle/np_ops.py conv2d_forward / conv2d_backward, ops_cpu.py:347-400 in the
reference) on the zero-padded input.  Covers the row-group shapes (C*R = 1, 2, 3, 6, 9, 12 filter rows), a ragged last
row block, asymmetric padding and a Conv1d-shaped problem; FP32 bar: rtol 1e-4 scale-relative."""
import numpy as np
import pytest

import helpers as H
from oracle import np_ops

#        N  C   H   W   K  R  pad_t pad_l
SHAPES = [
    (4, 3, 32, 32, 64, 3, 1, 1),     # the VGG first layer at a small batch
    (6, 1, 28, 30, 64, 3, 0, 0),     # MNIST-like, one filter-row group
    (6, 2, 19, 40, 96, 3, 1, 1),     # ragged last row block, K = 96
    (4, 4, 33, 36, 256, 3, 1, 1),    # 12 filter rows, one output row per tile
    (40, 3, 1, 128, 64, 1, 0, 1),    # Conv1d: H = R = 1
    (4, 2, 34, 36, 64, 1, 0, 1),     # 1x3 filter on a 2-D image (2 filter rows)
]


def _check(shape):
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    lib = B.lib()
    N, C, Hh, W, K, R, pt, pl = shape
    S = 3
    OH, OW = Hh + 2 * pt - R + 1, W + 2 * pl - S + 1
    assert OW % 4 == 0 and N * OH * OW >= 4096, "shape would not reach the direct kernel"
    rng = np.random.RandomState(sum(shape))
    x = rng.randn(N, C, Hh, W).astype(np.float32)
    dy = rng.randn(N, K, OH, OW).astype(np.float32)
    xd, dyd, dwd = DeviceBuffer(x.shape, hostbuf=x), DeviceBuffer(dy.shape, hostbuf=dy), DeviceBuffer((K, C, R, S))
    lib.ng_conv2d_wgrad(dwd.ptr, dyd.ptr, xd.ptr, N, C, Hh, W, K, R, S, 1, pt, pl, OH, OW, 0)
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (pt, pt), (pl, pl)))
    w0 = np.zeros((K, C, R, S))
    _, cols = np_ops.conv2d_forward(xp, w0, 1)
    _, want = np_ops.conv2d_backward(dy.astype(np.float64), xp, w0, 1, cols)
    H.assert_close(dwd.numpy(), want, H.RTOL_FP32, f"wgrad {shape}")


#              N  C   H   W   K  pad  relu
FPROP_SHAPES = [
    (4, 3, 32, 32, 64, 1, 1),
    (6, 1, 28, 30, 8, 0, 0),
    (5, 2, 30, 32, 40, 1, 1),
    (3, 4, 40, 36, 100, 1, 0),
]


def _check_fprop(shape):
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    lib = B.lib()
    N, C, Hh, W, K, pad, relu = shape
    OH, OW = Hh + 2 * pad - 2, W + 2 * pad - 2
    assert OW % 4 == 0 and N * OH * OW >= 4096, "shape would not reach the direct kernel"
    rng = np.random.RandomState(sum(shape))
    x = rng.randn(N, C, Hh, W).astype(np.float32)
    w = (rng.randn(K, C, 3, 3) / 3).astype(np.float32)
    b = rng.randn(K).astype(np.float32)
    xd, wd, bd = DeviceBuffer(x.shape, hostbuf=x), DeviceBuffer(w.shape, hostbuf=w), DeviceBuffer((K,), hostbuf=b)
    yd = DeviceBuffer((N, K, OH, OW))
    lib.ng_conv2d_fprop(yd.ptr, xd.ptr, wd.ptr, bd.ptr, N, C, Hh, W, K, 3, 3, 1, pad, pad, OH, OW,
                        B.F_RELU if relu else 0)
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    want, _ = np_ops.conv2d_forward(xp, w.astype(np.float64), 1)
    want = want + b.astype(np.float64)[None, :, None, None]
    if relu:
        want = np.maximum(want, 0)
    H.assert_close(yd.numpy(), want, H.RTOL_FP32, f"fprop {shape}")


@pytest.mark.parametrize("shape", FPROP_SHAPES[:3], ids=lambda s: "x".join(map(str, s)))
def test_first_layer_fprop_emulated(shape):
    H.use_emulated_library()
    _check_fprop(shape)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", FPROP_SHAPES + [(512, 3, 32, 32, 64, 1, 0)], ids=lambda s: "x".join(map(str, s)))
def test_first_layer_fprop_on_b200(shape):
    H.use_cuda_library()
    _check_fprop(shape)


@pytest.mark.parametrize("shape", SHAPES[:3], ids=lambda s: "x".join(map(str, s)))
def test_first_layer_wgrad_emulated(shape):
    H.use_emulated_library()
    _check(shape)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES + [(512, 3, 32, 32, 64, 3, 1, 1)], ids=lambda s: "x".join(map(str, s)))
def test_first_layer_wgrad_on_b200(shape):
    H.use_cuda_library()
    _check(shape)
