"""Direct convolution kernels for narrow layers (both channel counts <= 32, stride 1: BASELINE configs[1] at
C = K = 16 / 32): conv_direct_kernel (fprop, dgrad), conv_wgrad_slide_kernel and conv_wgrad_direct_kernel against a float64 evaluation of
ops_cpu.py:345-400's definition, over ragged output widths (strips of 4 with a tail), 1x1 / 3x3 / 5x5 filters,
padding, output-channel tiles of 8 / 16 / 32.  The same kernels run under the host emulation in the CPU suite
(fixture-sized cases of tests/test_device_parity.py)."""
import numpy as np
import pytest

import helpers as H

#        N   C   H   W   K  R  S  pad
SHAPES = [(256, 16, 32, 32, 16, 3, 3, 0), (64, 32, 32, 32, 32, 3, 3, 1), (70, 8, 30, 34, 24, 3, 3, 1),
          (40, 16, 45, 47, 12, 5, 5, 2), (128, 24, 24, 24, 8, 1, 1, 0), (66, 12, 33, 31, 30, 3, 3, 0)]


def _truth(x, w, dy, pad):
    n, c, h, wd = x.shape
    k, _, r, s = w.shape
    oh, ow = dy.shape[2:]
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    dyp = np.pad(dy.astype(np.float64), ((0, 0), (0, 0), (r - 1 - pad, r - 1 - pad), (s - 1 - pad, s - 1 - pad)))
    w64 = w.astype(np.float64)
    y, dx, dw = np.zeros(dy.shape), np.zeros(x.shape), np.zeros(w.shape)
    for i in range(r):
        for j in range(s):
            y += np.einsum("nchw,kc->nkhw", xp[:, :, i:i + oh, j:j + ow], w64[:, :, i, j])
            dx += np.einsum("nkhw,kc->nchw", dyp[:, :, r - 1 - i:r - 1 - i + h, s - 1 - j:s - 1 - j + wd], w64[:, :, i, j])
            dw[:, :, i, j] = np.einsum("nkhw,nchw->kc", dy.astype(np.float64), xp[:, :, i:i + oh, j:j + ow])
    return y, dx, dw


# fixture-sized shapes for the host emulation (>= 1024 output pixels): even / odd widths (8- and 4-byte cp.async tiles
# of the sliding-window wgrad), padding 0 / 1, K % 4 != 0 (the register-tile wgrad), ragged last row block
EMU_SHAPES = [(5, 16, 18, 18, 16, 3, 3, 0), (6, 8, 16, 15, 8, 3, 3, 1), (5, 12, 17, 18, 20, 3, 3, 1),
              (6, 16, 16, 16, 10, 3, 3, 0), (3, 32, 22, 20, 16, 3, 3, 0), (3, 32, 20, 20, 32, 3, 3, 1)]


def _check(shape):
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    lib = B.lib()
    N, C, Hh, W, K, R, S, pad = shape
    OH, OW = Hh + 2 * pad - R + 1, W + 2 * pad - S + 1
    rng = np.random.RandomState(sum(shape))
    x = rng.randn(N, C, Hh, W).astype(np.float32)
    w = (rng.randn(K, C, R, S) / np.sqrt(R * S * C)).astype(np.float32)
    b = rng.randn(K).astype(np.float32)
    dy = rng.randn(N, K, OH, OW).astype(np.float32)
    xd, wd, bd, dyd = (DeviceBuffer(a.shape, hostbuf=a) for a in (x, w, b, dy))
    dims = (N, C, Hh, W, K, R, S, 1, pad, pad, OH, OW)
    y, dx, dw = DeviceBuffer((N, K, OH, OW)), DeviceBuffer((N, C, Hh, W)), DeviceBuffer((K, C, R, S))
    lib.ng_conv2d_fprop(y.ptr, xd.ptr, wd.ptr, bd.ptr, *dims, B.F_RELU)
    lib.ng_conv2d_dgrad(dx.ptr, dyd.ptr, wd.ptr, *dims, 0)
    lib.ng_conv2d_wgrad(dw.ptr, dyd.ptr, xd.ptr, *dims, 0)
    ty, tdx, tdw = _truth(x, w, dy, pad)
    ty = np.maximum(ty + b.reshape(1, -1, 1, 1), 0)
    H.assert_close(y.numpy(), ty, H.RTOL_FP32, f"fprop {shape}")
    H.assert_close(dx.numpy(), tdx, H.RTOL_FP32, f"dgrad {shape}")
    H.assert_close(dw.numpy(), tdw, H.RTOL_FP32, f"wgrad {shape}")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES + EMU_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_narrow_layer_kernels_match_float64(shape):
    H.use_cuda_library()
    _check(shape)


@pytest.mark.parametrize("shape", EMU_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_narrow_layer_kernels_match_float64_emulated(shape):
    H.use_emulated_library()
    _check(shape)
