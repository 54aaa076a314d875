"""tcgen05 convolution kernels vs the FP32 SIMT kernels (which the golden-vector tests pin to the
reference), through the C ABI, over tile-edge cases: ragged widths/heights, channel counts that are
not multiples of the 32-channel stage, more than 256 output channels (two N tiles), several images
per 128-pixel tile, 1x1 and 5x5 filters, Conv1d-shaped problems (H = R = 1).

FP32-accurate 3xTF32 mode: rtol 1e-4 (scale-relative); plain TF32 mode: rtol 1e-2."""
import numpy as np
import pytest

import helpers as H

#        N   C   H   W    K  R  S  pad_h pad_w
SHAPES = [
    (2, 32, 32, 32, 32, 3, 3, 1, 1), (4, 64, 16, 16, 128, 3, 3, 1, 1), (4, 128, 8, 8, 256, 3, 3, 1, 1),
    (3, 16, 32, 32, 48, 3, 3, 0, 0), (2, 8, 28, 28, 16, 3, 3, 0, 0), (2, 32, 32, 32, 32, 1, 1, 0, 0),
    (5, 40, 12, 12, 24, 5, 5, 2, 2), (3, 64, 9, 13, 320, 3, 3, 1, 1), (7, 96, 5, 5, 40, 3, 3, 1, 1),
    (2, 32, 4, 4, 32, 3, 3, 1, 1), (130, 32, 2, 2, 64, 1, 1, 0, 0), (2, 256, 20, 36, 32, 3, 3, 0, 0),
    (1, 32, 33, 65, 32, 3, 3, 2, 2), (5, 32, 1, 100, 48, 1, 3, 0, 0), (9, 64, 1, 37, 64, 1, 5, 0, 2),
]


# channel counts that are multiples of 64: the FP16x3 engine (ragged sizes, two N tiles, 1x1 / 5x5 / Conv1d shapes)
F16_SHAPES = [
    (2, 64, 32, 32, 64, 3, 3, 1, 1), (4, 64, 16, 16, 128, 3, 3, 1, 1), (4, 128, 8, 8, 256, 3, 3, 1, 1),
    (3, 64, 9, 13, 320, 3, 3, 1, 1), (2, 256, 20, 36, 64, 3, 3, 0, 0), (130, 64, 2, 2, 64, 1, 1, 0, 0),
    (1, 64, 33, 65, 192, 3, 3, 2, 2), (5, 128, 12, 12, 64, 5, 5, 2, 2), (2, 192, 7, 7, 128, 3, 3, 1, 1),
    (9, 64, 1, 37, 64, 1, 5, 0, 2), (3, 320, 6, 10, 64, 3, 3, 1, 1),
]


@pytest.mark.gpu
@pytest.mark.parametrize("mode,rtol", [("tf32x3", H.RTOL_FP32), ("tf32", H.RTOL_TF32), ("f16x3", H.RTOL_FP32)])
@pytest.mark.parametrize("shape", SHAPES + F16_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_tensor_core_conv_matches_simt(shape, mode, rtol):
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    H.use_cuda_library()
    lib = B.lib()
    N, C, Hh, W, K, R, S, ph, pw = shape
    OH, OW = Hh + 2 * ph - R + 1, W + 2 * pw - S + 1
    rng = np.random.RandomState(sum(shape))
    x = DeviceBuffer((N, C, Hh, W), hostbuf=rng.randn(N, C, Hh, W).astype(np.float32))
    w = DeviceBuffer((K, C, R, S), hostbuf=(rng.randn(K, C, R, S) / np.sqrt(R * S * C)).astype(np.float32))
    b = DeviceBuffer((K,), hostbuf=rng.randn(K).astype(np.float32))
    dy = DeviceBuffer((N, K, OH, OW), hostbuf=rng.randn(N, K, OH, OW).astype(np.float32))
    dims = (N, C, Hh, W, K, R, S, 1, ph, pw, OH, OW)
    flag = {"tf32x3": B.F_TF32X3 | B.F_TC_FORCE, "tf32": B.F_TF32, "f16x3": B.F_F16X3 | B.F_TC_FORCE}[mode]
    if mode == "f16x3" and (C % 64 or K % 64):
        pytest.skip("the FP16x3 engine takes channel counts that are multiples of 64")
    calls = {
        "fprop": ((N, K, OH, OW), lambda o, f: lib.ng_conv2d_fprop(o.ptr, x.ptr, w.ptr, b.ptr, *dims, f)),
        "dgrad": ((N, C, Hh, W), lambda o, f: lib.ng_conv2d_dgrad(o.ptr, dy.ptr, w.ptr, *dims, f)),
        "wgrad": ((K, C, R, S), lambda o, f: lib.ng_conv2d_wgrad(o.ptr, dy.ptr, x.ptr, *dims, f)),
    }
    for name, (oshape, fn) in calls.items():
        ref, got = DeviceBuffer(oshape), DeviceBuffer(oshape)
        fn(ref, 0)
        fn(got, flag)
        H.assert_close(got.numpy(), ref.numpy(), rtol, f"{name} {mode} {shape}")


GEMM_SHAPES = [(512, 256, 4096), (256, 4096, 512), (128, 64, 96), (148, 96, 132), (64, 32, 32), (300, 160, 1000),
               (512, 10, 256)]


@pytest.mark.gpu
@pytest.mark.parametrize("mode,rtol", [("tf32x3", H.RTOL_FP32), ("tf32", H.RTOL_TF32)])
@pytest.mark.parametrize("mnk", GEMM_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_tensor_core_gemm_matches_simt(mnk, mode, rtol):
    """ng_gemm on tcgen05 (K-major and MN-major operands read in place by TMA, split-K, bias + ReLU
    epilogue) vs the FP32 SIMT GEMM for every transpose combination; shapes the tensor-core kernel does
    not take (N = 10, M % 32 != 0 with a transposed A) exercise the per-call fallback."""
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    H.use_cuda_library()
    lib = B.lib()
    M, N, K = mnk
    rng = np.random.RandomState(M + N + K)
    flag = {"tf32x3": B.F_TF32X3 | B.F_TC_FORCE, "tf32": B.F_TF32}[mode]
    for ta in (0, 1):
        for tb in (0, 1):
            a = rng.randn(*((K, M) if ta else (M, K))).astype(np.float32)
            b = rng.randn(*((N, K) if tb else (K, N))).astype(np.float32)
            bias = rng.randn(N).astype(np.float32)
            ad, bd, biasd = DeviceBuffer(a.shape, hostbuf=a), DeviceBuffer(b.shape, hostbuf=b), DeviceBuffer((N,), hostbuf=bias)
            ref, got = DeviceBuffer((M, N)), DeviceBuffer((M, N))
            lib.ng_gemm(ref.ptr, ad.ptr, bd.ptr, M, N, K, ta, tb, 1, biasd.ptr, B.F_RELU)
            lib.ng_gemm(got.ptr, ad.ptr, bd.ptr, M, N, K, ta, tb, 1, biasd.ptr, B.F_RELU | flag)
            H.assert_close(got.numpy(), ref.numpy(), rtol, f"gemm {mode} {mnk} ta={ta} tb={tb}")


@pytest.mark.gpu
@pytest.mark.parametrize("nchw", [(3, 40, 5, 10), (16, 64, 32, 32), (5, 33, 1, 37), (70, 256, 8, 8)],
                         ids=lambda s: "x".join(map(str, s)))
def test_stage_nhwc_with_channel_sums(nchw):
    """ng_conv2d_stage_nhwc_sum: the channels-last staging copy of dy (bit-exact transpose) and the
    bias gradient from the same read (Add.backward's unbroadcast, ops_cpu.py:12-18), ragged tiles."""
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    H.use_cuda_library()
    lib = B.lib()
    n, c, h, w = nchw
    rng = np.random.RandomState(n + c)
    a = rng.randn(n, c, h, w).astype(np.float32)
    src, dst, sums = DeviceBuffer(a.shape, hostbuf=a), DeviceBuffer(a.shape), DeviceBuffer((c,))
    lib.ng_conv2d_stage_nhwc_sum(dst.ptr, src.ptr, sums.ptr, n, c, h * w, 0)
    H.assert_exact(dst.numpy().reshape(n, h * w, c), a.reshape(n, c, h * w).transpose(0, 2, 1), "staged copy")
    H.assert_close(sums.numpy(), a.astype(np.float64).sum(axis=(0, 2, 3)), H.RTOL_FP32, "channel sums")


@pytest.mark.gpu
@pytest.mark.parametrize("spread", [1.0, 3.0, 6.0])
def test_f16x3_conv_dynamic_range(spread):
    """FP16x3 operands carry one power-of-two scale per tensor, so the engine is checked on tensors whose
    magnitudes spread over many binades (log-normal with sigma = `spread` natural-log units: 6.0 puts the
    smallest elements ~1e-11 of the largest) against a float64 evaluation of sampled outputs: the error bar
    stays the FP32 one (rtol 1e-4 of the output scale), and elements far below the tensor maximum lose only
    relative, never absolute, accuracy."""
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    H.use_cuda_library()
    lib = B.lib()
    N, C, Hh, W, K, R, S, ph, pw = 3, 64, 12, 12, 128, 3, 3, 1, 1
    rng = np.random.RandomState(int(spread * 10))
    def wide(*shape):
        return (rng.randn(*shape) * np.exp(spread * rng.randn(*shape))).astype(np.float32)
    x, w, dy = wide(N, C, Hh, W), wide(K, C, R, S), wide(N, K, Hh, W)
    xd, wd, dyd = (DeviceBuffer(a.shape, hostbuf=a) for a in (x, w, dy))
    dims = (N, C, Hh, W, K, R, S, 1, ph, pw, Hh, W)
    flag = B.F_F16X3 | B.F_TC_FORCE
    y, dx, dw = DeviceBuffer((N, K, Hh, W)), DeviceBuffer((N, C, Hh, W)), DeviceBuffer((K, C, R, S))
    lib.ng_conv2d_fprop(y.ptr, xd.ptr, wd.ptr, 0, *dims, flag)
    lib.ng_conv2d_dgrad(dx.ptr, dyd.ptr, wd.ptr, *dims, flag)
    lib.ng_conv2d_wgrad(dw.ptr, dyd.ptr, xd.ptr, *dims, flag)
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (1, 1), (1, 1)))
    dyp = np.pad(dy.astype(np.float64), ((0, 0), (0, 0), (1, 1), (1, 1)))
    w64 = w.astype(np.float64)
    y_ref = np.zeros((N, K, Hh, W))
    dx_ref = np.zeros((N, C, Hh, W))
    dw_ref = np.zeros((K, C, R, S))
    for r in range(R):
        for s in range(S):
            y_ref += np.einsum("nchw,kc->nkhw", xp[:, :, r:r + Hh, s:s + W], w64[:, :, r, s])
            dx_ref += np.einsum("nkhw,kc->nchw", dyp[:, :, 2 - r:2 - r + Hh, 2 - s:2 - s + W], w64[:, :, r, s])
            dw_ref[:, :, r, s] = np.einsum("nkhw,nchw->kc", dy.astype(np.float64), xp[:, :, r:r + Hh, s:s + W])
    H.assert_close(y.numpy(), y_ref, H.RTOL_FP32, f"fprop spread {spread}")
    H.assert_close(dx.numpy(), dx_ref, H.RTOL_FP32, f"dgrad spread {spread}")
    H.assert_close(dw.numpy(), dw_ref, H.RTOL_FP32, f"wgrad spread {spread}")


@pytest.mark.gpu
@pytest.mark.parametrize("nchw", [(3, 64, 5, 10), (16, 128, 32, 32), (5, 192, 1, 37), (70, 256, 8, 8)],
                         ids=lambda s: "x".join(map(str, s)))
def test_stage_f16_pairs_with_channel_sums(nchw):
    """ng_conv2d_stage_nhwc_sum with NG_F_STAGE_F16: staged[n][p][c/64][hi|lo][c%64] in FP16 with a power-of-two
    scale that puts max|src| in [2^14, 2^15); hi + lo reproduces the scaled value to 2^-22 relative
    (11 + 11 significant bits) and the channel sums ride on the same read."""
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    H.use_cuda_library()
    lib = B.lib()
    n, c, h, w = nchw
    rng = np.random.RandomState(n + c)
    a = (rng.randn(n, c, h, w) * np.exp(rng.randn(n, c, h, w))).astype(np.float32)
    src, dst, sums = DeviceBuffer(a.shape, hostbuf=a), DeviceBuffer((a.size + 16,)), DeviceBuffer((c,))
    lib.ng_conv2d_stage_nhwc_sum(dst.ptr, src.ptr, sums.ptr, n, c, h * w, B.F_STAGE_F16)
    raw = dst.numpy()
    inv_scale = float(raw[a.size])
    amax = float(raw[a.size + 1:a.size + 2].view(np.float32)[0])
    assert amax == float(np.abs(a).max())
    scale = 1.0 / inv_scale
    assert 2.0 ** 14 <= amax * scale < 2.0 ** 15 and np.log2(scale) == np.round(np.log2(scale))
    halves = raw[:a.size].view(np.float16).reshape(n, h * w, c // 64, 2, 64).astype(np.float64)
    rebuilt = (halves[:, :, :, 0, :] + halves[:, :, :, 1, :]).reshape(n, h * w, c) * inv_scale
    want = a.reshape(n, c, h * w).transpose(0, 2, 1).astype(np.float64)
    err = np.abs(rebuilt - want)
    assert np.all(err <= np.maximum(np.abs(want) * 2.0 ** -21, amax * 2.0 ** -39)), float((err / np.abs(want).max()).max())
    H.assert_close(sums.numpy(), a.astype(np.float64).sum(axis=(0, 2, 3)), H.RTOL_FP32, "channel sums")
    # without the sums (the x operand): the same bytes
    dst2 = DeviceBuffer((a.size + 16,))
    lib.ng_conv2d_stage_nhwc(dst2.ptr, src.ptr, n, c, h * w, B.F_STAGE_F16)
    H.assert_exact(dst2.numpy()[:a.size + 2].view(np.uint32), raw[:a.size + 2].view(np.uint32), "staged copy without sums")
