"""Elementwise parity at BASELINE.json's full sizes:

  * config 2 — the six 3x3 / pad 1 conv layers of the VGG at batch 512, and its Linear(4096, 256);
  * config 1 — the five sweep shapes x(256, C, 32, 32) * w(C, C, 3, 3), C in {16, 32, 64, 128, 256}, no pad.

For fprop, dgrad and wgrad of every shape:
  (1) the tensor-core engines (tcgen05 3xTF32 = the default FP32-accurate mode, and opt-in TF32) are
      compared ELEMENTWISE, over the whole output tensor, with the FP32 SIMT kernels
      (rtol 1e-4 / 1e-2 of max|ref|, the north-star bars);
  (2) a few thousand sampled output elements of every engine are compared with a float64 NumPy direct
      evaluation of the reference's definition (ops_cpu.py:345-400: valid cross-correlation of the padded
      input; dX = sum over taps of dY @ W scattered; dW = tensordot(dY, windows)) — the full float64 oracle
      needs minutes per pass at these sizes, the sample needs a second.
Shapes the tensor-core kernels do not take (the 3-channel layer, C = 16 wgrad ...) fall back to SIMT
inside the library; the comparison still runs, and `ng_conv2d_tc_plan` is reported in the assertion text.
"""
import ctypes

import numpy as np
import pytest

import helpers as H

CFG2_LAYERS = [("cfg2", 512, c, hw, k, 1) for c, hw, k in
               [(3, 32, 64), (64, 32, 64), (64, 16, 128), (128, 16, 128), (128, 8, 256), (256, 8, 256)]]
CFG1_SWEEP = [("cfg1", 256, c, 32, c, 0) for c in (16, 32, 64, 128, 256)]
N_SAMPLES = {"fprop": 4096, "dgrad": 4096, "wgrad": 384}


def _ids(s):
    return f"{s[0]}-N{s[1]}-C{s[2]}-HW{s[3]}-K{s[4]}-pad{s[5]}"


def _geometry(shape):
    _, n, c, hw, k, pad = shape
    oh = hw + 2 * pad - 2
    return n, c, hw, k, pad, oh


def _run_pass(lib, which, shape, a, b, flags):
    from nanograd_b200.nn.buffer import DeviceBuffer
    n, c, hw, k, pad, oh = _geometry(shape)
    dims = (n, c, hw, hw, k, 3, 3, 1, pad, pad, oh, oh)
    out = DeviceBuffer({"fprop": (n, k, oh, oh), "dgrad": (n, c, hw, hw), "wgrad": (k, c, 3, 3)}[which])
    if which == "fprop":
        lib.ng_conv2d_fprop(out.ptr, a.ptr, b.ptr, 0, *dims, flags)
    elif which == "dgrad":
        lib.ng_conv2d_dgrad(out.ptr, a.ptr, b.ptr, *dims, flags)
    else:
        lib.ng_conv2d_wgrad(out.ptr, a.ptr, b.ptr, *dims, flags)
    return out.numpy()


# ------------------------------------------------------------------ float64 direct evaluation (samples)
def _sample_fprop(x, w, pad, oh, rng, count):
    n, c, hw, _ = x.shape
    k = w.shape[0]
    idx = np.stack([rng.integers(0, n, count), rng.integers(0, k, count), rng.integers(0, oh, count),
                    rng.integers(0, oh, count)], axis=1)
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    ref = np.empty(count)
    for i, (ni, ki, yi, xi) in enumerate(idx):
        ref[i] = np.sum(xp[ni, :, yi:yi + 3, xi:xi + 3].astype(np.float64) * w[ki].astype(np.float64))
    return idx, ref


def _sample_dgrad(dy, w, pad, hw, rng, count):
    n, k, oh, _ = dy.shape
    c = w.shape[1]
    idx = np.stack([rng.integers(0, n, count), rng.integers(0, c, count), rng.integers(0, hw, count),
                    rng.integers(0, hw, count)], axis=1)
    # dx[n,c,h,w] = sum_{k,r,s} dy[n,k,h+pad-r,w+pad-s] * W[k,c,r,s]  (ops_cpu.py:390-396 scattered the other way)
    dyp = np.pad(dy, ((0, 0), (0, 0), (2, 2), (2, 2)))
    wf = w[:, :, ::-1, ::-1].astype(np.float64)
    ref = np.empty(count)
    for i, (ni, ci, hi, wi) in enumerate(idx):
        y0, x0 = hi + pad - 2 + 2, wi + pad - 2 + 2          # window start in the padded dy
        ref[i] = np.sum(dyp[ni, :, y0:y0 + 3, x0:x0 + 3].astype(np.float64) * wf[:, ci])
    return idx, ref


def _sample_wgrad(dy, x, pad, rng, count):
    n, k, oh, _ = dy.shape
    c = x.shape[1]
    idx = np.stack([rng.integers(0, k, count), rng.integers(0, c, count), rng.integers(0, 3, count),
                    rng.integers(0, 3, count)], axis=1)
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    ref = np.empty(count)
    for i, (ki, ci, r, s) in enumerate(idx):
        ref[i] = np.sum(dy[:, ki].astype(np.float64) * xp[:, ci, r:r + oh, s:s + oh].astype(np.float64))
    return idx, ref


def _pick(arr, idx):
    return arr[tuple(idx[:, j] for j in range(idx.shape[1]))].astype(np.float64)


def _rel(got, ref, scale):
    return float(np.max(np.abs(got - ref)) / scale)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", CFG2_LAYERS + CFG1_SWEEP, ids=_ids)
def test_conv_engines_elementwise_and_against_float64_at_full_size(shape):
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    H.use_cuda_library()
    lib = B.lib()
    n, c, hw, k, pad, oh = _geometry(shape)
    rng = np.random.default_rng(1000 + c + k + hw)
    x = rng.standard_normal((n, c, hw, hw), dtype=np.float32)
    w = (rng.standard_normal((k, c, 3, 3), dtype=np.float32) / np.float32(np.sqrt(9 * c)))
    dy = rng.standard_normal((n, k, oh, oh), dtype=np.float32)
    xd, wd, dyd = DeviceBuffer(x.shape, hostbuf=x), DeviceBuffer(w.shape, hostbuf=w), DeviceBuffer(dy.shape, hostbuf=dy)
    plan = (ctypes.c_int * 4)()
    lib.ng_conv2d_tc_plan(n, c, hw, hw, k, 3, 3, 1, oh, oh, B.F_TF32X3 | B.F_TC_FORCE, plan)
    plan = dict(zip(("fprop", "dgrad", "wgrad"), map(bool, plan)))

    samples = {"fprop": _sample_fprop(x, w, pad, oh, rng, N_SAMPLES["fprop"]),
               "dgrad": _sample_dgrad(dy, w, pad, hw, rng, N_SAMPLES["dgrad"]),
               "wgrad": _sample_wgrad(dy, x, pad, rng, N_SAMPLES["wgrad"])}
    operands = {"fprop": (xd, wd), "dgrad": (dyd, wd), "wgrad": (dyd, xd)}
    engines = [("simt", 0, H.RTOL_FP32), ("tf32x3", B.F_TF32X3 | B.F_TC_FORCE, H.RTOL_FP32),
               ("tf32", B.F_TF32, H.RTOL_TF32)]
    if c % 64 == 0 and k % 64 == 0:
        engines.insert(2, ("f16x3", B.F_F16X3 | B.F_TC_FORCE, H.RTOL_FP32))
    report = []
    for which in ("fprop", "dgrad", "wgrad"):
        idx, ref = samples[which]
        simt = None
        for name, flags, rtol in engines:
            got = _run_pass(lib, which, shape, *operands[which], flags)
            scale = float(np.max(np.abs(got)))
            assert np.isfinite(scale) and scale > 0, (which, name)
            err64 = _rel(_pick(got, idx), ref, max(float(np.max(np.abs(ref))), 1e-30))
            what = f"{_ids(shape)} {which} [{name}, tensor-core kernel taken: {plan[which]}]"
            assert err64 <= rtol, f"{what}: {err64:.3e} from float64 on {len(ref)} sampled elements (bar {rtol})"
            if simt is None:
                simt = got
                report.append(f"{which}/{name}: f64 {err64:.1e}")
                continue
            err = _rel(got.astype(np.float64), simt.astype(np.float64), float(np.max(np.abs(simt))))
            assert err <= rtol, f"{what}: {err:.3e} elementwise from the SIMT kernel over {got.size} elements (bar {rtol})"
            report.append(f"{which}/{name}: f64 {err64:.1e} simt {err:.1e}")
    print(_ids(shape), "; ".join(report))


@pytest.mark.gpu
def test_float64_sampler_detects_a_dropped_tap():
    """The float64 sampling check has teeth: fprop with one of the nine taps zeroed is rejected."""
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    H.use_cuda_library()
    shape = ("cfg2", 512, 64, 16, 128, 1)
    n, c, hw, k, pad, oh = _geometry(shape)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((n, c, hw, hw), dtype=np.float32)
    w = (rng.standard_normal((k, c, 3, 3), dtype=np.float32) / np.float32(np.sqrt(9 * c)))
    idx, ref = _sample_fprop(x, w, pad, oh, rng, 2048)
    w_cut = w.copy()
    w_cut[:, :, 1, 2] = 0
    got = _run_pass(B.lib(), "fprop", shape, DeviceBuffer(x.shape, hostbuf=x), DeviceBuffer(w.shape, hostbuf=w_cut),
                    B.F_TF32X3 | B.F_TC_FORCE)
    assert _rel(_pick(got, idx), ref, float(np.max(np.abs(ref)))) > 100 * H.RTOL_FP32


@pytest.mark.gpu
def test_linear_4096x256_engines_elementwise_and_against_float64():
    """config 2's Linear(4096, 256) at batch 512: y = x W^T, dx = dy W, dW = dy^T x in the three engines."""
    from nanograd_b200 import backend as B
    from nanograd_b200.nn.buffer import DeviceBuffer
    H.use_cuda_library()
    lib = B.lib()
    rng = np.random.default_rng(6)
    x = rng.standard_normal((512, 4096), dtype=np.float32)
    w = rng.standard_normal((256, 4096), dtype=np.float32) / np.float32(64)
    dy = rng.standard_normal((512, 256), dtype=np.float32)
    x64, w64, dy64 = x.astype(np.float64), w.astype(np.float64), dy.astype(np.float64)
    truth = {"y": x64 @ w64.T, "dx": dy64 @ w64, "dw": dy64.T @ x64}
    xd, wd, dyd = DeviceBuffer(x.shape, hostbuf=x), DeviceBuffer(w.shape, hostbuf=w), DeviceBuffer(dy.shape, hostbuf=dy)
    for name, flags, rtol in [("simt", 0, H.RTOL_FP32), ("tf32x3", B.F_TF32X3 | B.F_TC_FORCE, H.RTOL_FP32),
                              ("tf32", B.F_TF32, H.RTOL_TF32)]:
        y, dx, dw = DeviceBuffer((512, 256)), DeviceBuffer((512, 4096)), DeviceBuffer((256, 4096))
        lib.ng_gemm(y.ptr, xd.ptr, wd.ptr, 512, 256, 4096, 0, 1, 1, 0, flags)
        lib.ng_gemm(dx.ptr, dyd.ptr, wd.ptr, 512, 4096, 256, 0, 0, 1, 0, flags)
        lib.ng_gemm(dw.ptr, dyd.ptr, xd.ptr, 256, 4096, 512, 1, 0, 1, 0, flags)
        for key, buf in (("y", y), ("dx", dx), ("dw", dw)):
            H.assert_close(buf.numpy(), truth[key], rtol, f"Linear(4096,256) {key} [{name}] vs float64")
