"""Pins the NumPy oracle (oracle/) to the real reference.

(a) every tests/cases.py case, run on this repo's Tensor with the oracle registered as the
    Device.CPU namespace, must reproduce the golden vectors that tests/golden/make_golden.py
    produced from the REAL reference (/root/reference) — on any machine;
(b) when /root/reference is present (the build container) each oracle/np_ops.py function is
    also compared bit-for-bit with the reference's own ops_cpu Function on random inputs.
"""
import numpy as np
import pytest

import cases
import helpers as H
from nanograd_b200.device import Device

CASES = cases.all_cases()

# float64 NumPy on another CPU / BLAS build differs only by summation order
RTOL_PIN = 1e-6


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_golden(name):
    fn, kwargs = CASES[name]
    fw = H.device_framework(Device.CPU)
    with np.errstate(all="ignore"):
        got = fn(fw, **kwargs)
    want = H.golden(name)
    H.compare_case(name, got, {k: want[k] for k in want.files}, RTOL_PIN, cases.EXACT_KEYS.get(name, ()))


class _Ctx:
    """Minimal stand-in for a Function context when calling the reference's static methods."""

    def __init__(self):
        self.saved_tensors = []

    def save_for_backward(self, *args):
        self.saved_tensors.extend(args)


needs_reference = pytest.mark.skipif(H.load_reference() is None, reason="/root/reference not present")


@needs_reference
@pytest.mark.parametrize("k,s", [(2, 1), (3, 1), (3, 2), (5, 3)])
def test_conv2d_functions_match_reference_bitwise(k, s):
    from oracle import np_ops as K
    ref = H.load_reference()["ops_cpu"]
    rng = np.random.RandomState(k * 7 + s)
    x = rng.normal(0, 1, (3, 4, 11, 11)).astype(np.float32)
    w = rng.normal(0, 1, (5, 4, k, k)).astype(np.float32)
    ctx = _Ctx()
    y_ref = ref.Conv2d.forward(ctx, x, w, s)
    y, cols = K.conv2d_forward(x, w, s)
    H.assert_exact(y, y_ref, "conv2d forward")
    g = rng.normal(0, 1, y.shape).astype(np.float32)
    dx_ref, dw_ref = ref.Conv2d.backward(ctx, g)
    dx, dw = K.conv2d_backward(g, x, w, s, cols)
    H.assert_exact(dx, dx_ref, "conv2d dgrad")
    H.assert_exact(dw, dw_ref, "conv2d wgrad")


@needs_reference
@pytest.mark.parametrize("k,s", [(2, 1), (3, 2), (5, 4)])
def test_conv1d_functions_match_reference_bitwise(k, s):
    from oracle import np_ops as K
    ref = H.load_reference()["ops_cpu"]
    rng = np.random.RandomState(k * 11 + s)
    x = rng.normal(0, 1, (3, 4, 37)).astype(np.float32)
    w = rng.normal(0, 1, (5, 4, k)).astype(np.float32)
    ctx = _Ctx()
    y_ref = ref.Conv1d.forward(ctx, x, w, s)
    y, cols = K.conv1d_forward(x, w, s)
    H.assert_exact(y, y_ref, "conv1d forward")
    g = rng.normal(0, 1, y.shape).astype(np.float32)
    dx_ref, dw_ref = ref.Conv1d.backward(ctx, g)
    dx, dw = K.conv1d_backward(g, x, w, s, cols)
    H.assert_exact(dx, dx_ref, "conv1d dgrad")
    H.assert_exact(dw, dw_ref, "conv1d wgrad")


@needs_reference
def test_elementwise_and_reduction_functions_match_reference_bitwise():
    from oracle import np_ops as K
    ref = H.load_reference()["ops_cpu"]
    rng = np.random.RandomState(3)
    x = np.maximum(np.round(rng.normal(0, 1, (4, 5, 6)) * 2) / 2, 0).astype(np.float32)
    g = rng.normal(0, 1, (4, 6)).astype(np.float32)
    ctx = _Ctx()
    out_ref = ref.Max.forward(ctx, x, 1)
    out, keep, axis = K.extreme_forward(x, 1, "max")
    H.assert_exact(out, out_ref, "max forward")
    H.assert_exact(K.extreme_backward(g, x, axis, keep), ref.Max.backward(ctx, g), "max backward (tie split)")
    H.assert_exact(K.relu_backward(g, x[:, 0, :]), ref.ReLU.backward(_ctx_with(x[:, 0, :]), g), "relu backward")
    a_shape, b_shape = (4, 5, 6), (4, 1, 6)
    gg = rng.normal(0, 1, a_shape).astype(np.float32)
    c2 = _Ctx()
    c2.saved_tensors = [a_shape, b_shape]
    for ours, theirs in zip(K.add_backward(gg, a_shape, b_shape), ref.Add.backward(c2, gg)):
        H.assert_exact(ours, theirs, "add backward / unbroadcast")
        assert ours.dtype == theirs.dtype == np.float64  # the reference's float64 creep, ops_cpu.py:174
    idx = [(1, 3), (-2, 7), (0, 4)]
    H.assert_exact(K.pad_slice(x, idx), ref.inner_slice(x, idx), "inner_slice")
    labels = rng.randint(0, 10, 12).astype(np.float32)
    H.assert_exact(K.onehot_forward(labels, 10), ref.OneHot.forward(_Ctx(), labels, 10), "onehot")


def _ctx_with(*saved):
    c = _Ctx()
    c.saved_tensors = list(saved)
    return c


@needs_reference
def test_reference_relu_subgradient_and_nll_scaling_quirks():
    """Behaviours the oracle must keep: relu'(0) = 1, NLL divides by batch*classes."""
    ref = H.load_reference()
    T = ref["tensor"].Tensor
    x = T(np.zeros((1, 4), dtype=np.float32), requires_grad=True)
    x.relu().sum().backward()
    assert np.all(x.grad.data == 1.0)
    logp = T(np.log(np.full((2, 5), 0.2, dtype=np.float32)), requires_grad=True)
    loss = ref["module"].NLLLoss()(logp, T(np.array([1, 3], dtype=np.float32)))
    np.testing.assert_allclose(loss.data, -np.log(0.2) / 5, rtol=1e-6)
