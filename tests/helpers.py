"""Shared test helpers: tolerance definition, reference loader, device selection, model zoo."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOT = "/root/reference"
# the unmodified reference pip-installed by tools/make_oracle_ref.sh (git-ignored, shipped to the GPU box)
REFERENCE_INSTALL = os.path.join(ROOT, "oracle", "_ref")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# FP32 parity bar of BASELINE.json's north_star: rtol 1e-4 (1e-2 in TF32 mode), evaluated
# scale-relative — allclose(rtol=R, atol=R*max|ref|) — because outputs that are sums of
# zero-mean products land arbitrarily close to zero and a pure elementwise rtol is unattainable
# even for the oracle against itself.
RTOL_FP32 = 1e-4
RTOL_TF32 = 1e-2


def assert_close(actual, expected, rtol=RTOL_FP32, what=""):
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    assert actual.shape == expected.shape, f"{what}: shape {actual.shape} != {expected.shape}"
    if expected.size == 0:
        return
    scale = float(np.max(np.abs(expected)))
    np.testing.assert_allclose(actual, expected, rtol=rtol, atol=rtol * scale + 1e-30, err_msg=what)


def assert_exact(actual, expected, what=""):
    actual, expected = np.asarray(actual), np.asarray(expected)
    assert actual.shape == expected.shape, f"{what}: shape {actual.shape} != {expected.shape}"
    assert np.array_equal(actual, expected), f"{what}: not bit-exact ({np.sum(actual != expected)} mismatches)"


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


_REF = None


def load_reference():
    """Import the real reference next to this repo's packages: /root/reference (the build
    container) or its pip-installed copy oracle/_ref (the GPU box; tools/make_oracle_ref.sh).

    Returns a dict of its modules, or None when neither is present.
    The reference imports ``graphviz`` unconditionally; tests/_shim carries a stub.
    """
    global _REF
    if _REF is not None:
        return _REF or None
    root = next((r for r in (REFERENCE_ROOT, REFERENCE_INSTALL) if os.path.isdir(os.path.join(r, "nanograd"))), None)
    if root is None:
        _REF = {}
        return None
    import importlib
    import warnings
    shim = os.path.join(ROOT, "tests", "_shim")
    saved_path = list(sys.path)
    dont_write = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # /root/reference is read-only
    try:
        sys.path.insert(0, shim)
        sys.path.insert(0, root)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mods = {
                "root": root,
                "package": importlib.import_module("nanograd"),
                "tensor": importlib.import_module("nanograd.tensor"),
                "module": importlib.import_module("nanograd.nn.module"),
                "optim": importlib.import_module("nanograd.optim.optimizer"),
                "ops_cpu": importlib.import_module("nanograd.nn.ops_cpu"),
                "device": importlib.import_module("nanograd.device"),
            }
    finally:
        sys.path[:] = saved_path
        sys.dont_write_bytecode = dont_write
    _REF = mods
    return mods


# ---------------------------------------------------------------------------- device libs
def use_emulated_library():
    """Bind the host-emulated build of the kernels (tests/emu) — CPU-only test runs."""
    from nanograd_b200 import backend
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    path = build_emu.build()
    if backend.library_path() not in (None, path):
        import pytest
        pytest.skip("a different device library is already bound in this process")
    backend.load_library(path)
    backend.init(0)


def use_cuda_library():
    from nanograd_b200 import backend
    if backend.library_path() not in (None, backend.DEFAULT_LIB):
        import pytest
        pytest.skip("the emulated library is already bound in this process")
    backend.load_library(backend.DEFAULT_LIB)
    backend.init()


# ---------------------------------------------------------------------------- op harness
def run_on_both(shapes, fn, discrete=False, loc=30.0, scale=1.0, backward=True, seed=0, inputs=None):
    """The reference's make_test_ops pattern (tests/helpers.py:86-132) with the NumPy oracle in
    place of PyTorch: same seeded inputs on Device.CPU (oracle) and Device.GPU, loss = out.mean().
    Returns (out_cpu, out_gpu, grads_cpu, grads_gpu) as NumPy arrays."""
    from nanograd_b200.device import Device
    from nanograd_b200.tensor import Tensor
    rng = np.random.RandomState(seed)
    if inputs is None:
        inputs = []
        for shape in shapes:
            if discrete:
                inputs.append(rng.randint(0, 10, shape).astype(np.float32))
            else:
                inputs.append(rng.normal(loc, scale, shape).astype(np.float32))
    results = []
    for device in (Device.CPU, Device.GPU):
        tensors = [Tensor(a.copy(), requires_grad=backward, device=device) for a in inputs]
        out = fn(*tensors)
        grads = None
        if backward:
            out.mean().backward()
            grads = [t.grad.numpy() for t in tensors]
        results.append((out.numpy(), grads))
    return results[0][0], results[1][0], results[0][1], results[1][1]


def check_op(shapes, fn, rtol=RTOL_FP32, **kw):
    out_c, out_g, gr_c, gr_g = run_on_both(shapes, fn, **kw)
    assert_close(out_g, out_c, rtol, "forward")
    if gr_c is not None:
        for i, (gc, gg) in enumerate(zip(gr_c, gr_g)):
            assert_close(gg, gc, rtol, f"grad of input {i}")


# ---------------------------------------------------------------------------- models
def build_mnist_cnn(nn):
    """The README / tests/test_mnist.py:73-85 CNN: Conv2d(1,8,3)-relu-pool-Conv2d(8,16,3)-relu-pool-
    flatten-Linear(400,10)-log_softmax."""
    class CNN(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(1, 8, 3)
            self.conv2 = nn.Conv2d(8, 16, 3)
            self.l1 = nn.Linear(400, 10)

        def forward(self, x):
            x = x.reshape(shape=[-1, 1, 28, 28])
            x = self.conv1(x).relu().max_pool2d()
            x = self.conv2(x).relu().max_pool2d()
            x = x.flatten()
            return self.l1(x).log_softmax()
    return CNN()


def build_vgg(nn, widths=(64, 128, 256), in_ch=3, image=32, hidden=256, classes=10):
    """BASELINE config 2: [Conv3x3(pad 1)-BN-ReLU]x2-MaxPool per stage, then FC."""
    layers, c = [], in_ch
    for w in widths:
        layers += [nn.Conv2d(c, w, 3, 1, 1), nn.BatchNorm2d(w), nn.ReLU(),
                   nn.Conv2d(w, w, 3, 1, 1), nn.BatchNorm2d(w), nn.ReLU(), nn.MaxPool2d(2)]
        c = w
        image //= 2
    layers += [nn.Flatten(), nn.Linear(c * image * image, hidden), nn.ReLU(), nn.Linear(hidden, classes)]
    return nn.Sequential(*layers)


def params_to_numpy(model):
    return [p.numpy().copy() for p in model.parameters()]


# ---------------------------------------------------------------------------- case comparison
def compare_case(name, got, want, rtol, exact_keys=(), params_on_model_scale=False, mask_keys=None,
                 trajectory_rtol=None, grads_on_model_scale=False):
    """Compare one tests/cases.py result dict with its reference values.

    Tolerance: scale-relative allclose per array, each gradient on ITS OWN scale (see RTOL_FP32
    above).  Two documented refinements for whole-model cases, both about quantities that are
    analytically zero:
      * a parameter gradient whose reference maximum is below 1e-6 of the model's largest
        gradient (the bias of a conv that feeds BatchNorm: exactly zero analytically, ~1e-18 in
        the float64 oracle, rounding noise ~1e-9 on any float32 path) is compared on the scale of
        the model's largest gradient instead of its own;
      * after an Adam/AdamW step such a parameter moves by lr*g/(|g|+eps), i.e. by amplified
        noise, on every implementation; those parameters are excluded from the post-step check.
    ``grads_on_model_scale`` (opt-in TF32 mode, whole-model cases only): gradients are compared on the
    scale of the model's largest gradient.  The 1e-2 bar is a per-operator bar; through four TF32
    conv layers and their BatchNorm backward passes (which subtract means, i.e. cancel) the 3e-4
    per-op rounding compounds to ~5 % of the first layer's small gradient (measured: 3.6e-4 absolute
    against a first-layer maximum of 6.2e-3 and a model maximum of 0.06).
    ``mask_keys``: arrays whose support (!= 0) must match bit-exactly besides the value check.
    ``trajectory_rtol``: looser bar for the end of a multi-step trajectory (see cases.TRAJECTORY_RTOL).
    """
    if mask_keys is None:
        import cases as _cases
        mask_keys = _cases.MASK_KEYS.get(name, ())
        if trajectory_rtol is None:
            trajectory_rtol = _cases.TRAJECTORY_RTOL.get(name)
    assert set(got.keys()) == set(want.keys()), f"{name}: keys differ: {sorted(got)} vs {sorted(want)}"
    grad_keys = [k for k in want.keys() if k.startswith("grad")]
    gmax = max([float(np.max(np.abs(want[k]))) for k in grad_keys], default=0.0)
    # TF32-mode whole-model runs: post-step weights are compared on the scale of the model's largest
    # parameter tensor, like gradients are on the largest gradient (a zero-initialised BatchNorm beta
    # after two steps IS a gradient, scaled by lr)
    pmax = max([float(np.max(np.abs(want[k]))) for k in want.keys() if k.startswith("param")], default=0.0)
    for key in want.keys():
        w, g = np.asarray(want[key]), np.asarray(got[key])
        what = f"{name}[{key}]"
        if key in exact_keys:
            assert_exact(g.astype(np.float64), w.astype(np.float64), what)
            continue
        if key in mask_keys:
            assert_exact(g != 0, w != 0, what + " support (equality mask)")
        if key.startswith("grad"):
            assert g.shape == w.shape, what
            own = float(np.max(np.abs(w)))
            scale = own if own >= 1e-6 * gmax else gmax  # analytically-zero gradients: see docstring
            if grads_on_model_scale:
                scale = max(own, gmax)
            np.testing.assert_allclose(g, w, rtol=rtol, atol=rtol * scale, err_msg=what)
            continue
        if trajectory_rtol is not None and (key.startswith("param") or key in ("loss_last", "loss_history")):
            gk = "grad" + key[len("param"):]
            if key.startswith("param") and gk in want and float(np.max(np.abs(want[gk]))) < 1e-6 * gmax:
                continue
            assert_close(g, w, max(rtol, trajectory_rtol), what)
            continue
        if key.startswith("param") and gmax > 0:
            gk = "grad" + key[len("param"):]
            if gk in want and float(np.max(np.abs(want[gk]))) < 1e-6 * gmax:
                continue  # analytically-zero gradient: see docstring
            if params_on_model_scale:
                assert g.shape == w.shape, what
                np.testing.assert_allclose(g, w, rtol=rtol, atol=rtol * pmax, err_msg=what)
                continue
        assert_close(g, w, rtol, what)


def device_framework(device):
    """tests/cases.py handle for this repo's Tensor on the given device."""
    import cases
    import nanograd_b200.nn.module as nn
    import nanograd_b200.optim.optimizer as optim
    from nanograd_b200.tensor import Tensor
    return cases.Framework(Tensor, nn, optim, device, lambda t: np.array(t.numpy(), copy=True))
