"""The drop-in boundary, executed: the UNMODIFIED reference's ``Tensor`` / ``Module`` /
``Optimizer`` (PABannier/nanograd, imported from /root/reference or from its pip-installed copy
oracle/_ref on the GPU box) bound to this repo's CUDA operator namespace by
``nanograd_b200.integrate.install()`` — the binding INTEGRATION.md documents.

Every parity case of tests/cases.py is run through the reference's own classes on
``Device.GPU`` and compared with the golden vectors the same reference produced on its NumPy
path: with the fused routes (one device op per layer) and with the boundary alone
(``fused=False``: the reference's own compositions over the 21 registered primitives).
``-m gpu`` runs on the sm_100a library; the CPU suite runs a subset on the host emulation.
"""
import os
import re

import numpy as np
import pytest

import cases
import helpers as H

CASES = cases.all_cases()

EMU_FUSED = ["conv2d_k3_s1", "conv1d_k3_s2", "matmul", "maxpool2d_p2", "avgpool2d_p2", "maxpool1d_p2",
             "logsoftmax_nll", "elementwise", "reductions", "shapes", "batchnorm2d", "batchnorm1d", "mnist_cnn_sgd",
             "vgg_bn_adam", "attention_adamw", "tinynet_sgdm", "tinynet_adamw2", "linear_stack",
             "activation_mlp_sgdm", "conv2d_wide_c64_k32_p1"]
EMU_PRIMITIVES = ["maxpool2d_p3", "avgpool2d_p4", "logsoftmax_nll", "batchnorm2d", "batchnorm1d", "vgg_bn_sgd",
                  "tinynet_adam", "activation_mlp_adam", "conv2d_wide_c64_k32_p1"]


def _bound_reference(fused):
    ref = H.load_reference()
    if ref is None:
        pytest.skip("the reference is neither at /root/reference nor installed in oracle/_ref "
                    "(tools/make_oracle_ref.sh)")
    from nanograd_b200 import integrate
    integrate.install(ref["package"])
    integrate.set_fused(fused)
    return ref


def _to_numpy(t):
    d = t.data
    return np.array(d.numpy() if hasattr(d, "numpy") else d, copy=True)


def _run(name, fused):
    ref = _bound_reference(fused)
    fw = cases.Framework(ref["tensor"].Tensor, ref["module"], ref["optim"], ref["device"].Device.GPU, _to_numpy)
    fn, kwargs = CASES[name]
    try:
        with np.errstate(all="ignore"):
            got = fn(fw, **kwargs)
    finally:
        from nanograd_b200 import integrate
        integrate.set_fused(True)
    want = H.golden(name)
    H.compare_case(name, got, {k: want[k] for k in want.files}, H.RTOL_FP32, cases.EXACT_KEYS.get(name, ()))


@pytest.mark.parametrize("name", EMU_FUSED)
def test_reference_classes_on_emulated_backend(name):
    H.use_emulated_library()
    _run(name, fused=True)


@pytest.mark.parametrize("name", EMU_PRIMITIVES)
def test_reference_classes_on_emulated_backend_primitives_only(name):
    H.use_emulated_library()
    _run(name, fused=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_classes_on_b200(name):
    H.use_cuda_library()
    _run(name, fused=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(n for n in CASES if not n.startswith(("conv2d_k", "conv1d_k", "tinynet"))))
def test_reference_classes_on_b200_primitives_only(name):
    H.use_cuda_library()
    _run(name, fused=False)


# ---------------------------------------------------------------------------- the documented snippet
def _integration_md_snippets():
    text = open(os.path.join(H.ROOT, "INTEGRATION.md")).read()
    return re.findall(r"```python\n# \[executed by tests/test_integration_reference.py\]\n(.*?)```", text, re.S)


def _exec_snippets():
    ref = _bound_reference(True)  # makes `import nanograd` resolve to the reference
    snippets = _integration_md_snippets()
    assert snippets, "INTEGRATION.md lost its executable snippet"
    scope = {}
    for code in snippets:
        exec(compile(code, "INTEGRATION.md", "exec"), scope)
    assert scope["loss_history"][-1] < scope["loss_history"][0], scope["loss_history"]
    assert type(scope["model"].conv1.weight) is ref["tensor"].Tensor
    assert scope["model"].conv1.weight.device == ref["device"].Device.GPU


def test_integration_md_snippet_runs_verbatim_emulated():
    H.use_emulated_library()
    _exec_snippets()


@pytest.mark.gpu
def test_integration_md_snippet_runs_verbatim_on_b200():
    H.use_cuda_library()
    _exec_snippets()


def test_optimizer_built_before_gpu_move():
    """The reference's flow (tests/test_mnist.py:95-96): optimizer on a CPU model, then model.gpu()."""
    H.use_emulated_library()
    ref = _bound_reference(True)
    Tensor, nn, optim, Device = ref["tensor"].Tensor, ref["module"], ref["optim"], ref["device"].Device
    for make in (lambda p: optim.SGD(p, lr=0.01, momentum=0.9), lambda p: optim.Adam(p), lambda p: optim.AdamW(p)):
        np.random.seed(0)
        model = nn.Sequential(nn.Linear(6, 5), nn.ReLU(), nn.Linear(5, 3))
        opt = make(model.parameters())
        model.gpu()
        before = [_to_numpy(p) for p in model.parameters()]
        x = Tensor(np.random.randn(4, 6).astype(np.float32), device=Device.GPU)
        model(x).sum().backward()
        opt.step()
        after = [_to_numpy(p) for p in model.parameters()]
        assert any(np.abs(a - b).max() > 0 for a, b in zip(after, before))
