"""The reference's own Device.GPU unittest classes — written for its OpenCL backend, comparing with PyTorch on the
CPU at rtol 1e-3 — run unedited against this backend bound into the unmodified reference
(tools/run_reference_gpu_tests.py; the emulated kernel sources here, in a subprocess).  43 tests:
tests/test_ops.py::TestGPU (39), tests/test_step.py::TestStepGPU (3), tests/test_module.py::TestModuleGPU (1) — with the
fused routes and with the boundary only — plus the flow of the reference's tests/test_mnist.py (its three models, its
train() / evaluate() loops with Adam built before model.gpu()) on a synthetic idx-gz data set against the same seeded
run on the reference's NumPy path.
Needs /root/reference (its tests are not part of the pip-installed oracle/_ref), so it runs in the build container."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    if not os.path.isdir("/root/reference/tests"):
        pytest.skip("/root/reference/tests is not present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_gpu_tests.py"), "--lib", "emu"] + list(flags),
                         capture_output=True, text=True, timeout=1500, cwd=os.environ.get("TMPDIR", "/tmp"))
    m = re.search(r"REFTESTS run=(\d+) failures=(\d+) errors=(\d+) skipped=(\d+)", res.stdout)
    assert m, res.stdout[-2000:] + res.stderr[-4000:]
    run, failures, errors, skipped = map(int, m.groups())
    assert res.returncode == 0 and failures == 0 and errors == 0, res.stdout[-2000:] + res.stderr[-6000:]
    assert run >= 43 and skipped == 0, m.group(0)
    return res.stdout


def test_reference_gpu_test_classes_and_mnist_flow_pass_on_this_backend():
    out = _run("--mnist-synthetic", "30")
    flows = re.findall(r"REFMNIST (\w+) steps=30 accuracy cpu=([\d.]+) gpu=([\d.]+) (\w+)", out)
    assert [f[0] for f in flows] == ["LinearModel", "CNNModel1d", "CNNModel2d"], out[-1500:]
    assert all(f[3] == "ok" for f in flows), flows


def test_reference_gpu_test_classes_pass_with_the_boundary_only():
    _run("--primitives-only")
