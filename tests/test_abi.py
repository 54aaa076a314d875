"""The C-ABI library builds, loads and exports every symbol include/nanograd_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nanograd_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ng_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def library_path():
    from nanograd_b200 import build as ngbuild
    try:
        return ngbuild.build_library()
    except RuntimeError as e:  # nvcc missing
        pytest.skip(str(e))


def test_header_and_bridge_agree():
    from nanograd_b200 import backend
    assert declared_symbols() == backend.EXPORTED_SYMBOLS


def test_library_exports_every_declared_symbol(library_path):
    handle = ctypes.CDLL(library_path)
    missing = [s for s in declared_symbols() if not hasattr(handle, s)]
    assert not missing, f"symbols declared in the header but not exported: {missing}"


def test_library_is_sm100a_and_has_no_torch_dependency(library_path):
    out = subprocess.run(["cuobjdump", "-lelf", library_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
    dynamic = subprocess.run(["readelf", "-d", library_path], capture_output=True, text=True).stdout
    needed = "\n".join(line for line in dynamic.splitlines() if "(NEEDED)" in line)
    assert needed, dynamic
    for forbidden in ("torch", "c10", "cublas", "cudnn", "nccl"):
        assert forbidden not in needed.lower(), f"unexpected link-time dependency on {forbidden}"


def test_bridge_binds_all_prototypes(library_path):
    # in a subprocess: binding the CUDA library here would make the emulated parity tests of the
    # same pytest process skip (one device library per process)
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from nanograd_b200 import backend\n"
            "lib = backend.load_library(%r)\n"
            "assert all(getattr(lib, n) is not None for n in backend.EXPORTED_SYMBOLS)\n"
            "print('BOUND', len(backend.EXPORTED_SYMBOLS))\n") % (ROOT, library_path)
    res = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert "BOUND" in res.stdout, res.stderr


def test_missing_library_fails_loudly(tmp_path):
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from nanograd_b200 import backend\n"
            "try:\n"
            "    backend.load_library(%r)\n"
            "except backend.DeviceLibraryError as e:\n"
            "    print('RAISED', e)\n") % (ROOT, str(tmp_path / "nope.so"))
    out = subprocess.run(["python", "-c", code], capture_output=True, text=True).stdout
    assert "RAISED" in out and "no CPU fallback" in out


def test_cpu_tensor_ops_are_not_silently_served():
    """Device.CPU arithmetic is the reference's job; the package itself registers GPU ops only."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import numpy as np\n"
            "from nanograd_b200.tensor import Tensor\n"
            "from nanograd_b200.device import Device\n"
            "print('cpu_ops', len(Tensor.ops[Device.CPU]))\n"
            "try:\n"
            "    Tensor(np.ones(3, dtype=np.float32)) + 1\n"
            "except Exception as e:\n"
            "    print('RAISED', e)\n") % ROOT
    out = subprocess.run(["python", "-c", code], capture_output=True, text=True).stdout
    assert "cpu_ops 0" in out and "RAISED" in out
