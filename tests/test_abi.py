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


def _header_prototypes():
    """{name: [C parameter type, ...]} parsed from the header (comments stripped)."""
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const\s+char\s*\*)\s*(ng_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        params = [p.strip() for p in m.group(2).split(",")]
        if params in ([""], ["void"]):
            params = []
        types = []
        for p in params:
            p = re.sub(r"\s+", " ", p)
            base = p.rsplit(" ", 1)[0] if (" " in p and not p.endswith("*")) else p   # drop the parameter name
            stars = p.count("*")
            types.append(re.sub(r"[\s*]|const", "", base) + "*" * stars)
        protos[m.group(1)] = types
    return protos


def test_header_parameter_lists_match_the_ctypes_prototypes():
    """Same number of parameters, and the same width / pointer-ness for each, in include/nanograd_b200.h and in
    backend._PROTOTYPES: a drifted prototype would corrupt the arguments of a call without any error."""
    from nanograd_b200 import backend
    header = _header_prototypes()
    kinds = {"uint64_t": ctypes.c_uint64, "int64_t": ctypes.c_int64, "int": ctypes.c_int, "float": ctypes.c_float,
             "int32_t": ctypes.c_int32}
    for name, argtypes in backend._PROTOTYPES.items():
        assert name in header, name
        ctypes_of = header[name]
        assert len(ctypes_of) == len(argtypes), f"{name}: header has {len(ctypes_of)} parameters, the bridge {len(argtypes)}"
        for i, (c, a) in enumerate(zip(ctypes_of, argtypes)):
            if c.endswith("*"):
                assert a in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(a, "_type_") and issubclass(a, ctypes._Pointer), \
                    f"{name} parameter {i}: header {c}, bridge {a}"
                base = kinds.get(c.rstrip("*"))
                if base is not None and issubclass(a, ctypes._Pointer):
                    assert ctypes.sizeof(a._type_) == ctypes.sizeof(base), f"{name} parameter {i}: header {c}, bridge {a}"
            else:
                assert c in kinds, f"{name} parameter {i}: unknown C type {c}"
                assert ctypes.sizeof(a) == ctypes.sizeof(kinds[c]) and (a is ctypes.c_float) == (c == "float"), \
                    f"{name} parameter {i}: header {c}, bridge {a}"


def test_every_python_call_site_matches_a_prototype():
    """Static check of the product's Python sources: every `.ng_*(...)` call names a declared entry point and passes
    the declared number of arguments — including the branches only a B200 run takes."""
    import ast
    from nanograd_b200 import backend
    files = []
    for top in ("nanograd_b200", "bench.py", "__graft_entry__.py"):
        path = os.path.join(ROOT, top)
        if os.path.isfile(path):
            files.append(path)
        for d, _, names in os.walk(path):
            files += [os.path.join(d, n) for n in names if n.endswith(".py")]
    problems = []
    for path in files:
        for node in ast.walk(ast.parse(open(path).read())):
            if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr.startswith("ng_")):
                continue
            name = node.func.attr
            if name == "ng_emu_set_collective":   # test-only hook of the emulated library (dist.init(use_device="emu"))
                continue
            if name == "ng_last_error":
                continue
            if name not in backend._PROTOTYPES:
                problems.append(f"{path}:{node.lineno}: {name} is not an entry point")
            elif not any(isinstance(a, ast.Starred) for a in node.args) and len(node.args) != len(backend._PROTOTYPES[name]):
                problems.append(f"{path}:{node.lineno}: {name} called with {len(node.args)} arguments, "
                                f"prototype has {len(backend._PROTOTYPES[name])}")
    assert not problems, "\n".join(problems)
