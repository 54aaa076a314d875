"""Undefined-behaviour check of the kernel sources (tools/kernel_ubcheck.py): the host build of the kernels under
UndefinedBehaviorSanitizer.  The tool first proves it would see a float4 load from a 4-byte-aligned address (an
illegal-address fault on the GPU, silently fine on x86), then runs the emulated tests — parity cases, direct / narrow
convolutions with ragged widths (their 128- / 64-bit loads are alignment-gated), BatchNorm split kernels, fused layers,
edge shapes, loader, checkpoints: no misaligned vector access, signed index overflow or bad shift may be reported."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_undefined_behaviour_in_the_emulated_kernels():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kernel_ubcheck.py")],
                         capture_output=True, text=True, timeout=2400, cwd=ROOT)
    m = re.search(r"UBCHECK selfcheck=(\w+) tests='([^']*)' reports=(\d+)", res.stdout)
    assert m, res.stdout[-3000:] + res.stderr[-3000:]
    assert m.group(1) == "ok", "the detector did not see the deliberately misaligned load: " + res.stdout[-1500:]
    assert int(m.group(3)) == 0, res.stdout[-3000:]
    assert "passed" in m.group(2) and "failed" not in m.group(2) and res.returncode == 0, res.stdout[-3000:]
