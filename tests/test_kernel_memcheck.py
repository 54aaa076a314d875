"""Out-of-bounds check of the kernel sources (tools/kernel_memcheck.py --quick): the host build of the kernels under
AddressSanitizer with exactly-sized device allocations (no pool rounding behind a tensor).  The tool first proves it
would see a kernel writing one element past its buffer, then runs the emulated tests of the direct / narrow convolution
kernels, the BatchNorm split kernels, the fused layers and the optimizer kernels: no report may involve this library.
(The full selections — golden-vector parity cases, edge shapes, the reference-bound integration tests and the
reference's own GPU tests — are recorded in profiles/r2_memcheck.txt.)"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_out_of_bounds_access_in_the_emulated_kernels():
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan.so is not installed")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kernel_memcheck.py"), "--quick"],
                         capture_output=True, text=True, timeout=2400, cwd=ROOT)
    m = re.search(r"MEMCHECK selfcheck=(\w+) tests='([^']*)' errors=(\d+)", res.stdout)
    assert m, res.stdout[-3000:] + res.stderr[-3000:]
    assert m.group(1) == "ok", "the detector did not see the deliberate overrun: " + res.stdout[-1500:]
    assert int(m.group(3)) == 0, res.stdout[-3000:]
    assert "passed" in m.group(2) and "failed" not in m.group(2) and res.returncode == 0, res.stdout[-3000:]
