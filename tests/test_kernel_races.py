"""Shared-memory race check of the kernel sources (tools/kernel_race_check.py --quick): the host build of the kernels
instrumented with ThreadSanitizer, one real host thread per CUDA thread, __syncthreads / warp exchanges as real
barriers.  The tool first proves it would see a missing barrier (a deliberately racy kernel is reported, its
synchronised twin is not), then runs the emulated tests of the direct / narrow convolution kernels, the BatchNorm split
kernels, the fused layers and the optimizer kernels: no report may involve this library.  (The full selection, with
the golden-vector parity cases, is recorded in profiles/r2_race_check.txt.)"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_shared_memory_races_in_the_emulated_kernels():
    tsan = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(tsan) or not os.path.exists(tsan):
        pytest.skip("libtsan.so is not installed")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kernel_race_check.py"), "--quick"],
                         capture_output=True, text=True, timeout=2400, cwd=ROOT)
    m = re.search(r"RACECHECK selfcheck=(\w+) tests='([^']*)' races=(\d+)", res.stdout)
    assert m, res.stdout[-3000:] + res.stderr[-3000:]
    assert m.group(1) == "ok", "the detector did not see the deliberately racy kernel: " + res.stdout[-1500:]
    assert int(m.group(3)) == 0, res.stdout[-3000:]
    assert "passed" in m.group(2) and "failed" not in m.group(2) and res.returncode == 0, res.stdout[-3000:]
