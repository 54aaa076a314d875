"""Uninitialised-memory check (tools/kernel_initcheck.py): with NG_EMU_POISON_NAN=1 the emulated library fills every
block its caching allocator hands out, fresh or recycled, with NaNs; a kernel that consumed memory nobody wrote would
turn the golden-vector parity results into NaN.  The tool proves the poison is in effect, then runs the emulated tests."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_kernel_consumes_unwritten_device_memory():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "kernel_initcheck.py")],
                         capture_output=True, text=True, timeout=2400, cwd=ROOT)
    m = re.search(r"INITCHECK selfcheck=(\w+) tests='([^']*)'", res.stdout)
    assert m, res.stdout[-3000:] + res.stderr[-3000:]
    assert m.group(1) == "ok", res.stdout[-1500:]
    assert "passed" in m.group(2) and "failed" not in m.group(2) and res.returncode == 0, res.stdout[-3000:]
