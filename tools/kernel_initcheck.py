"""Uninitialised-memory check of the kernel sources, without a GPU.

Under NG_EMU_POISON_NAN=1 the host emulation of the kernels (tests/emu) fills every block the library's caching
allocator hands out — fresh or recycled — with NaNs.  A kernel that accumulates into, or reads, device memory nobody
wrote (the bug `compute-sanitizer --tool initcheck` finds on a GPU; the reference's OpenCL kernels K2 / K5 had it, SURVEY
§2.4) then yields NaN instead of whatever the block happened to contain, and the golden-vector parity tests fail.

    python tools/kernel_initcheck.py [pytest selection ...]

1. self-check: a freshly allocated buffer and a recycled one must read back as NaN;
2. the emulated tests of the selection run with the poison on.
Prints `INITCHECK selfcheck=ok tests=<pytest summary>`; exit code 0 iff the self-check held and every test passed.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DEFAULT = ["tests/test_device_parity.py", "tests/test_narrow_conv.py", "tests/test_first_layer_conv.py",
           "tests/test_lazy_batchnorm.py", "tests/test_fused_layers.py", "tests/test_edge_shapes.py",
           "tests/test_optimizer_placement.py", "tests/test_checkpoint.py", "tests/test_data.py",
           "tests/test_integration_reference.py"]

SELFCHECK = r"""
import sys
sys.path[:0] = [%(root)r, %(root)r + "/tests", %(root)r + "/tests/emu"]
import numpy as np
import build_emu
from nanograd_b200 import backend as B
B.load_library(build_emu.build()); B.init(0)
from nanograd_b200.nn.buffer import DeviceBuffer
a = DeviceBuffer((1000,)); fresh = a.numpy()
B.lib().ng_fill(a.ptr, 1.0, 1000); ptr = a.ptr; del a
b = DeviceBuffer((1000,)); recycled = b.numpy()
print("SELF", int(np.isnan(fresh).all()), int(b.ptr == ptr), int(np.isnan(recycled).all()))
"""


def main():
    selection = sys.argv[1:] or DEFAULT
    env = dict(os.environ, NG_EMU_POISON_NAN="1")
    res = subprocess.run([sys.executable, "-c", SELFCHECK % {"root": ROOT}], capture_output=True, text=True, env=env)
    flags = res.stdout.strip().split()[-3:] if "SELF" in res.stdout else []
    ok = flags == ["1", "1", "1"]
    print("self-check: fresh block all NaN, block recycled at the same address, recycled block all NaN -> %s: %s" %
          (flags, "ok" if ok else "FAILED " + res.stderr[-500:]), flush=True)
    run = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "not gpu", "-p", "no:cacheprovider"] + selection,
                         capture_output=True, text=True, env=env, cwd=ROOT)
    tail = [ln for ln in run.stdout.strip().splitlines() if ln.strip()][-1] if run.stdout.strip() else "no output"
    if run.returncode != 0:
        print(run.stdout[-3000:])
    print("INITCHECK selfcheck=%s tests=%r" % ("ok" if ok else "failed", tail))
    return 0 if ok and run.returncode == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
