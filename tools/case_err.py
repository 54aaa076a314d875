"""Per-array error of one tests/cases.py case against its golden vectors, per math mode (diagnostic).
python tools/case_err.py CASE [modes...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import helpers as H  # noqa: E402
from nanograd_b200 import backend  # noqa: E402
from nanograd_b200.device import Device  # noqa: E402


def main():
    name = sys.argv[1]
    modes = sys.argv[2:] or ["simt", "tf32x3", "f16x3", "fp32"]
    H.use_cuda_library()
    fn, kwargs = cases.all_cases()[name]
    want = H.golden(name)
    for mode in modes:
        backend.set_math_mode(mode)
        with np.errstate(all="ignore"):
            got = fn(H.device_framework(Device.GPU), **kwargs)
        worst = []
        for k in want.files:
            w, g = np.asarray(want[k], dtype=np.float64), np.asarray(got[k], dtype=np.float64)
            s = float(np.abs(w).max())
            if s < 1e-12:
                continue   # analytically-zero quantities (conv bias under BatchNorm)
            worst.append((float(np.abs(g - w).max()) / max(s, 1e-300), k, s))
        worst.sort(reverse=True)
        print(mode, " ".join(f"{k}:{e:.1e}(max {s:.1e})" for e, k, s in worst[:8]), flush=True)


if __name__ == "__main__":
    main()
