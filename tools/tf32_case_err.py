"""Per-array scale-relative error of a tests/cases.py case in TF32 mode vs its golden file."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, helpers as H
from nanograd_b200 import backend
from nanograd_b200.device import Device
H.use_cuda_library()
for name in sys.argv[1:]:
    for mode in ("fp32", "tf32"):
        backend.set_math_mode(mode)
        fn, kw = cases.all_cases()[name]
        got = fn(H.device_framework(Device.GPU), **kw)
        want = H.golden(name)
        gkeys = [k for k in want.files if k.startswith("grad")]
        gmax = max([float(np.abs(want[k]).max()) for k in gkeys], default=0.0)
        errs = {}
        for k in want.files:
            scale = max(np.abs(want[k]).max(), 1e-30)
            if k.startswith("grad"):
                scale = max(scale, gmax)
            if k.startswith("param") and gmax > 0:
                gk = "grad" + k[5:]
                if gk in want.files and float(np.abs(want[gk]).max()) < 1e-6 * gmax:
                    continue
            errs[k] = float(np.abs(got[k] - want[k]).max() / scale)
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
        print(name, mode, " ".join(f"{k}:{v:.2e}" for k, v in worst))
