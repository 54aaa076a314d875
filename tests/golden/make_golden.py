"""Generate tests/golden/*.npz by running tests/cases.py on the REAL reference.

Run in the build container (the reference lives read-only at /root/reference; it cannot travel
to the GPU box, the vectors can):

    python tests/golden/make_golden.py            # writes one .npz per case + MANIFEST.json
    python tests/golden/make_golden.py NAME ...   # only these cases (the manifest keeps the others)

The reference is imported unmodified through tests/helpers.load_reference() (a stub graphviz
module is the only shim; pyopencl's absence just leaves its GPU namespace empty).
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import helpers  # noqa: E402


def reference_framework():
    ref = helpers.load_reference()
    if ref is None:
        raise SystemExit("the reference is not available at /root/reference")
    return cases.Framework(ref["tensor"].Tensor, ref["module"], ref["optim"], ref["device"].Device.CPU,
                           lambda t: np.array(t.data, copy=True))


def main():
    fw = reference_framework()
    manifest = {}
    only = set(sys.argv[1:])
    if only and os.path.exists(os.path.join(HERE, "MANIFEST.json")):
        with open(os.path.join(HERE, "MANIFEST.json")) as f:
            manifest = json.load(f)["cases"]
    warnings.simplefilter("ignore")
    for name, (fn, kwargs) in cases.all_cases().items():
        if only and name not in only:
            continue
        with np.errstate(all="ignore"):
            out = fn(fw, **kwargs)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        manifest[name] = {k: [list(v.shape), str(v.dtype)] for k, v in out.items()}
        print(f"{name}: {len(out)} arrays, {os.path.getsize(path)} bytes")
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump({"source": "PABannier/nanograd 1.0.4 (/root/reference), NumPy CPU path",
                   "numpy": np.__version__, "cases": manifest}, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
