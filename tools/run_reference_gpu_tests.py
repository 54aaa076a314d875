"""Run the REFERENCE'S OWN Device.GPU test classes against this backend.

PABannier/nanograd ships unittest classes for its (OpenCL) GPU backend — tests/test_ops.py::TestGPU (39 operator tests:
conv2d over 16 filter-size x stride combinations, conv1d, max / avg pooling, pad, slice, broadcasting, reductions,
log_softmax, one_hot ...), tests/test_step.py::TestStepGPU (SGD with and without momentum, Adam, AdamW) and
tests/test_module.py::TestModuleGPU — each comparing nanograd on Device.GPU with PyTorch on the CPU at rtol 1e-3.
This script imports the unmodified reference from /root/reference, binds this repo's backend into it with
nanograd_b200.integrate.install() and runs exactly those classes, unedited:

    python tools/run_reference_gpu_tests.py [--lib emu|cuda] [module ...]     (default: all three modules, emu)

--lib emu   the host emulation of the kernel sources (tests/emu) — no GPU needed; what the CPU test-suite runs
--lib cuda  the sm_100a library on a B200 (needs /root/reference, i.e. a GPU in the build container)
Prints one summary line `REFTESTS run=N failures=F errors=E skipped=S`; exit code 0 iff everything passed.
"""
import argparse
import importlib
import os
import sys
import unittest
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("NANOGRAD_REFERENCE_ROOT", "/root/reference")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default="emu", choices=["emu", "cuda"])
    ap.add_argument("--primitives-only", action="store_true",
                    help="bind the boundary only (integrate.install(fused=False)): every layer runs the reference's own "
                         "composition over the 21 registered primitives")
    ap.add_argument("modules", nargs="*", default=["tests.test_ops", "tests.test_step", "tests.test_module"])
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(REFERENCE, "tests")):
        print("REFTESTS unavailable: %s/tests not present" % REFERENCE)
        return 3
    # the reference's `tests` package (it has an __init__.py) must win over this repo's tests/ directory
    sys.path[:0] = [REFERENCE, os.path.join(ROOT, "tests", "_shim"), ROOT]
    sys.dont_write_bytecode = True          # /root/reference is read-only
    warnings.simplefilter("ignore")
    if args.lib == "emu":
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        library = build_emu.build()
    else:
        from nanograd_b200 import backend
        library = backend.DEFAULT_LIB
    import nanograd
    import nanograd_b200.integrate as integrate
    integrate.install(nanograd, fused=not args.primitives_only, library=library)

    suite = unittest.TestSuite()
    for name in args.modules:
        mod = importlib.import_module(name)
        assert os.path.abspath(mod.__file__).startswith(os.path.abspath(REFERENCE)), mod.__file__
        for attr in sorted(dir(mod)):
            obj = getattr(mod, attr)
            if isinstance(obj, type) and issubclass(obj, unittest.TestCase) and "GPU" in attr:
                suite.addTests(unittest.defaultTestLoader.loadTestsFromTestCase(obj))
    res = unittest.TextTestRunner(verbosity=0, stream=sys.stderr).run(suite)
    print("REFTESTS run=%d failures=%d errors=%d skipped=%d" % (res.testsRun, len(res.failures), len(res.errors), len(res.skipped)))
    return 0 if res.wasSuccessful() and res.testsRun > 0 else 1


if __name__ == "__main__":
    sys.exit(main())
