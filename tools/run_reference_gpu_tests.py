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
    ap.add_argument("--mnist-synthetic", type=int, default=0, metavar="STEPS",
                    help="also run the reference's tests/test_mnist.py flow (its models, its train() / evaluate() loops, "
                         "the optimizer built before model.gpu()) on Device.GPU for STEPS steps per model, on a synthetic "
                         "idx-gz data set written to a scratch directory (the MNIST files are not shipped)")
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
    ok = res.wasSuccessful() and res.testsRun > 0
    if args.mnist_synthetic > 0:
        ok = mnist_flow(args.mnist_synthetic) and ok
    return 0 if ok else 1


def _write_synthetic_mnist(root, n_train=2048, n_test=512):
    """Ten fixed random 28x28 prototypes plus noise, in the idx-gz layout tests/test_mnist.py:20-38 reads."""
    import gzip
    import numpy as np
    rng = np.random.RandomState(7)
    protos = rng.rand(10, 784) > 0.7
    os.makedirs(os.path.join(root, "data"), exist_ok=True)
    for split, n in (("train", n_train), ("t10k", n_test)):
        y = rng.randint(0, 10, n).astype(np.uint8)
        x = np.clip(protos[y] * 200.0 + rng.randn(n, 784) * 40.0, 0, 255).astype(np.uint8)
        with open(os.path.join(root, "data", f"{split}-images-idx3-ubyte.gz"), "wb") as f:
            f.write(gzip.compress(bytes(16) + x.tobytes()))
        with open(os.path.join(root, "data", f"{split}-labels-idx1-ubyte.gz"), "wb") as f:
            f.write(gzip.compress(bytes(8) + y.tobytes()))


def mnist_flow(steps):
    """tests/test_mnist.py:88-109 on Device.GPU (the reference keeps that class commented out, :112-115): for each of
    its three models, `optim.Adam(model.parameters())` on the host model, `train(..., device=Device.GPU)` (which calls
    model.gpu() after the optimizer exists and reads out / loss back every step), then `evaluate`.  The same seeded run
    on Device.CPU (the reference's NumPy path) is the yardstick."""
    import tempfile
    import numpy as np
    scratch = tempfile.mkdtemp(prefix="ng_ref_mnist_")
    _write_synthetic_mnist(scratch)
    cwd = os.getcwd()
    os.chdir(scratch)                     # tests/test_mnist.py opens 'data/...' relative to the working directory
    try:
        tm = importlib.import_module("tests.test_mnist")
    finally:
        os.chdir(cwd)
    from nanograd.device import Device
    import nanograd.optim.optimizer as optim
    ok = True
    for model_cls in (tm.LinearModel, tm.CNNModel1d, tm.CNNModel2d):
        acc = {}
        for device in (Device.CPU, Device.GPU):
            np.random.seed(0)
            model = model_cls()
            optimizer = optim.Adam(model.parameters(), lr=1e-3)
            tm.train(model, tm.X_train, tm.Y_train, optimizer, steps=steps, batch_size=64, device=device)
            acc[device.name] = float(tm.evaluate(model, tm.X_test, tm.Y_test, device=device))
        # the yardstick is the reference's own NumPy run of the same seeded loop (a few dozen steps do not reach the
        # > 0.91 of tests/test_mnist.py:105-109 on every model): the same test-set accuracy within 2 points
        good = abs(acc["GPU"] - acc["CPU"]) <= 0.02
        ok = ok and good
        print("REFMNIST %s steps=%d accuracy cpu=%.4f gpu=%.4f %s" % (model_cls.__name__, steps, acc["CPU"], acc["GPU"],
                                                                       "ok" if good else "MISMATCH"))
    return ok


if __name__ == "__main__":
    sys.exit(main())
