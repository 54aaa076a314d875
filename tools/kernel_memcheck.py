"""Out-of-bounds check of the kernel sources, without a GPU.

The kernels of nanograd_b200/csrc (all but the tcgen05 files) compile for the host under -DNG_EMU (tests/emu).  Here
they are built with AddressSanitizer, and the library's caching allocator hands out exactly-sized blocks
(-DNG_EMU_EXACT_ALLOC: no 512-byte rounding, no slack behind a tensor): a kernel that reads or writes one element past
a tensor, a workspace or a parameter table — e.g. a 128-bit load on a ragged tail — is reported as a heap-buffer-overflow
(on the GPU the pool's rounding usually hides it).

    python tools/kernel_memcheck.py [--quick] [pytest selection ...]

1. self-check: a kernel that writes element [n] of an n-element device buffer must be reported, its in-bounds twin not;
2. the emulated tests of the selection run under the instrumented library; every AddressSanitizer report is counted.
Prints `MEMCHECK selfcheck=ok tests=<pytest summary> errors=N`; exit code 0 iff the self-check held and N == 0.
Diagnostic only (compute-sanitizer --tool memcheck is the on-GPU counterpart).
"""
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")

RACY = r"""
#include "ng_emu.h"
#include <cstdio>
static void kernel(float* out, int shift) { out[threadIdx.x + shift] = (float)threadIdx.x; }
int main(int argc, char** argv) {
  float* out = nullptr;
  cudaMalloc(&out, 64 * sizeof(float));
  const int shift = argc > 1 ? 0 : 1;        // without an argument thread 63 writes out[64]
  ng_emu::launch(dim3(1), dim3(64), 0, [=]() { kernel(out, shift); });
  printf("done %f\n", out[1]);
  return 0;
}
"""


def libasan():
    path = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(path) or not os.path.exists(path):
        raise SystemExit("libasan.so not found (gcc -print-file-name=libasan.so -> %r)" % path)
    return path


def selfcheck(scratch):
    src = os.path.join(scratch, "racy.cpp")
    with open(src, "w") as f:
        f.write(RACY)
    exe = os.path.join(scratch, "racy")
    subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=address", "-DNG_EMU", "-DNG_EMU_EXACT_ALLOC", "-pthread", "-I", EMU,
                    src, os.path.join(EMU, "ng_emu.cpp"), "-o", exe], check=True)
    env = dict(os.environ, NG_EMU_THREADS="1", ASAN_OPTIONS="detect_leaks=0 halt_on_error=0 exitcode=0")
    racy = subprocess.run([exe], capture_output=True, text=True, env=env)
    clean = subprocess.run([exe, "inbounds"], capture_output=True, text=True, env=env)
    n_racy = (racy.stdout + racy.stderr).count("ERROR: AddressSanitizer: heap-buffer-overflow")
    n_clean = (clean.stdout + clean.stderr).count("ERROR: AddressSanitizer")
    return n_racy >= 1 and n_clean == 0, n_racy, n_clean


def main():
    args = sys.argv[1:]
    quick = "--quick" in args
    selection = [a for a in args if a != "--quick"]
    if not selection:
        selection = ["tests/test_narrow_conv.py", "tests/test_first_layer_conv.py", "tests/test_lazy_batchnorm.py",
                     "tests/test_fused_layers.py", "tests/test_optimizer_placement.py"]
        if not quick:
            selection += ["tests/test_device_parity.py", "tests/test_checkpoint.py", "tests/test_data.py"]
    sys.path.insert(0, EMU)
    import build_emu
    lib = build_emu.build(asan=True)
    scratch = tempfile.mkdtemp(prefix="ng_memcheck_")
    ok, n_racy, n_clean = selfcheck(scratch)
    print("self-check: kernel writing one element past its buffer -> %d report(s), in bounds -> %d: %s" %
          (n_racy, n_clean, "ok" if ok else "FAILED"), flush=True)
    log_prefix = os.path.join(scratch, "asan")
    # real host threads (the fiber executor switches stacks behind AddressSanitizer's back)
    env = dict(os.environ, NG_EMU_THREADS="1", NG_EMU_ASAN="1", LD_PRELOAD=libasan(), OPENBLAS_NUM_THREADS="1",
               OMP_NUM_THREADS="1", ASAN_OPTIONS="log_path=%s detect_leaks=0 halt_on_error=0 exitcode=0" % log_prefix)
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "not gpu", "-p", "no:cacheprovider"] + selection,
                         capture_output=True, text=True, env=env, cwd=ROOT)
    tail = [ln for ln in res.stdout.strip().splitlines() if ln.strip()][-1] if res.stdout.strip() else "no output"
    races, foreign, where = 0, 0, set()
    for path in glob.glob(log_prefix + ".*"):
        for report in re.split(r"(?==+\d+==ERROR: AddressSanitizer)", open(path, errors="replace").read()):
            if "ERROR: AddressSanitizer" not in report:
                continue
            if "libng_emu.so" in report or "ng::" in report or "ng_emu" in report:
                races += 1
                where.update(re.findall(r"#\d+ (ng::[\w:]+)", report))
            else:
                foreign += 1    # e.g. a third-party thread pool: not this library's code
    if foreign:
        print("ignored %d report(s) that do not involve this library" % foreign)
    print("library: %s" % os.path.relpath(lib, ROOT))
    print("pytest: %s (rc %d)" % (tail, res.returncode))
    if races:
        print("functions in the reports: %s" % ", ".join(sorted(where)[:20]))
        print("logs: %s.*" % log_prefix)
    print("MEMCHECK selfcheck=%s tests=%r errors=%d" % ("ok" if ok else "failed", tail, races))
    return 0 if ok and races == 0 and res.returncode == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
