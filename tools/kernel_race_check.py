"""Shared-memory race check of the kernel sources, without a GPU.

The kernels of nanograd_b200/csrc (all but the tcgen05 files) compile for the host under -DNG_EMU (tests/emu).  Here
they are built with ThreadSanitizer and run with one REAL host thread per CUDA thread (NG_EMU_THREADS=1), __syncthreads
and the warp exchanges being std::barrier: a kernel that reads a shared-memory slot another thread wrote without a
barrier in between — the bug that shows up on a GPU as a rare wrong value — is reported as a data race.

    python tools/kernel_race_check.py [--quick] [pytest selection ...]

1. self-check: a 64-thread kernel with the barrier left out must be reported, the same kernel with it must not;
2. the emulated tests of the selection (default: direct / narrow convolutions, BatchNorm split kernels, fused layers,
   and the golden-vector parity cases) run under the instrumented library; every ThreadSanitizer report is counted.
Prints `RACECHECK selfcheck=ok tests=<pytest summary> races=N`; exit code 0 iff the self-check held and N == 0.
Diagnostic only (compute-sanitizer --tool racecheck is the on-GPU counterpart).
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
static void kernel(float* out, bool with_sync) {
  __shared__ float buf[64];
  buf[threadIdx.x] = (float)threadIdx.x;
  if (with_sync) __syncthreads();
  out[threadIdx.x] = buf[63 - threadIdx.x];   // a slot another thread wrote
}
int main(int argc, char** argv) {
  static float out[64];
  const bool with_sync = argc > 1;
  ng_emu::launch(dim3(1), dim3(64), 0, [=]() { kernel(out, with_sync); });
  printf("done %f\n", out[0]);
  return 0;
}
"""


def libtsan():
    path = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(path) or not os.path.exists(path):
        raise SystemExit("libtsan.so not found (gcc -print-file-name=libtsan.so -> %r)" % path)
    return path


def selfcheck(scratch):
    src = os.path.join(scratch, "racy.cpp")
    with open(src, "w") as f:
        f.write(RACY)
    exe = os.path.join(scratch, "racy")
    subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-Wno-tsan", "-DNG_EMU", "-pthread", "-I", EMU,
                    src, os.path.join(EMU, "ng_emu.cpp"), "-o", exe], check=True)
    env = dict(os.environ, NG_EMU_THREADS="1", TSAN_OPTIONS="exitcode=0")
    racy = subprocess.run([exe], capture_output=True, text=True, env=env)
    clean = subprocess.run([exe, "sync"], capture_output=True, text=True, env=env)
    n_racy = (racy.stdout + racy.stderr).count("WARNING: ThreadSanitizer: data race")
    n_clean = (clean.stdout + clean.stderr).count("WARNING: ThreadSanitizer: data race")
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
    lib = build_emu.build(tsan=True)
    scratch = tempfile.mkdtemp(prefix="ng_racecheck_")
    ok, n_racy, n_clean = selfcheck(scratch)
    print("self-check: kernel without its barrier -> %d report(s), with it -> %d: %s" % (n_racy, n_clean, "ok" if ok else "FAILED"),
          flush=True)
    log_prefix = os.path.join(scratch, "tsan")
    # OpenBLAS' own worker threads (NumPy, uninstrumented) would be reported too: keep the oracle single-threaded
    env = dict(os.environ, NG_EMU_THREADS="1", NG_EMU_TSAN="1", LD_PRELOAD=libtsan(), OPENBLAS_NUM_THREADS="1",
               OMP_NUM_THREADS="1", TSAN_OPTIONS="log_path=%s exitcode=0 report_signal_unsafe=0" % log_prefix)
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "not gpu", "-p", "no:cacheprovider"] + selection,
                         capture_output=True, text=True, env=env, cwd=ROOT)
    tail = [ln for ln in res.stdout.strip().splitlines() if ln.strip()][-1] if res.stdout.strip() else "no output"
    races, foreign, where = 0, 0, set()
    for path in glob.glob(log_prefix + ".*"):
        for report in open(path, errors="replace").read().split("==================")[1:]:
            if "WARNING: ThreadSanitizer" not in report:
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
    print("RACECHECK selfcheck=%s tests=%r races=%d" % ("ok" if ok else "failed", tail, races))
    return 0 if ok and races == 0 and res.returncode == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
