"""Undefined-behaviour check of the kernel sources, without a GPU.

The kernels of nanograd_b200/csrc (all but the tcgen05 files) compile for the host under -DNG_EMU (tests/emu).  Built
with UndefinedBehaviorSanitizer they report what an x86 host silently tolerates and a GPU does not: a 128- / 64-bit
vector access through a pointer that is not that aligned (an illegal-address fault on the GPU), signed overflow in
index arithmetic, out-of-range shifts, out-of-bounds indexing of fixed-size arrays.

    python tools/kernel_ubcheck.py [pytest selection ...]

1. self-check: a kernel that loads a float4 from a 4-byte-aligned address must be reported, its aligned twin must not;
2. the emulated tests of the selection (default: parity cases, direct / narrow convolutions with ragged widths,
   BatchNorm split kernels, fused layers, edge shapes, loader, checkpoints) run under the instrumented library.
Prints `UBCHECK selfcheck=ok tests=<pytest summary> reports=N`; exit code 0 iff the self-check held and N == 0.
"""
import glob
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")

PROBE = r"""
#include "ng_emu.h"
#include <cstdio>
static void kernel(float* out, const float* in, int shift) {
  const float4 v = *reinterpret_cast<const float4*>(in + 4 * threadIdx.x + shift);
  out[threadIdx.x] = v.x + v.w;
}
int main(int argc, char** argv) {
  float *in = nullptr, *out = nullptr;
  cudaMalloc(&in, (4 * 32 + 4) * sizeof(float));
  cudaMalloc(&out, 32 * sizeof(float));
  for (int i = 0; i < 4 * 32 + 4; ++i) in[i] = (float)i;
  const int shift = argc > 1 ? 0 : 1;        // without an argument every load is 4 bytes off a 16-byte boundary
  ng_emu::launch(dim3(1), dim3(32), 0, [=]() { kernel(out, in, shift); });
  printf("done %f\n", out[1]);
  return 0;
}
"""

DEFAULT = ["tests/test_device_parity.py", "tests/test_narrow_conv.py", "tests/test_first_layer_conv.py",
           "tests/test_lazy_batchnorm.py", "tests/test_fused_layers.py", "tests/test_edge_shapes.py",
           "tests/test_optimizer_placement.py", "tests/test_checkpoint.py", "tests/test_data.py"]


def selfcheck(scratch):
    src, exe = os.path.join(scratch, "probe.cpp"), os.path.join(scratch, "probe")
    with open(src, "w") as f:
        f.write(PROBE)
    subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=undefined", "-DNG_EMU", "-pthread", "-I", EMU,
                    src, os.path.join(EMU, "ng_emu.cpp"), "-o", exe], check=True)
    bad = subprocess.run([exe], capture_output=True, text=True)
    good = subprocess.run([exe, "aligned"], capture_output=True, text=True)
    n_bad = (bad.stdout + bad.stderr).count("misaligned address")
    n_good = (good.stdout + good.stderr).count("runtime error")
    return n_bad >= 1 and n_good == 0, n_bad, n_good


def main():
    selection = sys.argv[1:] or DEFAULT
    sys.path.insert(0, EMU)
    import build_emu
    lib = build_emu.build(ubsan=True)
    scratch = tempfile.mkdtemp(prefix="ng_ubcheck_")
    ok, n_bad, n_good = selfcheck(scratch)
    print("self-check: float4 load from a 4-byte-aligned address -> %d report(s), aligned -> %d: %s" %
          (n_bad, n_good, "ok" if ok else "FAILED"), flush=True)
    log_prefix = os.path.join(scratch, "ubsan")
    env = dict(os.environ, NG_EMU_UBSAN="1", UBSAN_OPTIONS="print_stacktrace=1 log_path=%s" % log_prefix)
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "not gpu", "-p", "no:cacheprovider"] + selection,
                         capture_output=True, text=True, env=env, cwd=ROOT)
    tail = [ln for ln in res.stdout.strip().splitlines() if ln.strip()][-1] if res.stdout.strip() else "no output"
    reports = []
    for path in glob.glob(log_prefix + ".*"):
        reports += [ln.strip() for ln in open(path, errors="replace") if "runtime error" in ln]
    reports += [ln.strip() for ln in (res.stdout + res.stderr).splitlines() if "runtime error" in ln]
    print("library: %s" % os.path.relpath(lib, ROOT))
    print("pytest: %s (rc %d)" % (tail, res.returncode))
    for ln in sorted(set(reports))[:20]:
        print("  " + ln[:200])
    print("UBCHECK selfcheck=%s tests=%r reports=%d" % ("ok" if ok else "failed", tail, len(reports)))
    return 0 if ok and not reports and res.returncode == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
