"""TEST INFRASTRUCTURE ONLY: compile nanograd_b200/csrc/*.cu for the host (-DNG_EMU) with g++.

The result (tests/emu/_build/libng_emu.so) exposes the same C ABI as the CUDA library and lets
the CPU-only test run exercise kernel indexing and the ctypes marshalling.  It is never
imported by the nanograd_b200 package; tests load it explicitly through
``nanograd_b200.backend.load_library(path)``.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "nanograd_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libng_emu.so")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu") and not f.startswith("umma_"))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs += [os.path.join(HERE, "ng_emu.h"), os.path.join(HERE, "ng_emu.cpp"),
             os.path.join(ROOT, "include", "nanograd_b200.h")]
    return hdrs


def build(force=False, verbose=False, tsan=None, asan=None, ubsan=None):
    """Build (or reuse) the emulated library.  Safe under pytest-xdist: one process builds at a time (file lock),
    objects and the library are written under temporary names and renamed into place.

    tsan=True (or NG_EMU_TSAN=1): a ThreadSanitizer-instrumented build in tests/emu/_build_tsan, for
    tools/kernel_race_check.py — run with NG_EMU_THREADS=1 (one host thread per CUDA thread) and libtsan preloaded,
    it reports shared-memory races (a missing __syncthreads) in the kernel sources."""
    import fcntl
    global OUT_DIR, LIB
    if tsan is None:
        tsan = os.environ.get("NG_EMU_TSAN", "0") == "1"
    if asan is None:
        asan = os.environ.get("NG_EMU_ASAN", "0") == "1"
    # asan=True (or NG_EMU_ASAN=1): AddressSanitizer build in tests/emu/_build_asan whose cudaMalloc hands out
    # exactly-sized blocks (tools/kernel_memcheck.py: out-of-bounds reads / writes of the kernels)
    if ubsan is None:
        ubsan = os.environ.get("NG_EMU_UBSAN", "0") == "1"
    # ubsan=True (or NG_EMU_UBSAN=1): UndefinedBehaviorSanitizer build in tests/emu/_build_ubsan — misaligned vector
    # accesses (an illegal-address fault on the GPU), signed index overflow, bad shifts (tools/kernel_ubcheck.py)
    out_dir = os.path.join(HERE, "_build_tsan" if tsan else "_build_asan" if asan else "_build_ubsan" if ubsan else "_build")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        OUT_DIR, LIB = out_dir, os.path.join(out_dir, "libng_emu.so")
        return _build_locked(force, verbose, tsan, asan and not tsan, ubsan and not tsan and not asan)


def _build_locked(force, verbose, tsan=False, asan=False, ubsan=False):
    srcs = _sources() + [os.path.join(HERE, "ng_emu.cpp")]
    newest = max(os.path.getmtime(p) for p in srcs + _deps())
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    flags = ["-std=c++20", "-O2", "-g", "-DNG_EMU", "-fPIC", "-pthread", "-I", HERE, "-I", CSRC,
             "-Wno-unused-result", "-Wno-attributes"]
    link = []
    if tsan:
        flags = [f for f in flags if f != "-O2"] + ["-O1", "-fsanitize=thread", "-Wno-tsan"]
        link = ["-fsanitize=thread"]
    if asan:
        flags = [f for f in flags if f != "-O2"] + ["-O1", "-fsanitize=address", "-fno-omit-frame-pointer", "-DNG_EMU_EXACT_ALLOC"]
        link = ["-fsanitize=address"]
    if ubsan:
        flags = [f for f in flags if f != "-O2"] + ["-O1", "-fsanitize=undefined", "-fno-omit-frame-pointer"]
        link = ["-fsanitize=undefined"]
    objs = []

    def compile_one(src):
        obj = os.path.join(OUT_DIR, os.path.basename(src) + ".o")
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) >= max(os.path.getmtime(src), max(os.path.getmtime(d) for d in _deps()))):
            return obj
        tmp = obj + ".tmp%d" % os.getpid()
        cmd = ["g++", "-x", "c++"] + flags + ["-c", src, "-o", tmp]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        os.replace(tmp, obj)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, srcs))
    tmp = LIB + ".tmp%d" % os.getpid()
    subprocess.run(["g++", "-shared", "-pthread"] + link + ["-o", tmp] + objs, check=True)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
