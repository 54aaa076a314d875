"""In-tree build of the device library: nvcc -> nanograd_b200/lib/libnanograd_b200.so (sm_100a).

``python -m nanograd_b200.build`` (or ``__graft_entry__.build()``) cross-compiles every file in
``nanograd_b200/csrc`` for sm_100a; no GPU is needed to build.  The shared object is
self-contained (static cudart; NCCL is dlopen()ed at run time) and exposes exactly the C ABI
declared in ``include/nanograd_b200.h``.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libnanograd_b200.so")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]
# NANOGRAD_B200_ABLATE=1 builds the tensor-core kernels with their NG_UMMA_EXP ablation switches
# (development only: the switches cost time on the per-stage paths, so release builds leave them out)
if os.environ.get("NANOGRAD_B200_ABLATE"):
    NVCC_FLAGS.append("-DNG_UMMA_ABLATE")


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the nanograd_b200 device library cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(ROOT, "include", "nanograd_b200.h"))
    return hs


def build_library(force=False, verbose=False):
    """Compile (if stale) and return the path of the shared library."""
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sources()
    dep_time = max(os.path.getmtime(h) for h in headers())
    log_path = os.path.join(LIB_DIR, "ptxas.log")

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), dep_time):
            return obj, ""
        cmd = [nvcc] + ARCH_FLAGS + NVCC_FLAGS + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, res.stdout))
        return obj, "### %s\n%s\n" % (os.path.basename(src), res.stdout)

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [r[0] for r in results]
    logs = "".join(r[1] for r in results)
    if logs:
        with open(log_path, "a") as f:
            f.write(logs)
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs):
        cmd = [nvcc] + ARCH_FLAGS + ["-shared", "-o", LIB_PATH] + objs + ["-cudart", "static", "-ldl", "-lpthread"]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s" % res.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
