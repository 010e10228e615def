"""In-tree build of libarflow_b200.so (nvcc, sm_100a only).

    python -m arflow_b200.build [--force]

Every csrc/*.cu is compiled to build/*.o and linked into arflow_b200/libarflow_b200.so, next to
the Python shim, so the binary travels with the source tree (no JIT cache, no site-packages).
nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(PKG, "libarflow_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3",
    "-Xptxas", "-v",
]
if os.environ.get("ARF_RELEASE") == "1":      # compile the test hooks of arf_debug_set out (csrc/common.cuh)
    NVCC_FLAGS.append("-DARF_TEST_HOOKS=0")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "arflow_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, hdr_mtime, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    path = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) >= max(os.path.getmtime(path), hdr_mtime)):
        return obj, False, ""
    cmd = [NVCC, *NVCC_FLAGS, "-c", path, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    log = p.stdout + p.stderr
    with open(obj[:-2] + ".ptxas.log", "w") as f:
        f.write(log)
    if verbose:
        print(log)
    return obj, True, log


def build_library(force=False, verbose=False):
    """Compile + link; returns the path of the shared library."""
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    hdr_mtime = _deps_mtime()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force, hdr_mtime, verbose), srcs))
    objs = [r[0] for r in res]
    rebuilt = any(r[1] for r in res)
    if rebuilt or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    return LIB


if __name__ == "__main__":
    lib = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(lib)
