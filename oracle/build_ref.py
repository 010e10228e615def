"""ORACLE — builds the reference's own CUDA correlation package for sm_100a into oracle/_ref/ (git-ignored).

    python oracle/build_ref.py        # build container only: needs /root/reference

The sources are compiled where they lie (models/correlation_package/correlation_cuda.cc and
correlation_cuda_kernel.cu), unmodified; oracle/ref_compat.h (force-included) bridges the two torch API
removals they trip over.  The reference's setup.py is not used (sm_50-61 only, -std=c++11, CUDA 9 paths).
The resulting correlation_cuda*.so is a torch extension: tests/test_ref_cuda_gpu.py imports it on the GPU box
and compares the B200 kernels against the reference's kernels run on the same device.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ARFLOW_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def build(verbose=False):
    src = os.path.join(REF, "models", "correlation_package")
    if not os.path.isdir(src):
        return None
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    shim = os.path.join(HERE, "ref_compat.h")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    load(name="correlation_cuda", sources=[os.path.join(src, "correlation_cuda.cc"),
                                           os.path.join(src, "correlation_cuda_kernel.cu")],
         extra_cflags=["-include", shim, "-O2", "-w"],
         extra_cuda_cflags=["-include", shim, "-O3", "-w", "-gencode", "arch=compute_100a,code=sm_100a"],
         build_directory=OUT, verbose=verbose, is_python_module=False)
    so = os.path.join(OUT, "correlation_cuda.so")
    return so if os.path.exists(so) else None


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
