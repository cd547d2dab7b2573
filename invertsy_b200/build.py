"""
Builds invertsy_b200/libisy_b200.so (the C-ABI CUDA library of include/isy_b200.h) in-tree with
nvcc for sm_100a.  The library links only the CUDA runtime: no torch types cross the boundary.

    python -m invertsy_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libisy_b200.so")
SOURCES = ["isy_capi.cu"]
DEPS = ["isy_capi.cu", "isy_kernels.cuh", "isy_common.cuh", "isy_layout.cuh", "isy_setup.cuh", "isy_project.cuh",
        "isy_bins.cuh", "isy_raster.cuh", "isy_shade.cuh", "isy_render.cuh", "isy_sky_kernels.cuh", "isy_fam.cuh",
        "isy_willshaw.cuh",
        os.path.join("..", "..", "include", "isy_b200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not (force or needs_build()):
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
