"""Build the in-tree native library (libaadp.so) with nvcc for sm_100a.

Invoked by __graft_entry__.build() and lazily by alignment_algos_b200._lib when the .so is
missing or older than its sources.  Cross-compiles without a GPU.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libaadp.so")
SOURCES = ["aadp_api.cu"]
DEPS = ["aadp_api.cu", "aadp_kernels.cuh", "aadp_packed.cuh", "aadp_general.cuh", "aadp_frec.cuh", "aadp_enum.cuh", "aadp_pruned.h", os.path.join("..", "..", "include", "aadp.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libaadp.so cannot be built")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: an alternative build of the same library for A/B measurements (loaded with AADP_LIB=...)."""
    if out is None and not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines]
    # the image's default host compiler links libstdc++ statically (dangling .so link);
    # use the system g++ so the library shares libstdc++ with the host process
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", out or LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return out or LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
