"""In-tree build of libmpmb.so (hand-written sm_100a CUDA + the C-ABI of include/mpmb.h).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels with gpurun snapshots.
"""
import os
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SRC = [os.path.join(_PKG, "csrc", "mpmb_engine.cu")]
DEPS = SRC + [os.path.join(_PKG, "csrc", "mpmb_math.cuh"), os.path.join(_ROOT, "include", "mpmb.h")]
LIB = os.path.join(_PKG, "lib", "libmpmb.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    """Compile libmpmb.so for sm_100a if it is missing or stale.  Returns the path."""
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SRC
    env = dict(os.environ)
    # this image exports CC/CXX pointing at a wrapper without OpenMP specs; nvcc should use the system g++
    env.pop("CC", None)
    env.pop("CXX", None)
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building libmpmb.so")
    if verbose:
        sys.stderr.write(r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
