"""devis_b200/build.py -- builds the in-tree C-ABI library with nvcc for sm_100a.

    python -m devis_b200.build [--force] [--verbose]

The output, devis_b200/libdevis_msda.so, is git-ignored but travels to the GPU box with the
repository snapshot.  Replaces the reference's src/models/ops/setup.py + make.sh (a torch
CUDAExtension without arch flags); this library has no torch / pybind dependency at all.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdevis_msda.so")
SOURCES = [os.path.join(CSRC, "msda_capi.cu"), os.path.join(CSRC, "deform_conv_capi.cu")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--expt-relaxed-constexpr",
]


def _deps():
    deps = [os.path.join(HERE, "..", "include", h) for h in ("devis_msda.h", "devis_deform_conv.h")]
    for name in os.listdir(CSRC):
        if name.endswith((".cu", ".cuh", ".h")):
            deps.append(os.path.join(CSRC, name))
    return deps


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
