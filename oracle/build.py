"""oracle/build.py -- TEST INFRASTRUCTURE.  Build recipes for the checker side.

  python -m oracle.build            # C oracle (gcc), and oracle/_ref if /root/reference exists
  python -m oracle.build --c-only
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
C_SRC = os.path.join(HERE, "msda_oracle.c")
C_LIB = os.path.join(HERE, "libmsda_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_OPS = "/root/reference/src/models/ops/src"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_c_oracle(force=False):
    """gcc -O2 -ffp-contract=off: no FMA contraction, so fp32 results are the plain
    IEEE sequence the restatement spells out."""
    if force or _stale(C_LIB, [C_SRC]):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", C_LIB, C_SRC, "-lm"]
        subprocess.check_call(cmd)
    return C_LIB


def build_reference_cuda_op(force=False):
    """Compile the reference's own CUDA op FROM WHERE IT LIES under /root/reference into
    oracle/_ref/ (git-ignored, travels to the GPU box).  See oracle/ref_shim.h for the one
    compatibility overload torch 2.11 needs; no reference source is copied or edited."""
    from . import ref_cuda_build
    return ref_cuda_build.build(force=force)


def main(argv):
    build_c_oracle(force="--force" in argv)
    print("built", C_LIB)
    if "--c-only" not in argv and os.path.isdir(REFERENCE_OPS):
        try:
            print("built", build_reference_cuda_op(force="--force" in argv))
        except Exception as exc:  # the reference build is optional evidence, never required
            print("reference CUDA op not built:", exc)


if __name__ == "__main__":
    main(sys.argv[1:])
