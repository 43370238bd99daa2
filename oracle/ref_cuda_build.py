"""oracle/ref_cuda_build.py -- TEST INFRASTRUCTURE.

Compiles the reference's ORIGINAL CUDA op (Deformable-DETR MultiScaleDeformableAttention) from the sources where
they lie under /root/reference/src/models/ops/src into oracle/_ref/ (git-ignored, shipped to the GPU box), for
sm_100a.  No reference source is copied or edited; the only addition is the force-included oracle/ref_shim.h.
It is used by tests and benchmarks as a second oracle ("its original CUDA op", BASELINE.json north_star) and as the
GPU baseline -- never by the product path.

    python -m oracle.ref_cuda_build         # build (about a minute)
    oracle.ref_cuda_build.load()            # import the built module on the GPU box (no sources needed)
"""
import glob
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SRC = "/root/reference/src/models/ops/src"
NAME = "MultiScaleDeformableAttention_ref"
SHIM = os.path.join(HERE, "ref_shim.h")


def built_path():
    hits = glob.glob(os.path.join(REF_DIR, NAME + "*.so"))
    return hits[0] if hits else None


def build(force=False):
    if built_path() and not force:
        return built_path()
    if not os.path.isdir(SRC):
        raise RuntimeError("reference sources are not mounted; oracle/_ref can only be built in the build container")
    os.makedirs(REF_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils import cpp_extension
    sources = [os.path.join(SRC, "vision.cpp"), os.path.join(SRC, "cpu", "ms_deform_attn_cpu.cpp"),
               os.path.join(SRC, "cuda", "ms_deform_attn_cuda.cu")]
    cpp_extension.load(
        name=NAME, sources=sources, extra_include_paths=[SRC], build_directory=REF_DIR, is_python_module=False,
        extra_cflags=["-DWITH_CUDA", "-include", SHIM, "-w"],
        # the reference's own nvcc defines (setup.py:41-44) + an explicit Blackwell target (it passes no -gencode)
        extra_cuda_cflags=["-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                           "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__", "-include", SHIM,
                           "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-w"],
        verbose=False)
    return built_path()


def load():
    """Import the prebuilt module (GPU box: the .so travelled with the snapshot).  Returns None if it was never built."""
    path = built_path()
    if path is None:
        return None
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
