"""Builds alternative libraries for kernel A/B runs: build/variants/lib_<name>.so, each the product sources compiled with
extra -D flags (the compile-time switches of devis_b200/csrc: DEVIS_HINTS, DEVIS_L2_HINTS, DEVIS_BWD_MIN_BLOCKS, ...).
    python benchmarks/build_variants.py l2_0:-DDEVIS_L2_HINTS=0 l2_1:-DDEVIS_L2_HINTS=1 ...
variant_sweep.py then times every build in its own process (DEVIS_MSDA_LIB selects the library)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from devis_b200 import build as product_build  # noqa: E402


def build_one(spec):
    name, _, flags = spec.partition(":")
    out = os.path.join(ROOT, "build", "variants", f"lib_{name}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [product_build.nvcc_path()] + product_build.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-o", out] + product_build.SOURCES
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    with ThreadPoolExecutor(max_workers=4) as pool:
        for path in pool.map(build_one, sys.argv[1:]):
            print(path, flush=True)
