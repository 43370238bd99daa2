"""A/B timing of library builds (build/variants/*.so) at the DeVIS layer-clip shape; each build runs in its own process.
    python benchmarks/variant_sweep.py [--kind fwd|bwd] [--dtype fp32]"""
import argparse
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from benchmarks.sweep import RawClip, time_us
from devis_b200 import synthetic, clip_geometry, _lib
kind, dtype, dist = sys.argv[1], sys.argv[2], sys.argv[3]
clip = synthetic.make_clip(device="cuda", dtype={"fp32": torch.float32, "bf16": torch.bfloat16}[dtype], dist=dist)
geom = clip_geometry.ClipGeometry(clip["shapes"], 6, clip["frame_table"])
rc = RawClip(clip, geom.tile_order("cuda"))
out = []
for threads in (128, 256):
    _lib.set_tuning(0 if kind == "fwd" else 2, threads)
    out.append("%%d:%%.1f" %% (threads, time_us(rc.fwd if kind == "fwd" else rc.bwd, 20)))
print(" ".join(out))
''' % ROOT

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="fwd")
ap.add_argument("--dtype", default="fp32")
ap.add_argument("--dist", default="local")
a = ap.parse_args()
libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so")))
for lib in [os.path.join(ROOT, "devis_b200", "libdevis_msda.so")] + libs:
    env = dict(os.environ, DEVIS_MSDA_LIB=lib, DEVIS_MSDA_TUNING="1")
    r = subprocess.run([sys.executable, "-c", CHILD, a.kind, a.dtype, a.dist], env=env, capture_output=True, text=True)
    print(f"{os.path.basename(lib):28s} {a.kind} {a.dtype} {a.dist}  us by threads: {r.stdout.strip()} {r.stderr.strip()[-200:] if r.returncode else ''}", flush=True)
