"""A/B timing of library builds (the in-tree library + build/variants/*.so, see build_variants.py) at the DeVIS layer-clip
shape; each build runs in its own process (DEVIS_MSDA_LIB selects the library).  Per build: median per-launch time of the
whole-clip forward and backward (events around each launch), fp32 and optionally bf16.
    python benchmarks/variant_sweep.py [--dtype fp32] [--dist local] [--iters 30] [--out gpurun_out/variants.json]"""
import argparse
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, statistics, sys, torch
sys.path.insert(0, %r)
from benchmarks.sweep import RawClip
from devis_b200 import synthetic, clip_geometry
dtype, dist, iters = sys.argv[1], sys.argv[2], int(sys.argv[3])
clip = synthetic.make_clip(device="cuda", dtype={"fp32": torch.float32, "bf16": torch.bfloat16}[dtype], dist=dist)
geom = clip_geometry.ClipGeometry(clip["shapes"], 6, clip["frame_table"])
rc = RawClip(clip, geom.tile_order("cuda"))
def med(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev) * 1e3
print(json.dumps({"fwd_us": round(med(rc.fwd), 1), "bwd_us": round(med(rc.bwd), 1)}))
''' % ROOT


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--dist", default="local")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    libs = [os.path.join(ROOT, "devis_b200", "libdevis_msda.so")] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so")))
    rows = {}
    for rep in range(2):                      # two passes over the builds: run-to-run drift shows up as a difference
        for lib in libs:
            env = dict(os.environ, DEVIS_MSDA_LIB=lib)
            r = subprocess.run([sys.executable, "-c", CHILD, a.dtype, a.dist, str(a.iters)], env=env, capture_output=True, text=True)
            name = os.path.basename(lib)
            try:
                rows.setdefault(name, []).append(json.loads(r.stdout.strip().splitlines()[-1]))
            except Exception:  # noqa: BLE001
                rows.setdefault(name, []).append({"error": r.stderr.strip()[-300:]})
            print(name, a.dtype, a.dist, rows[name][-1], flush=True)
    if a.out:
        with open(a.out, "w") as fh:
            json.dump({"dtype": a.dtype, "dist": a.dist, "iters": a.iters, "rows": rows}, fh, indent=1)


if __name__ == "__main__":
    main()
