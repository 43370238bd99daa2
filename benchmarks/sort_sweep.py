"""Developer tool: the sorted whole-clip backward (msda_bwd_sort.cuh) against the direct-scatter kernel at the DeVIS
R50 T=6 encoder shape -- time, and the difference of grad_value (normalised max / rms) -- for window margins and the
first level that gets a window.

    python benchmarks/sort_sweep.py [--dtype fp32|bf16] [--dist local|uniform] [--iters 20] [--out file.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devis_b200 import _lib, clip_geometry, synthetic  # noqa: E402
from benchmarks.sweep import RawClip, time_us  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--dist", default="local")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="")
    ap.add_argument("--margins", default="3,6,9")
    ap.add_argument("--min-levels", default="0,1,2")
    a = ap.parse_args()
    dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[a.dtype]
    clip = synthetic.make_clip(dist=a.dist, dtype=dtype, device="cuda")
    geom = clip_geometry.ClipGeometry(clip["shapes"], 6, clip["frame_table"])
    rc = RawClip(clip, geom.tile_order("cuda", 8, 8))
    rows = []
    _lib.set_tuning(6, 1)                       # direct scatter
    rc.bwd()
    torch.cuda.synchronize()
    ref = [t.clone() for t in (rc.gv, rc.glc, rc.gac, rc.glt, rc.gat)]
    us = time_us(rc.bwd, a.iters)
    rows.append(dict(kind="direct", us=round(us, 1)))
    print(rows[-1], flush=True)
    _lib.set_tuning(6, 3)
    for min_level in [int(x) for x in a.min_levels.split(",")]:
        for margin in [int(x) for x in a.margins.split(",")]:
            _lib.set_tuning(7, margin)
            _lib.set_tuning(9, min_level + 1)
            rc.gv.fill_(float("nan"))
            rc.bwd()
            torch.cuda.synchronize()
            d = (rc.gv.double() - ref[0].double())
            scale = ref[0].double().abs().max()
            same = all(torch.equal(x, y) for x, y in zip((rc.glc, rc.gac, rc.glt, rc.gat), ref[1:]))
            us = time_us(rc.bwd, a.iters)
            rows.append(dict(kind="sorted", margin=margin, min_level=min_level, us=round(us, 1),
                             gv_max_err=float(d.abs().max() / scale), gv_rms_err=float(d.pow(2).mean().sqrt() / scale),
                             other_grads_bit_identical=same))
            print(rows[-1], flush=True)
    _lib.set_tuning(6, 0); _lib.set_tuning(7, 0); _lib.set_tuning(9, 0)
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()
