"""Developer tool: deterministic whole-clip backward at the DeVIS R50 T=6 encoder shape -- the direct 64-bit fixed-point
scatter (tuning key 6 = 1) against the sorted pre-aggregation (msda_bwds_kernel<DET>, the default in the encoder form) and
the default float-atomic backward; checks that the two deterministic paths agree bit for bit.

    python benchmarks/det_bench.py [--dtype fp32|bf16] [--dist local|uniform] [--iters 20] [--out file.json]
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
    ap.add_argument("--margins", default="6")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[a.dtype]
    clip = synthetic.make_clip(dist=a.dist, dtype=dtype, device="cuda")
    geom = clip_geometry.ClipGeometry(clip["shapes"], 6, clip["frame_table"])
    rc = RawClip(clip, geom.tile_order("cuda", 8, 8))
    det = _lib.FLAG_DETERMINISTIC
    res = {"dtype": a.dtype, "dist": a.dist}
    _lib.set_tuning(6, 1)
    res["float_atomics_us"] = round(time_us(rc.bwd, a.iters), 1)
    rc.bwd(det)
    torch.cuda.synchronize()
    ref = rc.gv.clone()
    res["deterministic_direct_scatter_us"] = round(time_us(lambda: rc.bwd(det), a.iters), 1)
    _lib.set_tuning(6, 0)
    for margin in [int(x) for x in a.margins.split(",")]:
        _lib.set_tuning(7, margin)
        rc.gv.fill_(float("nan"))
        rc.bwd(det)
        torch.cuda.synchronize()
        res[f"deterministic_sorted_margin{margin}_us"] = round(time_us(lambda: rc.bwd(det), a.iters), 1)
        res[f"deterministic_sorted_margin{margin}_bit_identical"] = bool(torch.equal(rc.gv, ref))
    _lib.set_tuning(7, 0)
    print(json.dumps(res), flush=True)
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    os.environ["DEVIS_MSDA_TUNING"] = "1"      # developer knobs (devis_msda_set_tuning) are inert without it
    main()
