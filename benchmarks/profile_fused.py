"""Runs a few launches of the fused-prologue whole-clip op at the DeVIS R50 T=6 encoder shape: the command ncu wraps.
    python benchmarks/profile_fused.py [--dtype fp32|bf16] [--iters 2]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devis_b200 import TemporalMSDeformAttnFusedFunction, clip_geometry, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="fp32")
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[a.dtype]
torch.manual_seed(0)
T, shapes_l, M, D, pc, pt = 6, synthetic.DEVIS_SHAPES, 8, 32, 4, 4
nl, wt = len(shapes_l), T - 1
S = sum(h * w for h, w in shapes_l)
geom = clip_geometry.ClipGeometry(shapes_l, T, clip_geometry.all_frames_table(T))
order = geom.tile_order("cuda")
ref = synthetic.pixel_reference_points(shapes_l, T, "cuda")
value = torch.randn(T, S, M, D, device="cuda", dtype=dtype).requires_grad_(True)
off_c = (2.0 * torch.randn(T, S, M, nl, pc, 2, device="cuda")).requires_grad_(True)
off_t = (2.0 * torch.randn(T, S, M, wt * nl, pt, 2, device="cuda")).requires_grad_(True)
lg_c = torch.randn(T, S, M, nl * pc, device="cuda").requires_grad_(True)
lg_t = torch.randn(T, S, M, wt * nl * pt, device="cuda").requires_grad_(True)
gout = torch.randn(T, S, M * D, device="cuda", dtype=dtype)
for _ in range(a.iters):
    TemporalMSDeformAttnFusedFunction.apply(value, ref, off_c, lg_c, off_t, lg_t, geom, order).backward(gout)
torch.cuda.synchronize()
