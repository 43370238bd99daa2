"""Runs the tensor-core deformable-conv forward at a mask-head layer shape: the command ncu wraps.
    python benchmarks/profile_dcn_igemm.py [--layer lay1|lay3|lay4]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devis_b200.deform_conv import deform_conv2d  # noqa: E402

LAYERS = {"lay1": (264, 264, 12, 20), "lay3": (136, 64, 23, 40), "lay4": (72, 32, 45, 80)}
ap = argparse.ArgumentParser()
ap.add_argument("--layer", default="lay1")
a = ap.parse_args()
c, cout, h, w = LAYERS[a.layer]
n = 60
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, c, h, w, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
off = 1.5 * torch.randn(n, 18, h, w, device="cuda", generator=g)
msk = torch.rand(n, 9, h, w, device="cuda", generator=g)
wt = torch.randn(cout, c, 3, 3, device="cuda", generator=g) / (9 * c) ** 0.5
b = torch.randn(cout, device="cuda", generator=g)
with torch.no_grad():
    for _ in range(3):
        deform_conv2d(x, off, wt, b, padding=1, mask=msk)
torch.cuda.synchronize()
