import os, sys, torch
sys.path.insert(0, os.getcwd())
from benchmarks.sweep import RawClip
from devis_b200 import _lib, clip_geometry, synthetic
clip = synthetic.make_clip(dist="local", device="cuda")
geom = clip_geometry.ClipGeometry(clip["shapes"], 6, clip["frame_table"])
rc = RawClip(clip, geom.tile_order("cuda", 8, 8))
_lib.set_tuning(6, 2); _lib.set_tuning(7, int(sys.argv[1])); _lib.set_tuning(8, int(sys.argv[2]))
for _ in range(2):
    rc.bwd()
torch.cuda.synchronize()
