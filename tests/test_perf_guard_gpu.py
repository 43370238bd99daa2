"""Coarse performance guards at the DeVIS R50 T=6 encoder layer-clip (the bench.py workload): a kernel that silently
falls off its launch shape or register budget (it happened: a launch-bounds hint made the bf16 forward 2x slower with
every parity test green) fails here.  Bounds are 50 % above the times measured on B200 (profiles/) -- room for a slower-clocked box, still
below a 2x regression; medians of 10 launches after warm-up, CUDA events."""
import statistics

import pytest
import torch

pytestmark = pytest.mark.gpu


def _median_us(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev) * 1e3


@pytest.mark.parametrize("dtype,fwd_bound,bwd_bound", [(torch.float32, 720.0, 2130.0), (torch.bfloat16, 540.0, 2080.0)])
def test_whole_clip_kernels_stay_near_their_measured_times(dtype, fwd_bound, bwd_bound):
    if torch.cuda.get_device_properties(0).multi_processor_count < 140:
        pytest.skip("bounds are B200 numbers")
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from benchmarks.sweep import RawClip
    from devis_b200 import clip_geometry, synthetic
    clip = synthetic.make_clip(device="cuda", dtype=dtype, dist="local")
    geom = clip_geometry.ClipGeometry(clip["shapes"], 6, clip["frame_table"])
    rc = RawClip(clip, geom.tile_order("cuda"))
    fwd, bwd = _median_us(rc.fwd), _median_us(rc.bwd)
    assert fwd < fwd_bound, f"forward {fwd:.0f} us (measured on B200: 478 fp32 / 359 bf16)"
    assert bwd < bwd_bound, f"backward {bwd:.0f} us (measured on B200: 1421 fp32 / 1387 bf16)"
