"""Seeded shape fuzzing of the drop-in op against the C oracle (fp32) and the PyTorch oracle (fp64): random level
pyramids, batch sizes, head / channel counts (grouped-lane kernels for D = 16, 32; generic kernels otherwise), point
counts, query counts, with taps partly outside the maps.  Everything goes through the C ABI."""
import numpy as np
import pytest
import torch

from conftest import nmax

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(seed)
    nl = int(rng.integers(1, 6))
    shapes = [(int(rng.integers(1, 14)), int(rng.integers(1, 18))) for _ in range(nl)]
    n = int(rng.integers(1, 4))
    m = int(rng.integers(1, 9))
    d = int(rng.choice([32, 32, 16, 8, 24, 33, 64]))
    lq = int(rng.integers(1, 70))
    p = int(rng.integers(1, 6))
    return shapes, n, m, d, lq, p


def _inputs(seed, dtype):
    from devis_b200 import synthetic
    shapes, n, m, d, lq, p = _case(seed)
    g = torch.Generator(device="cuda").manual_seed(seed)
    s = sum(h * w for h, w in shapes)
    sizes = torch.tensor([[w, h] for h, w in shapes], device="cuda", dtype=torch.float32)
    loc = torch.rand(n, lq, m, len(shapes), p, 2, generator=g, device="cuda") * 1.5 - 0.25     # ~1/3 of taps off-map
    loc = synthetic.make_boundary_safe(loc, sizes).to(dtype)
    aw = torch.softmax(torch.randn(n, lq, m, len(shapes) * p, generator=g, device="cuda"), -1).view(n, lq, m, len(shapes), p).to(dtype)
    value = torch.randn(n, s, m, d, generator=g, device="cuda").to(dtype)
    gout = torch.randn(n, lq, m * d, generator=g, device="cuda").to(dtype)
    shp = torch.tensor(shapes, device="cuda")
    areas = shp[:, 0] * shp[:, 1]
    lsi = torch.cat([areas.new_zeros(1), areas.cumsum(0)[:-1]])
    return value, shp, lsi, loc.contiguous(), aw.contiguous(), gout


def _run(value, shp, lsi, loc, aw, gout):
    from devis_b200 import MSDeformAttnFunction
    v, l_, a = (t.clone().requires_grad_(True) for t in (value, loc, aw))
    out = MSDeformAttnFunction.apply(v, shp, lsi, l_, a, 64)
    out.backward(gout)
    return out.detach(), v.grad, l_.grad, a.grad


@pytest.mark.parametrize("seed", list(range(24)))
def test_fuzz_fp32_vs_c_oracle(seed):
    from oracle import c_oracle
    args = _inputs(seed, torch.float32)
    got = _run(*args)
    value, shp, lsi, loc, aw, gout = (a.cpu().numpy() for a in args)
    want = (c_oracle.forward(value, shp, lsi, loc, aw),) + c_oracle.backward(value, shp, lsi, loc, aw, gout)
    for g, w, tol in zip(got, want, (1e-5, 1e-4, 1e-4, 1e-4)):
        assert g.shape == w.shape
        if np.abs(w).max() > 0:
            assert nmax(g.cpu().numpy(), w) < tol, (_case(seed),)


@pytest.mark.parametrize("seed", list(range(100, 108)))
def test_fuzz_fp64_vs_pytorch_oracle(seed):
    from oracle import msda_torch
    args = _inputs(seed, torch.float64)
    got = _run(*args)
    value, shp, lsi, loc, aw, gout = (a.cpu() for a in args)
    want = msda_torch.msda_forward_backward_torch(value, shp, loc, aw, gout)
    for g, w in zip(got, want):
        if w.abs().max() > 0:
            assert nmax(g.cpu().numpy(), w.numpy()) < 1e-11, (_case(seed),)


@pytest.mark.parametrize("seed", list(range(200, 206)))
def test_fuzz_whole_clip_vs_per_call(seed):
    """random clip geometries: the whole-clip op == sum of drop-in calls over (frame, slot) pairs"""
    from devis_b200 import MSDeformAttnFunction, clip_geometry, synthetic, temporal_ms_deform_attn
    rng = np.random.default_rng(seed)
    t = int(rng.integers(2, 6))
    shapes = tuple((int(rng.integers(2, 10)), int(rng.integers(2, 12))) for _ in range(int(rng.integers(1, 4))))
    wt = int(rng.integers(1, t))
    clip = synthetic.make_clip(n_frames=t, shapes=shapes, heads=int(rng.choice([2, 8])), channels=int(rng.choice([32, 16, 8])),
                               pc=int(rng.integers(1, 5)), pt=int(rng.integers(1, 5)), queries=int(rng.integers(1, 40)),
                               dist="uniform", seed=seed, device="cuda", t_window=wt)
    table = [[int(rng.integers(0, t)) for _ in range(wt)] for _ in range(t)]       # arbitrary, duplicates allowed
    geom = clip_geometry.ClipGeometry(shapes, t, table)
    out = temporal_ms_deform_attn(clip["value"], clip["loc_curr"], clip["aw_curr"], clip["loc_temporal"],
                                  clip["aw_temporal"], geom)
    shp = torch.tensor(shapes, device="cuda")
    areas = shp[:, 0] * shp[:, 1]
    lsi = torch.cat([areas.new_zeros(1), areas.cumsum(0)[:-1]])
    nl = len(shapes)
    for f in range(t):
        want = MSDeformAttnFunction.apply(clip["value"][f][None], shp, lsi, clip["loc_curr"][f][None].contiguous(),
                                          clip["aw_curr"][f][None].contiguous(), 64)
        for j, src in enumerate(table[f]):
            want = want + MSDeformAttnFunction.apply(
                clip["value"][src][None], shp, lsi, clip["loc_temporal"][f][None, :, :, j * nl:(j + 1) * nl].contiguous(),
                clip["aw_temporal"][f][None, :, :, j * nl:(j + 1) * nl].contiguous(), 64)
        assert nmax(out[f].cpu().numpy(), want[0].cpu().numpy()) < 2e-6


@pytest.mark.parametrize("m,p,seed", [(8, 4, 1), (8, 8, 2), (3, 4, 3), (1, 8, 4), (5, 4, 5), (8, 4, 6)])
def test_dead_corner_forward_path_with_taps_leaving_tiny_maps(m, p, seed):
    """The round-2 forward (msda_fwdv_kernel: D = 32, P % 4 == 0) predicates dead corners off and addresses the other three
    from a VIRTUAL top-left cell.  Maps of one row / one column / one pixel with a third of the taps outside exercise every
    combination of dead rows and columns, for 8 heads (row size as an immediate) and other head counts (run-time row size);
    forward and all three gradients against the C oracle, bf16 value against the fp32 result."""
    from devis_b200 import _lib, synthetic
    from oracle import c_oracle
    rng = np.random.default_rng(seed)
    shapes = [(1, 1), (1, int(rng.integers(2, 9))), (int(rng.integers(2, 9)), 1), (2, 2), (int(rng.integers(3, 12)), int(rng.integers(3, 12)))]
    n, lq, d = 2, 37, 32
    g = torch.Generator(device="cuda").manual_seed(seed)
    s = sum(h * w for h, w in shapes)
    sizes = torch.tensor([[w, h] for h, w in shapes], device="cuda", dtype=torch.float32)
    loc = torch.rand(n, lq, m, len(shapes), p, 2, generator=g, device="cuda") * 1.6 - 0.3
    loc = synthetic.make_boundary_safe(loc, sizes).contiguous()
    aw = torch.softmax(torch.randn(n, lq, m, len(shapes) * p, generator=g, device="cuda"), -1).view(n, lq, m, len(shapes), p).contiguous()
    value = torch.randn(n, s, m, d, generator=g, device="cuda")
    gout = torch.randn(n, lq, m * d, generator=g, device="cuda")
    shp = torch.tensor(shapes, device="cuda")
    areas = shp[:, 0] * shp[:, 1]
    lsi = torch.cat([areas.new_zeros(1), areas.cumsum(0)[:-1]])
    before = _lib.kernel_launches(_lib.KERNEL_FWD_GROUPED)
    got = _run(value, shp, lsi, loc, aw, gout)
    assert _lib.kernel_launches(_lib.KERNEL_FWD_GROUPED) == before + 1      # the grouped-lane forward, not the generic one
    args = [a.cpu().numpy() for a in (value, shp, lsi, loc, aw, gout)]
    want = (c_oracle.forward(*args[:5]),) + c_oracle.backward(*args)
    for a, w, tol in zip(got, want, (1e-5, 1e-4, 1e-4, 1e-4)):
        assert nmax(a.cpu().numpy(), w) < tol
    # bf16 value through the same kernel template (64-byte rows, row size 512 B as the immediate)
    from devis_b200 import MSDeformAttnFunction
    out_bf16 = MSDeformAttnFunction.apply(value.bfloat16(), shp, lsi, loc, aw, 64)
    assert nmax(out_bf16.float().cpu().numpy(), got[0].cpu().numpy()) < 1e-2


def test_out_of_range_taps_do_not_read_the_map():
    """Reference semantics (cuh:285-288, 56-78): a tap outside the map, or a corner outside it, reads nothing.  Round 1
    read a clamped in-map pixel with a zero factor, so a NaN / Inf there leaked into queries that never sample it; the
    round-2 forward does not load dead corners at all."""
    from devis_b200 import MSDeformAttnFunction
    m, d, p = 8, 32, 4
    shp = torch.tensor([[4, 4]], device="cuda")
    lsi = torch.zeros(1, dtype=torch.long, device="cuda")
    value = torch.randn(1, 16, m, d, device="cuda")
    loc = torch.full((1, 5, m, 1, p, 2), 0.9, device="cuda")          # cell (3, 3): right column and bottom row are dead
    loc[..., 0, :] = 1.5                                               # first point of every (query, head): out of range
    aw = torch.full((1, 5, m, 1, p), 0.25, device="cuda")
    clean = MSDeformAttnFunction.apply(value, shp, lsi, loc, aw, 64)
    poisoned = value.clone()
    poisoned[0, 0] = float("nan")                                      # pixel (0, 0): no live corner touches it
    poisoned[0, 1] = float("inf")
    out = MSDeformAttnFunction.apply(poisoned, shp, lsi, loc, aw, 64)
    assert torch.isfinite(out).all()
    assert torch.equal(out, clean)


@pytest.mark.parametrize("rows_h", [1900, 2100])
def test_forward_addresses_near_the_two_gib_boundary(rows_h):
    """The round-2 forward keeps SIGNED 32-bit byte offsets (a virtual top-left cell may lie before the map), so it serves
    value tensors below 2 GiB and hands larger ones to the unsigned-offset kernel (below 4 GiB).  1.95 GB and 2.15 GB of
    value, taps in the last rows of the map: the result must equal the same taps evaluated on a copy of just those rows."""
    from devis_b200 import MSDeformAttnFunction
    m, d, p, w, tail, lq = 8, 32, 4, 1000, 8, 64
    g = torch.Generator(device="cuda").manual_seed(rows_h)
    value = torch.empty(1, rows_h * w, m, d, device="cuda")
    value[:, -tail * w:] = torch.randn(1, tail * w, m, d, device="cuda", generator=g)
    shp_big = torch.tensor([[rows_h, w]], device="cuda")
    shp_small = torch.tensor([[tail, w]], device="cuda")
    lsi = torch.zeros(1, dtype=torch.long, device="cuda")
    # pixel coordinates inside the last `tail` rows, at least one row away from the slice's upper edge
    py = rows_h - tail + 1.25 + torch.rand(1, lq, m, 1, p, device="cuda", generator=g) * (tail - 2.5)
    px = torch.rand(1, lq, m, 1, p, device="cuda", generator=g) * (w - 1.0) + 0.25
    loc_big = torch.stack([(px + 0.5) / w, (py + 0.5) / rows_h], -1).contiguous()
    loc_small = torch.stack([(px + 0.5) / w, (py - (rows_h - tail) + 0.5) / tail], -1).contiguous()
    aw = torch.softmax(torch.randn(1, lq, m, p, device="cuda", generator=g), -1).view(1, lq, m, 1, p).contiguous()
    out_big = MSDeformAttnFunction.apply(value, shp_big, lsi, loc_big, aw, 64)
    out_small = MSDeformAttnFunction.apply(value[:, -tail * w:].contiguous(), shp_small, lsi, loc_small, aw, 64)
    # the two calls round y differently (y * H in float32 at H = 1900 vs 8): compare at that resolution
    assert nmax(out_big.cpu().numpy(), out_small.cpu().numpy()) < 2e-3
    assert torch.isfinite(out_big).all()
