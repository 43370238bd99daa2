"""Parity against the reference's ORIGINAL CUDA op (Deformable-DETR MultiScaleDeformableAttention), compiled
unmodified from /root/reference by oracle/ref_cuda_build.py into oracle/_ref/ ("Outputs must match ... its
original CUDA op on identical synthetic inputs", BASELINE.json north_star).  Skipped if oracle/_ref was not built."""
import pytest
import torch

from conftest import load_golden, nmax

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_cuda_build
    mod = ref_cuda_build.load()
    if mod is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return mod


def _both(ref, value, shapes, lsi, loc, aw, gout):
    from devis_b200 import MultiScaleDeformableAttention as ours
    o_ref = ref.ms_deform_attn_forward(value, shapes, lsi, loc, aw, 64)
    g_ref = ref.ms_deform_attn_backward(value, shapes, lsi, loc, aw, gout, 64)
    o = ours.ms_deform_attn_forward(value, shapes, lsi, loc, aw, 64)
    g = ours.ms_deform_attn_backward(value, shapes, lsi, loc, aw, gout, 64)
    return (o_ref, *g_ref), (o, *g)


@pytest.mark.parametrize("name", ["op_testpy", "op_ragged", "op_d32", "op_d30", "op_d71", "op_border"])
def test_fp64_bitwise_close_to_reference_cuda(ref, name):
    g = load_golden(name)
    args = [torch.from_numpy(g[k]).cuda() for k in ("value", "shapes", "lsi", "loc", "aw", "gout")]
    want, got = _both(ref, *args)
    for w, x in zip(want, got):
        assert nmax(x.cpu().numpy(), w.cpu().numpy()) < 1e-13


@pytest.mark.parametrize("kind,dist", [("enc_curr", "local"), ("enc_temporal", "local"), ("enc_temporal", "uniform"),
                                       ("dec_temporal", "uniform")])
def test_fp32_devis_call_shapes_match_reference_cuda(ref, kind, dist):
    from test_msda_gpu import _devis_call
    args = _devis_call(kind, dist, seed=21)
    want, got = _both(ref, *args)
    tol = (1e-5, 1e-4, 1e-4, 1e-4)
    for w, x, t in zip(want, got, tol):
        assert nmax(x.cpu().numpy(), w.cpu().numpy()) < t


def test_whole_clip_op_matches_reference_cuda_call_sequence(ref):
    """the reference's per-frame loop (ms_deform_attn.py:435-460) executed with ITS OWN CUDA op vs our single launch"""
    from devis_b200 import clip_geometry, synthetic, temporal_ms_deform_attn
    clip = synthetic.make_clip(dist="local", seed=5, device="cuda")
    geom = clip_geometry.ClipGeometry(clip["shapes"], 6, clip["frame_table"])
    out = temporal_ms_deform_attn(clip["value"], clip["loc_curr"], clip["aw_curr"], clip["loc_temporal"],
                                  clip["aw_temporal"], geom, geom.tile_order("cuda"))
    shapes = torch.tensor(clip["shapes"], device="cuda")
    areas = shapes.prod(1)
    lsi = torch.cat([areas.new_zeros(1), areas.cumsum(0)[:-1]])
    tshapes = shapes.repeat(5, 1)
    tareas = tshapes.prod(1)
    tlsi = torch.cat([tareas.new_zeros(1), tareas.cumsum(0)[:-1]])
    frames = []
    for t in range(6):
        cur = ref.ms_deform_attn_forward(clip["value"][t][None].contiguous(), shapes, lsi, clip["loc_curr"][t][None].contiguous(),
                                         clip["aw_curr"][t][None].contiguous(), 64)
        stacked = clip["value"][clip["frame_table"][t]].flatten(0, 1)[None].contiguous()
        tmp = ref.ms_deform_attn_forward(stacked, tshapes, tlsi, clip["loc_temporal"][t][None].contiguous(),
                                         clip["aw_temporal"][t][None].contiguous(), 64)
        frames.append(cur + tmp)
    want = torch.cat(frames, 0)
    assert nmax(out.cpu().numpy(), want.cpu().numpy()) < 1e-5
