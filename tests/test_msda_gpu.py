"""GPU parity of the drop-in op (MSDeformAttnFunction / MultiScaleDeformableAttention mirror) against
the golden fixtures produced by the reference's own PyTorch path, and against the CPU oracle.
Everything here goes through the C ABI (devis_b200/_lib.py -> libdevis_msda.so).

Tolerances (BASELINE.json north_star, normalised max error on boundary-safe taps):
forward 1e-5 fp32 / 1e-2 bf16; grad_value, grad_sampling_loc, grad_attn_weight 1e-4 fp32.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, nmax

pytestmark = pytest.mark.gpu

OP_CASES = ["op_testpy", "op_ragged", "op_d32", "op_d1", "op_d30", "op_d71", "op_border"]


def _cuda(g, dtype):
    dev = "cuda"
    f = lambda k: torch.from_numpy(g[k]).to(dev, dtype)
    i = lambda k: torch.from_numpy(g[k]).to(dev)
    return f("value"), i("shapes"), i("lsi"), f("loc"), f("aw"), f("gout")


def _run(value, shapes, lsi, loc, aw, gout, step=64):
    from devis_b200 import MSDeformAttnFunction
    v, l_, a = (t.clone().requires_grad_(True) for t in (value, loc, aw))
    out = MSDeformAttnFunction.apply(v, shapes, lsi, l_, a, step)
    out.backward(gout)
    return out.detach(), v.grad, l_.grad, a.grad


@pytest.mark.parametrize("name", OP_CASES)
def test_fp64_matches_reference_pytorch_core(name):
    """the reference's own check_forward_equal_with_pytorch_double (test.py:29-42), plus the gradients"""
    g = load_golden(name)
    out, gv, gl, ga = _run(*_cuda(g, torch.float64))
    assert torch.allclose(out.cpu(), torch.from_numpy(g["out"]))          # test.py:38 defaults
    assert nmax(out.cpu().numpy(), g["out"]) < 1e-12
    assert nmax(gv.cpu().numpy(), g["gvalue"]) < 1e-12
    assert nmax(gl.cpu().numpy(), g["gloc"]) < 1e-11
    assert nmax(ga.cpu().numpy(), g["gaw"]) < 1e-12


@pytest.mark.parametrize("name", [n for n in OP_CASES if n != "op_testpy"])
def test_fp32_within_north_star_tolerance(name):
    g = load_golden(name)
    out, gv, gl, ga = _run(*_cuda(g, torch.float32))
    assert torch.allclose(out.cpu().double(), torch.from_numpy(g["out"]), rtol=1e-2, atol=1e-3)   # test.py:54
    assert nmax(out.cpu().numpy(), g["out"]) < 1e-5
    assert nmax(gv.cpu().numpy(), g["gvalue"]) < 1e-4
    assert nmax(gl.cpu().numpy(), g["gloc"]) < 1e-4
    assert nmax(ga.cpu().numpy(), g["gaw"]) < 1e-4


@pytest.mark.parametrize("name", ["op_d32", "op_ragged", "op_d30"])
def test_bf16_forward_and_grads(name):
    """extension dtype: value/out/grad_out bf16, locations and weights fp32"""
    g = load_golden(name)
    value, shapes, lsi, loc, aw, gout = _cuda(g, torch.float32)
    out, gv, gl, ga = _run(value.bfloat16(), shapes, lsi, loc, aw, gout.bfloat16())
    assert out.dtype == torch.bfloat16 and gv.dtype == torch.bfloat16 and gl.dtype == torch.float32
    assert nmax(out.float().cpu().numpy(), g["out"]) < 1e-2
    assert nmax(gv.float().cpu().numpy(), g["gvalue"]) < 2e-2
    assert nmax(gl.cpu().numpy(), g["gloc"]) < 2e-2
    assert nmax(ga.cpu().numpy(), g["gaw"]) < 2e-2


def _devis_call(kind, dist, seed):
    """the per-call shapes of SURVEY.md section 3.4 built from one synthetic clip"""
    from devis_b200 import synthetic
    lq = {"enc_curr": None, "enc_temporal": None, "dec_curr": 30, "dec_temporal": 300}[kind]
    clip = synthetic.make_clip(queries=lq, dist=dist, seed=seed, device="cuda")
    shapes = torch.tensor(clip["shapes"], device="cuda")
    t = 2
    if kind.endswith("curr"):
        value, loc, aw = clip["value"][t][None], clip["loc_curr"][t][None], clip["aw_curr"][t][None]
    else:
        frames = clip["frame_table"][t]
        value = clip["value"][frames].flatten(0, 1)[None]
        loc, aw = clip["loc_temporal"][t][None], clip["aw_temporal"][t][None]
        shapes = shapes.repeat(len(frames), 1)
    areas = shapes[:, 0] * shapes[:, 1]
    lsi = torch.cat([areas.new_zeros(1), areas.cumsum(0)[:-1]])
    return value.contiguous(), shapes, lsi, loc.contiguous(), aw.contiguous(), clip["grad_out"][t][None].contiguous()


@pytest.mark.parametrize("kind,dist", [("enc_curr", "local"), ("enc_curr", "uniform"), ("enc_temporal", "local"),
                                       ("dec_curr", "uniform"), ("dec_temporal", "local")])
def test_fp32_matches_c_oracle_at_devis_call_shapes(kind, dist):
    """CUDA fp32 vs the C restatement of the reference CUDA kernels (also fp32) on the same inputs, at the
    shapes the DeVIS R50 T=6 model really issues."""
    from oracle import c_oracle
    args = _devis_call(kind, dist, seed=11)
    out, gv, gl, ga = _run(*args)
    value, shapes, lsi, loc, aw, gout = (a.cpu().numpy() for a in args)
    ref_out = c_oracle.forward(value, shapes, lsi, loc, aw)
    ref_gv, ref_gl, ref_ga = c_oracle.backward(value, shapes, lsi, loc, aw, gout)
    assert nmax(out.cpu().numpy(), ref_out) < 1e-5
    assert nmax(gv.cpu().numpy(), ref_gv) < 1e-4
    assert nmax(gl.cpu().numpy(), ref_gl) < 1e-4
    assert nmax(ga.cpu().numpy(), ref_ga) < 1e-4


def test_batched_with_im2col_step_matches_per_item_calls():
    g = load_golden("op_ragged")       # batch 2
    value, shapes, lsi, loc, aw, gout = _cuda(g, torch.float32)
    whole = _run(value, shapes, lsi, loc, aw, gout, step=1)
    whole2 = _run(value, shapes, lsi, loc, aw, gout, step=2)
    for a, b in zip(whole, whole2):
        assert torch.equal(a, b) or nmax(a.cpu().numpy(), b.cpu().numpy()) < 1e-6
    for n in range(2):
        part = _run(value[n:n + 1].contiguous(), shapes, lsi, loc[n:n + 1].contiguous(), aw[n:n + 1].contiguous(),
                    gout[n:n + 1].contiguous())
        assert torch.equal(part[0], whole[0][n:n + 1])


def test_random_locations_like_reference_test_quantile_gate():
    """torch.rand locations (test.py:32) are not kept away from cell borders: gate grad_loc on p99.9"""
    from oracle import c_oracle
    torch.manual_seed(5)
    shapes = torch.tensor([(24, 32), (12, 16)], device="cuda")
    lsi = torch.tensor([0, 24 * 32], device="cuda")
    n, m, d, lq, nl, p = 1, 8, 32, 500, 2, 4
    value = torch.randn(n, 24 * 32 + 12 * 16, m, d, device="cuda")
    loc = torch.rand(n, lq, m, nl, p, 2, device="cuda")
    aw = torch.softmax(torch.randn(n, lq, m, nl * p, device="cuda"), -1).view(n, lq, m, nl, p)
    gout = torch.randn(n, lq, m * d, device="cuda")
    out, gv, gl, ga = _run(value, shapes, lsi, loc, aw, gout)
    r = [t.cpu().numpy() for t in (value, shapes, lsi, loc, aw, gout)]
    ref_out = c_oracle.forward(*r[:5])
    ref_gv, ref_gl, ref_ga = c_oracle.backward(*r)
    assert nmax(out.cpu().numpy(), ref_out) < 1e-5
    assert nmax(gv.cpu().numpy(), ref_gv) < 1e-4
    err = np.abs(gl.cpu().numpy() - ref_gl) / np.abs(ref_gl).max()
    assert np.quantile(err, 0.999) < 1e-4


@pytest.mark.parametrize("channels", [30, 32, 64, 71, 1025, 2048, 3096])
def test_gradcheck_like_reference(channels):
    """test.py:61-76,83: torch.autograd.gradcheck of the Function in double for the reference's full channel sweep
    (30, 32, 64, 71, 1025, 2048, 3096 -- chosen there to hit every backward-kernel branch of cuh:978-1320; here
    grouped-lane for 32, generic otherwise)."""
    from torch.autograd import gradcheck
    from devis_b200 import MSDeformAttnFunction
    torch.manual_seed(3)
    n, m, lq, nl, p = 1, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long).cuda()
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    s = int(shapes.prod(1).sum())
    value = (torch.rand(n, s, m, channels).cuda() * 0.01).double().requires_grad_(True)
    loc = torch.rand(n, lq, m, nl, p, 2).cuda().double().requires_grad_(True)
    aw = torch.rand(n, lq, m, nl, p).cuda().double() + 1e-5
    aw = (aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)).requires_grad_(True)
    assert gradcheck(MSDeformAttnFunction.apply, (value, shapes, lsi, loc, aw, 2))


def test_error_behaviour_matches_reference():
    from devis_b200 import MSDeformAttnFunction, MultiScaleDeformableAttention as MSDA
    g = load_golden("op_ragged")
    value, shapes, lsi, loc, aw, gout = _cuda(g, torch.float32)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDA.ms_deform_attn_forward(value.cpu(), shapes.cpu(), lsi.cpu(), loc.cpu(), aw.cpu(), 64)
    with pytest.raises(RuntimeError, match="contiguous"):
        MSDA.ms_deform_attn_forward(value, shapes, lsi, loc.transpose(1, 2), aw, 64)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        MSDA.ms_deform_attn_forward(value, shapes.cpu(), lsi, loc, aw, 64)
    with pytest.raises(RuntimeError):
        MSDA.ms_deform_attn_forward(value.half(), shapes, lsi, loc.half(), aw.half(), 64)
    v3 = torch.cat([value, value[:1]], 0)
    with pytest.raises(RuntimeError, match="im2col_step"):
        MSDA.ms_deform_attn_forward(v3, shapes, lsi, torch.cat([loc, loc[:1]], 0), torch.cat([aw, aw[:1]], 0), 2)
    with pytest.raises(RuntimeError, match="contiguous"):
        MSDA.ms_deform_attn_backward(value, shapes, lsi, loc, aw, gout.transpose(0, 1).contiguous().transpose(0, 1), 64)
    out = MSDeformAttnFunction.apply(value, shapes, lsi, loc, aw, 64)
    assert out.shape == (2, 11, 24)


def test_empty_inputs():
    from devis_b200 import MultiScaleDeformableAttention as MSDA
    g = load_golden("op_ragged")
    value, shapes, lsi, loc, aw, gout = _cuda(g, torch.float32)
    out = MSDA.ms_deform_attn_forward(value, shapes, lsi, loc[:, :0].contiguous(), aw[:, :0].contiguous(), 64)
    assert out.shape == (2, 0, 24)
    gv, gl, ga = MSDA.ms_deform_attn_backward(value, shapes, lsi, loc[:, :0].contiguous(), aw[:, :0].contiguous(),
                                              gout[:, :0].contiguous(), 64)
    assert gv.shape == value.shape and not gv.any() and gl.numel() == 0 and ga.numel() == 0
    out0 = MSDA.ms_deform_attn_forward(value[:0], shapes, lsi, loc[:0], aw[:0], 64)
    assert out0.shape == (0, 11, 24)


def test_launches_are_counted_and_on_current_stream():
    from devis_b200 import _lib, MultiScaleDeformableAttention as MSDA
    g = load_golden("op_d32")
    value, shapes, lsi, loc, aw, gout = _cuda(g, torch.float32)
    before = _lib.launch_count()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        out = MSDA.ms_deform_attn_forward(value, shapes, lsi, loc, aw, 64)
    side.synchronize()
    assert _lib.launch_count() == before + 1
    assert nmax(out.cpu().numpy(), g["out"]) < 1e-5


def test_frozen_value_skips_grad_value():
    from devis_b200 import MSDeformAttnFunction
    g = load_golden("op_d32")
    value, shapes, lsi, loc, aw, gout = _cuda(g, torch.float32)
    l_, a = loc.clone().requires_grad_(True), aw.clone().requires_grad_(True)
    out = MSDeformAttnFunction.apply(value, shapes, lsi, l_, a, 64)       # value does not require grad
    out.backward(gout)
    assert nmax(l_.grad.cpu().numpy(), g["gloc"]) < 1e-4 and nmax(a.grad.cpu().numpy(), g["gaw"]) < 1e-4


def test_single_scale_long_clip_T36():
    """the reference's single-scale T=36 ablation (docs/TRAIN.md:41): 1 level, 35 temporal frames, 2 temporal points"""
    from devis_b200 import clip_geometry, synthetic, temporal_ms_deform_attn
    from oracle import temporal_torch
    shapes = ((12, 20),)
    clip = synthetic.make_clip(n_frames=36, shapes=shapes, queries=5, pt=2, dist="uniform", seed=9, device="cuda")
    geom = clip_geometry.ClipGeometry(shapes, 36, clip["frame_table"])
    out = temporal_ms_deform_attn(clip["value"], clip["loc_curr"], clip["aw_curr"], clip["loc_temporal"],
                                  clip["aw_temporal"], geom)
    cpu = lambda k: clip[k].double().cpu()
    offs = [torch.tensor([f - t for f in row]) for t, row in enumerate(clip["frame_table"])]
    ref = temporal_torch.temporal_core_per_frame(cpu("value"), cpu("loc_curr"), cpu("aw_curr"), cpu("loc_temporal"),
                                                 cpu("aw_temporal"), torch.tensor(shapes), offs)
    assert nmax(out.cpu().numpy(), ref.numpy()) < 1e-5
