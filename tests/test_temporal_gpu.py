"""GPU parity of the whole-clip temporal op and of the temporal modules.

 * against the golden module fixtures (outputs and gradients of the reference's own
   TemporalMSDeformAttnEncoder / Decoder / MSDeformAttn modules, float64),
 * against the PyTorch oracle's per-frame loop (oracle/temporal_torch.py) at a reduced clip,
 * and, at the full DeVIS R50 T=6 shape, through size-independent properties:
   whole-clip == the reference's 2T per-frame calls made with the drop-in op; linearity of the op in
   value and in the attention weights; Euler identities <grad_value, value> = <grad_aw, aw> = <grad_out, out>.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, nmax

pytestmark = pytest.mark.gpu


def _t(g, k, dtype=None):
    x = torch.from_numpy(g[k]).cuda()
    return x.to(dtype) if dtype is not None and x.is_floating_point() else x


def _load_sd(mod, g, dtype):
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}
    assert set(sd) == set(mod.state_dict())            # parameter names are the reference's
    mod = mod.double()                                  # fixtures are float64: load before any down-cast
    mod.load_state_dict(sd)
    return mod.to("cuda", dtype)


@pytest.mark.parametrize("tag", ["2d", "4d"])
def test_msdeformattn_module_matches_reference_module(tag):
    from devis_b200 import MSDeformAttn
    g = load_golden(f"mod_msda_{tag}")
    c, nl, heads, pts = [int(x) for x in g["cfg"]]
    for dtype, tol_o, tol_g in ((torch.float64, 1e-11, 1e-10), (torch.float32, 2e-5, 2e-4)):
        mod = _load_sd(MSDeformAttn(c, nl, heads, pts), g, dtype)
        q = _t(g, "query", dtype).requires_grad_(True)
        x = _t(g, "inp", dtype).requires_grad_(True)
        out, none = mod(q, _t(g, "ref", dtype), x, _t(g, "shapes"), _t(g, "lsi"), _t(g, "mask"))
        assert none is None
        out.backward(_t(g, "gout", dtype))
        assert nmax(out.detach().cpu().numpy(), g["out"]) < tol_o
        assert nmax(q.grad.cpu().numpy(), g["gquery"]) < tol_g
        assert nmax(x.grad.cpu().numpy(), g["ginp"]) < tol_g


@pytest.mark.parametrize("tag", ["all", "window"])
def test_temporal_encoder_module_matches_reference_module(tag):
    from devis_b200 import TemporalMSDeformAttnEncoder
    g = load_golden(f"mod_tenc_{tag}")
    t_frames, c, nl, t_window, heads, pc, pt = [int(x) for x in g["cfg"]]
    for dtype, tol_o, tol_g in ((torch.float64, 1e-11, 1e-10), (torch.float32, 2e-5, 2e-4)):
        mod = _load_sd(TemporalMSDeformAttnEncoder(t_frames, c, nl, t_window, heads, pc, pt), g, dtype)
        q = _t(g, "query", dtype).requires_grad_(True)
        x = _t(g, "inp", dtype).requires_grad_(True)
        offs = [row for row in _t(g, "temporal_offsets")]
        out, none = mod(q, _t(g, "ref", dtype), x, (_t(g, "shapes"), _t(g, "tshapes")), (_t(g, "lsi"), _t(g, "tlsi")), offs)
        assert none is None
        out.backward(_t(g, "gout", dtype))
        assert nmax(out.detach().cpu().numpy(), g["out"]) < tol_o
        assert nmax(q.grad.cpu().numpy(), g["gquery"]) < tol_g
        assert nmax(x.grad.cpu().numpy(), g["ginp"]) < tol_g
        for name, prm in mod.named_parameters():
            assert nmax(prm.grad.cpu().numpy(), g["pg." + name]) < tol_g, name


@pytest.mark.parametrize("tag", ["2d_ia", "2d_noia", "4d_ia", "4d_noia"])
def test_temporal_decoder_module_matches_reference_module(tag):
    from devis_b200 import TemporalMSDeformAttnDecoder
    g = load_golden(f"mod_tdec_{tag}")
    t_frames, c, nl, t_window, heads, pc, pt, ia = [int(x) for x in g["cfg"]]
    for dtype, tol_o, tol_g in ((torch.float64, 1e-11, 1e-10), (torch.float32, 2e-5, 2e-4)):
        mod = _load_sd(TemporalMSDeformAttnDecoder(t_frames, c, nl, t_window, heads, pc, pt, bool(ia)), g, dtype)
        q = _t(g, "query", dtype).requires_grad_(True)
        x = _t(g, "inp", dtype).requires_grad_(True)
        r = _t(g, "ref", dtype).requires_grad_(True)
        offs = [row for row in _t(g, "temporal_offsets")]
        out, lc, lt, awc, awt = mod(q, r, x, (_t(g, "shapes"), _t(g, "tshapes")), (_t(g, "lsi"), _t(g, "tlsi")), offs)
        out.backward(_t(g, "gout", dtype))
        assert out.shape == g["out"].shape and len(lc) == t_frames and len(lt) == t_frames
        assert tuple(lc[0].shape) == g["loc_curr"].shape[1:] and tuple(lt[0].shape) == g["loc_temporal"].shape[1:]
        assert nmax(out.detach().cpu().numpy(), g["out"]) < tol_o
        assert nmax(torch.stack(lc).detach().cpu().numpy(), g["loc_curr"]) < tol_o
        assert nmax(torch.stack(lt).detach().cpu().numpy(), g["loc_temporal"]) < tol_o
        assert nmax(awc.detach().cpu().numpy(), g["aw_curr"]) < tol_o
        assert nmax(awt.detach().cpu().numpy(), g["aw_temporal"]) < tol_o
        assert nmax(q.grad.cpu().numpy(), g["gquery"]) < tol_g
        assert nmax(x.grad.cpu().numpy(), g["ginp"]) < tol_g
        assert nmax(r.grad.cpu().numpy(), g["gref"]) < tol_g


def _clip_fn(clip, dtype=None, order=None):
    from devis_b200 import clip_geometry, temporal_ms_deform_attn
    geom = clip_geometry.ClipGeometry(clip["shapes"], clip["value"].shape[0], clip["frame_table"])
    leaves = [clip[k].clone().requires_grad_(True) for k in ("value", "loc_curr", "aw_curr", "loc_temporal", "aw_temporal")]
    out = temporal_ms_deform_attn(*leaves, geom, order)
    out.backward(clip["grad_out"])
    return out.detach(), [x.grad for x in leaves], geom


@pytest.mark.parametrize("dist", ["local", "uniform"])
@pytest.mark.parametrize("queries", [None, 7])
def test_whole_clip_op_matches_pytorch_oracle_per_frame_loop(dist, queries):
    """reduced clip (T=4, 3 levels) so the CPU oracle runs in seconds; fp32 CUDA vs fp64 oracle"""
    from devis_b200 import synthetic
    from oracle import temporal_torch
    shapes = ((18, 30), (9, 15), (5, 8))
    clip = synthetic.make_clip(n_frames=4, shapes=shapes, queries=queries, dist=dist, seed=3, device="cuda")
    out, grads, _ = _clip_fn(clip)
    cpu = {k: (v.detach().double().cpu() if isinstance(v, torch.Tensor) else v) for k, v in clip.items()}
    leaves = [cpu[k].clone().requires_grad_(True) for k in ("value", "loc_curr", "aw_curr", "loc_temporal", "aw_temporal")]
    offs = [torch.tensor([f - t for f in row]) for t, row in enumerate(clip["frame_table"])]
    ref = temporal_torch.temporal_core_per_frame(*leaves, torch.tensor(shapes), offs)
    ref.backward(cpu["grad_out"])
    assert nmax(out.cpu().numpy(), ref.detach().numpy()) < 1e-5
    for got, leaf in zip(grads, leaves):
        assert nmax(got.cpu().numpy(), leaf.grad.numpy()) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_whole_clip_equals_reference_call_sequence_at_full_devis_shape(dtype):
    """the reference's 2T calls + T gather copies (ms_deform_attn.py:435-460) issued with the drop-in op
    must give what the single whole-clip launch gives -- full R50 T=6 encoder shape"""
    from devis_b200 import MSDeformAttnFunction, synthetic
    clip = synthetic.make_clip(dist="local", seed=1, dtype=dtype, device="cuda")
    out, grads, geom = _clip_fn(clip)
    order = geom.tile_order("cuda")
    out_tiled, grads_tiled, _ = _clip_fn(clip, order=order)
    shapes = torch.tensor(clip["shapes"], device="cuda")
    areas = shapes.prod(1)
    lsi = torch.cat([areas.new_zeros(1), areas.cumsum(0)[:-1]])
    wt = len(clip["frame_table"][0])
    tshapes = shapes.repeat(wt, 1)
    tareas = tshapes.prod(1)
    tlsi = torch.cat([tareas.new_zeros(1), tareas.cumsum(0)[:-1]])
    leaves = [clip[k].clone().requires_grad_(True) for k in ("value", "loc_curr", "aw_curr", "loc_temporal", "aw_temporal")]
    v, lc, ac, lt, at = leaves
    frames_out = []
    for t in range(v.shape[0]):
        cur = MSDeformAttnFunction.apply(v[t][None], shapes, lsi, lc[t][None], ac[t][None], 64)
        stacked = v[clip["frame_table"][t]].flatten(0, 1)[None]
        tmp = MSDeformAttnFunction.apply(stacked, tshapes, tlsi, lt[t][None], at[t][None], 64)
        frames_out.append(cur + tmp)
    ref = torch.cat(frames_out, 0)
    ref.backward(clip["grad_out"])
    tol_o, tol_g = (2e-6, 2e-5) if dtype == torch.float32 else (1e-2, 2e-2)
    for o, gs in ((out, grads), (out_tiled, grads_tiled)):
        assert nmax(o.float().cpu().numpy(), ref.detach().float().cpu().numpy()) < tol_o
        for got, leaf in zip(gs, leaves):
            assert nmax(got.float().cpu().numpy(), leaf.grad.float().cpu().numpy()) < tol_g
    assert torch.equal(out, out_tiled)            # the visiting order must not change any output bit


def test_full_shape_linearity_and_euler_identities():
    from devis_b200 import synthetic
    clip = synthetic.make_clip(dist="local", seed=2, device="cuda")
    out, grads, geom = _clip_fn(clip)
    gv, glc, gac, glt, gat = grads
    gout = clip["grad_out"]
    total = (gout.double() * out.double()).sum().item()
    # out is linear in value, and linear in the (current, temporal) weights jointly
    assert abs((gv.double() * clip["value"].double()).sum().item() - total) < 1e-6 * abs(total) + 1e-3
    aw_side = (gac.double() * clip["aw_curr"].double()).sum().item() + (gat.double() * clip["aw_temporal"].double()).sum().item()
    assert abs(aw_side - total) < 1e-6 * abs(total) + 1e-3
    # linearity in value: f(2.5*v + v2) = 2.5*f(v) + f(v2)
    from devis_b200 import temporal_ms_deform_attn
    v2 = torch.randn_like(clip["value"])
    args = (clip["loc_curr"], clip["aw_curr"], clip["loc_temporal"], clip["aw_temporal"], geom)
    lhs = temporal_ms_deform_attn(2.5 * clip["value"] + v2, *args)
    rhs = 2.5 * out + temporal_ms_deform_attn(v2, *args)
    assert nmax(lhs.cpu().numpy(), rhs.cpu().numpy()) < 1e-5
    # a constant value map is reproduced exactly where all taps fall inside: sum of weights (=1) * const
    const = torch.ones_like(clip["value"])
    inside = temporal_ms_deform_attn(const, *args)
    assert inside.max().item() <= 1 + 1e-5 and inside.min().item() >= -1e-6


def test_decoder_shape_300_queries_whole_clip_vs_per_frame():
    from devis_b200 import synthetic
    from oracle import c_oracle
    clip = synthetic.make_clip(queries=300, dist="uniform", seed=4, device="cuda")
    out, grads, _ = _clip_fn(clip)
    # frame 3 via the C oracle: current call + temporal call on the gathered frames
    t = 3
    shapes = np.array(clip["shapes"], dtype=np.int64)
    areas = shapes.prod(1)
    lsi = np.concatenate([[0], np.cumsum(areas)[:-1]])
    cur = c_oracle.forward(clip["value"][t][None].cpu().numpy(), shapes, lsi, clip["loc_curr"][t][None].cpu().numpy(),
                           clip["aw_curr"][t][None].cpu().numpy())
    frames = clip["frame_table"][t]
    tshapes = np.tile(shapes, (len(frames), 1))
    tlsi = np.concatenate([[0], np.cumsum(tshapes.prod(1))[:-1]])
    tmp = c_oracle.forward(clip["value"][frames].flatten(0, 1)[None].cpu().numpy(), tshapes, tlsi,
                           clip["loc_temporal"][t][None].cpu().numpy(), clip["aw_temporal"][t][None].cpu().numpy())
    assert nmax(out[t].cpu().numpy(), (cur + tmp)[0]) < 1e-5


def test_duplicate_frames_in_table_and_no_temporal_window():
    """window mode can list a frame twice (devis_transformer.py:102-112); t_window = 0 degenerates to the plain op"""
    from devis_b200 import clip_geometry, synthetic, temporal_ms_deform_attn
    from oracle import temporal_torch
    shapes = ((10, 12), (5, 6))
    clip = synthetic.make_clip(n_frames=4, shapes=shapes, queries=9, dist="uniform", seed=8, device="cuda", t_window=2)
    table = clip_geometry.window_table(4, 2)
    assert table[0] == [1, 1]
    geom = clip_geometry.ClipGeometry(shapes, 4, table)
    out = temporal_ms_deform_attn(clip["value"], clip["loc_curr"], clip["aw_curr"], clip["loc_temporal"],
                                  clip["aw_temporal"], geom)
    offs = [torch.tensor([f - t for f in row]) for t, row in enumerate(table)]
    cpu = lambda k: clip[k].double().cpu()
    ref = temporal_torch.temporal_core_per_frame(cpu("value"), cpu("loc_curr"), cpu("aw_curr"), cpu("loc_temporal"),
                                                 cpu("aw_temporal"), torch.tensor(shapes), offs)
    assert nmax(out.cpu().numpy(), ref.numpy()) < 1e-5
    geom0 = clip_geometry.ClipGeometry(shapes, 4, None)
    out0 = temporal_ms_deform_attn(clip["value"], clip["loc_curr"], clip["aw_curr"], None, None, geom0)
    from oracle import msda_torch
    ref0 = msda_torch.msda_forward_torch(cpu("value"), torch.tensor(shapes), cpu("loc_curr"), cpu("aw_curr"))
    assert nmax(out0.cpu().numpy(), ref0.numpy()) < 1e-5


def test_deterministic_mode_is_bit_reproducible_and_accurate():
    """DEVIS_MSDA_FLAG_DETERMINISTIC: fixed-point integer accumulation of grad_value -- identical bits on every
    run, and at least as close to the fp64 oracle as the default float-atomic mode."""
    from devis_b200 import MultiScaleDeformableAttention as MSDA, synthetic
    from oracle import temporal_torch
    shapes = ((18, 30), (9, 15), (5, 8))
    clip = synthetic.make_clip(n_frames=4, shapes=shapes, dist="local", seed=13, device="cuda")
    _, grads_default, _ = _clip_fn(clip)
    MSDA.set_deterministic(True)
    try:
        runs = [_clip_fn(clip)[1] for _ in range(3)]
    finally:
        MSDA.set_deterministic(False)
    for other in runs[1:]:
        for a, b in zip(runs[0], other):
            assert torch.equal(a, b)
    cpu = {k: (v.detach().double().cpu() if isinstance(v, torch.Tensor) else v) for k, v in clip.items()}
    leaves = [cpu[k].clone().requires_grad_(True) for k in ("value", "loc_curr", "aw_curr", "loc_temporal", "aw_temporal")]
    offs = [torch.tensor([f - t for f in row]) for t, row in enumerate(clip["frame_table"])]
    ref = temporal_torch.temporal_core_per_frame(*leaves, torch.tensor(shapes), offs)
    ref.backward(cpu["grad_out"])
    want = leaves[0].grad.numpy()
    err_det = nmax(runs[0][0].cpu().numpy(), want)
    err_def = nmax(grads_default[0].cpu().numpy(), want)
    assert err_det < 1e-4 and err_det <= err_def * 1.5 + 1e-7
    # the other gradients do not depend on the mode
    for a, b in zip(runs[0][1:], grads_default[1:]):
        assert torch.equal(a, b)


def test_deterministic_mode_drop_in_op_bf16_and_generic_channels():
    from devis_b200 import MSDeformAttnFunction, MultiScaleDeformableAttention as MSDA
    for name, dtype in (("op_d32", torch.bfloat16), ("op_d30", torch.float32), ("op_ragged", torch.float32)):
        g = load_golden(name)
        f = lambda k: torch.from_numpy(g[k]).cuda()
        value, gout = f("value").to(dtype), f("gout").to(dtype)
        loc, aw = f("loc").float(), f("aw").float()

        def run():
            v = value.clone().requires_grad_(True)
            out = MSDeformAttnFunction.apply(v, f("shapes"), f("lsi"), loc, aw, 64)
            out.backward(gout)
            return v.grad
        MSDA.set_deterministic(True)
        try:
            a, b = run(), run()
        finally:
            MSDA.set_deterministic(False)
        assert torch.equal(a, b)
        tol = 2e-2 if dtype == torch.bfloat16 else 1e-4
        assert nmax(a.float().cpu().numpy(), g["gvalue"]) < tol
    with torch.no_grad():
        pass


def test_whole_clip_op_is_cuda_graph_capturable():
    """launch-bound regime (decoder shapes): the op must be capturable so a layer can be replayed as a CUDA graph"""
    from devis_b200 import clip_geometry, synthetic, temporal_ms_deform_attn
    clip = synthetic.make_clip(queries=30, dist="uniform", seed=6, device="cuda")
    geom = clip_geometry.ClipGeometry(clip["shapes"], 6, clip["frame_table"])
    args = [clip[k] for k in ("value", "loc_curr", "aw_curr", "loc_temporal", "aw_temporal")]
    eager = temporal_ms_deform_attn(*args, geom)
    static_value = clip["value"].clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            temporal_ms_deform_attn(static_value, *args[1:], geom)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_out = temporal_ms_deform_attn(static_value, *args[1:], geom)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, eager)
    static_value.mul_(2.0)
    graph.replay()
    torch.cuda.synchronize()
    assert nmax(static_out.cpu().numpy(), (2.0 * eager).cpu().numpy()) < 1e-6


def test_fused_prologue_matches_unfused_path_and_reference_module():
    """TemporalMSDeformAttnFusedFunction (softmax + location arithmetic inside the kernels) against the unfused module
    path at the full DeVIS encoder shape, forward and every gradient; and against the golden reference-module fixture
    (fp32; its d_model/heads give D = 8, which the fused kernels do not serve -> must fall back, same numbers)."""
    from devis_b200 import TemporalMSDeformAttnEncoder, synthetic
    torch.manual_seed(0)
    T, shapes_l = 6, synthetic.DEVIS_SHAPES
    S = sum(h * w for h, w in shapes_l)
    enc = TemporalMSDeformAttnEncoder(n_frames=T, d_model=256, n_levels=4, t_window=T - 1, n_heads=8, n_curr_points=4,
                                      n_temporal_points=4).cuda()
    with torch.no_grad():
        for lin in (enc.sampling_offsets, enc.temporal_sampling_offsets, enc.attention_weights, enc.temporal_attention_weights):
            lin.weight.normal_(0, 0.05)
    query = torch.randn(T, S, 256, device="cuda")
    inp = torch.randn(T, S, 256, device="cuda")
    ref = synthetic.pixel_reference_points(shapes_l, T, "cuda")
    shapes = torch.tensor(shapes_l, device="cuda")
    lsi = torch.tensor(synthetic.level_start_index(shapes_l), device="cuda")
    tshapes = shapes.repeat(T - 1, 1)
    tlsi = torch.cat([tshapes.new_zeros(1), tshapes.prod(1).cumsum(0)[:-1]])
    offsets = [torch.tensor([d for d in range(-t, T - t) if d != 0], device="cuda") for t in range(T)]
    gout = torch.randn(T, S, 256, device="cuda")
    results = {}
    for fused in (True, False):
        enc.fuse_prologue = fused
        enc.zero_grad(set_to_none=True)
        q = query.clone().requires_grad_(True)
        x = inp.clone().requires_grad_(True)
        out, _ = enc(q, ref, x, (shapes, tshapes), (lsi, tlsi), offsets)
        out.backward(gout)
        results[fused] = (out.detach(), q.grad, x.grad, {n: p.grad.clone() for n, p in enc.named_parameters()})
    f, u = results[True], results[False]
    assert nmax(f[0].cpu().numpy(), u[0].cpu().numpy()) < 1e-5
    assert nmax(f[1].cpu().numpy(), u[1].cpu().numpy()) < 2e-4
    assert nmax(f[2].cpu().numpy(), u[2].cpu().numpy()) < 2e-4
    for name in f[3]:
        assert nmax(f[3][name].cpu().numpy(), u[3][name].cpu().numpy()) < 2e-4, name


@pytest.mark.parametrize("M", [8, 3])
def test_fused_function_matches_pytorch_oracle_on_small_clip(M):
    """fused op vs the fp64 PyTorch oracle (projection-free): logits/offsets -> softmax/locations -> per-frame loop.
    8 heads: the value-row size is a compile-time immediate of the dead-corner consumer; 3 heads: the run-time form."""
    from devis_b200 import TemporalMSDeformAttnFusedFunction, clip_geometry, synthetic
    from oracle import temporal_torch
    torch.manual_seed(1)
    shapes_l, T, D, pc, pt = ((18, 30), (9, 15), (5, 8)), 4, 32, 4, 4
    nl, wt = len(shapes_l), T - 1
    S = sum(h * w for h, w in shapes_l)
    geom = clip_geometry.ClipGeometry(shapes_l, T, clip_geometry.all_frames_table(T))
    ref = synthetic.pixel_reference_points(shapes_l, T, "cuda")
    value = torch.randn(T, S, M, D, device="cuda")
    # offsets in pixels, kept away from cell borders by construction of the oracle comparison tolerance
    off_c = (2.0 * torch.randn(T, S, M, nl, pc, 2, device="cuda")).requires_grad_(True)
    off_t = (2.0 * torch.randn(T, S, M, wt * nl, pt, 2, device="cuda")).requires_grad_(True)
    lg_c = torch.randn(T, S, M, nl * pc, device="cuda").requires_grad_(True)
    lg_t = torch.randn(T, S, M, wt * nl * pt, device="cuda").requires_grad_(True)
    v = value.clone().requires_grad_(True)
    out = TemporalMSDeformAttnFusedFunction.apply(v, ref, off_c, lg_c, off_t, lg_t, geom, geom.tile_order("cuda"))
    gout = torch.randn_like(out)
    out.backward(gout)

    d = lambda t: t.detach().double().cpu()
    shapes = torch.tensor(shapes_l)
    o_c, o_t, l_c, l_t, vv = (d(x).requires_grad_(True) for x in (off_c, off_t, lg_c, lg_t, value))
    joint = torch.softmax(torch.cat([l_c, l_t], -1), -1)
    aw_c = joint[..., :nl * pc].reshape(T, S, M, nl, pc)
    aw_t = joint[..., nl * pc:].reshape(T, S, M, wt * nl, pt)
    loc_c, loc_t = temporal_torch.encoder_locations(d(ref), o_c, o_t, shapes, wt)
    offs = temporal_torch.all_frames_offsets(T)
    want = temporal_torch.temporal_core_per_frame(vv, loc_c, aw_c, loc_t, aw_t, shapes, offs)
    want.backward(d(gout))
    assert nmax(out.detach().cpu().numpy(), want.detach().numpy()) < 1e-5
    # taps are not boundary-safe here (raw random offsets): gate location gradients on a quantile, others on max
    assert nmax(v.grad.cpu().numpy(), vv.grad.numpy()) < 1e-4
    assert nmax(lg_c.grad.cpu().numpy(), l_c.grad.numpy()) < 1e-4
    assert nmax(lg_t.grad.cpu().numpy(), l_t.grad.numpy()) < 1e-4
    for got, ref_g in ((off_c.grad, o_c.grad), (off_t.grad, o_t.grad)):
        err = (got.double().cpu() - ref_g).abs() / ref_g.abs().max()
        assert torch.quantile(err.flatten()[:2_000_000], 0.999) < 1e-4


def test_bf16_grad_value_accumulation_mode():
    """DEVIS_MSDA_FLAG_BF16_GRAD_VALUE (opt-in): bf16 grad_value accumulated with packed bf16 reductions.
    Checked against the fp64 PyTorch oracle on bf16-rounded inputs; every partial sum rounds to bf16, so the bound is
    a bf16 one (the float-accumulated default is checked next to it and must be tighter); the other gradients are
    bit-identical in both modes; unsupported combinations are refused by the library, not silently served."""
    from devis_b200 import MSDeformAttnFunction, MultiScaleDeformableAttention as MSDA, _lib, synthetic
    from oracle import temporal_torch
    shapes = ((18, 30), (9, 15), (5, 8))
    clip = synthetic.make_clip(n_frames=4, shapes=shapes, queries=None, dist="local", seed=11, dtype=torch.bfloat16,
                               device="cuda")
    _, grads_default, _ = _clip_fn(clip)
    MSDA.set_bf16_grad_value_accumulation(True)
    try:
        _, grads_half, _ = _clip_fn(clip)
        g = load_golden("op_d32")
        f = lambda k: torch.from_numpy(g[k]).cuda()
        v = f("value").bfloat16().requires_grad_(True)
        out = MSDeformAttnFunction.apply(v, f("shapes"), f("lsi"), f("loc").float(), f("aw").float(), 64)
        out.backward(f("gout").bfloat16())
        assert v.grad.dtype == torch.bfloat16
        assert nmax(v.grad.float().cpu().numpy(), g["gvalue"]) < 1e-1
    finally:
        MSDA.set_bf16_grad_value_accumulation(False)
    cpu = {k: (t.detach().double().cpu() if isinstance(t, torch.Tensor) else t) for k, t in clip.items()}
    leaves = [cpu[k].clone().requires_grad_(True) for k in ("value", "loc_curr", "aw_curr", "loc_temporal", "aw_temporal")]
    offs = [torch.tensor([fr - t for fr in row]) for t, row in enumerate(clip["frame_table"])]
    ref = temporal_torch.temporal_core_per_frame(*leaves, torch.tensor(shapes), offs)
    ref.backward(cpu["grad_out"])
    want = leaves[0].grad.numpy()
    assert grads_half[0].dtype == torch.bfloat16
    err_half = nmax(grads_half[0].float().cpu().numpy(), want)
    err_default = nmax(grads_default[0].float().cpu().numpy(), want)
    # measured: 2.3e-3 (float accumulation, one rounding) against 4.3e-2 (a few hundred bf16 roundings per element)
    assert err_default < 1e-2 and err_half < 1e-1, (err_default, err_half)
    rms = float(np.sqrt(np.mean((grads_half[0].float().cpu().numpy() - want) ** 2)) / np.abs(want).max())
    assert rms < 1e-2, rms
    for a, b in zip(grads_half[1:], grads_default[1:]):
        assert torch.equal(a, b)
    # fp32 value + the flag, or the flag + deterministic mode: refused
    lib = _lib.load()
    z = torch.zeros(1, 4, 8, 32, device="cuda")
    shp = torch.tensor([[2, 2]], device="cuda")
    lsi = torch.zeros(1, dtype=torch.int64, device="cuda")
    loc, aw = torch.rand(1, 3, 8, 1, 4, 2, device="cuda"), torch.rand(1, 3, 8, 1, 4, device="cuda")
    go, gl, ga = torch.zeros(1, 3, 256, device="cuda"), torch.empty_like(loc), torch.empty_like(aw)
    rc = lib.devis_msda_backward(z.data_ptr(), shp.data_ptr(), lsi.data_ptr(), loc.data_ptr(), aw.data_ptr(), go.data_ptr(),
                                 z.clone().data_ptr(), gl.data_ptr(), ga.data_ptr(), 1, 4, 8, 32, 1, 3, 4, 64, _lib.F32,
                                 _lib.FLAG_BF16_GRAD_VALUE, None, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == -8
    torch.cuda.synchronize()


def test_bf16_grad_value_accumulation_fused_encoder_module():
    """the fused-prologue encoder path under the bf16 accumulation switch: input gradients stay within bf16 bounds of
    the default mode"""
    from devis_b200 import MultiScaleDeformableAttention as MSDA, TemporalMSDeformAttnEncoder, synthetic
    torch.manual_seed(0)
    T, shapes_l = 3, ((18, 30), (9, 15))
    S = sum(h * w for h, w in shapes_l)
    enc = TemporalMSDeformAttnEncoder(n_frames=T, d_model=256, n_levels=2, t_window=T - 1, n_heads=8, n_curr_points=4,
                                      n_temporal_points=4).cuda().bfloat16()
    query = torch.randn(T, S, 256, device="cuda").bfloat16()
    inp = torch.randn(T, S, 256, device="cuda").bfloat16()
    ref = synthetic.pixel_reference_points(shapes_l, T, "cuda")
    shapes = torch.tensor(shapes_l, device="cuda")
    lsi = torch.tensor(synthetic.level_start_index(shapes_l), device="cuda")
    tshapes = shapes.repeat(T - 1, 1)
    tlsi = torch.cat([tshapes.new_zeros(1), tshapes.prod(1).cumsum(0)[:-1]])
    offsets = [torch.tensor([d for d in range(-t, T - t) if d != 0], device="cuda") for t in range(T)]
    gout = torch.randn(T, S, 256, device="cuda").bfloat16()
    res = {}
    for half in (False, True):
        MSDA.set_bf16_grad_value_accumulation(half)
        try:
            enc.zero_grad(set_to_none=True)
            x = inp.clone().requires_grad_(True)
            out, _ = enc(query, ref, x, (shapes, tshapes), (lsi, tlsi), offsets)
            out.backward(gout)
            res[half] = (out.detach().float(), x.grad.float())
        finally:
            MSDA.set_bf16_grad_value_accumulation(False)
    assert torch.equal(res[False][0], res[True][0])
    assert nmax(res[True][1].cpu().numpy(), res[False][1].cpu().numpy()) < 5e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_deterministic_sorted_backward_is_bit_identical_to_direct_deterministic_scatter(dtype):
    """Deterministic mode in the encoder form (query order given) runs msda_bwds_kernel<DET>: the same 64-bit fixed-point
    integers as the direct deterministic scatter, pre-summed per row run in registers -- integer addition is associative,
    so grad_value must be bit-identical to the direct path (tuning key 6 = 1) and identical run to run; full DeVIS shape
    (local and uniform taps) and a ragged 4-level pyramid."""
    from devis_b200 import MultiScaleDeformableAttention as MSDA, _lib, clip_geometry, synthetic
    cases = [dict(dist="local", seed=21), dict(dist="uniform", seed=22),
             dict(n_frames=4, shapes=((18, 30), (9, 15), (5, 8), (3, 4)), dist="local", seed=23)]
    MSDA.set_deterministic(True)
    try:
        for kw in cases:
            clip = synthetic.make_clip(device="cuda", dtype=dtype, **kw)
            geom = clip_geometry.ClipGeometry(clip["shapes"], clip["value"].shape[0], clip["frame_table"])
            order = geom.tile_order("cuda")
            _lib.set_tuning(6, 1)
            _, direct, _ = _clip_fn(clip, order=order)
            _lib.set_tuning(6, 0)
            launches = _lib.load().devis_msda_launch_count()
            _, srt, _ = _clip_fn(clip, order=order)
            _, again, _ = _clip_fn(clip, order=order)
            assert _lib.load().devis_msda_launch_count() > launches
            for a, b, c in zip(srt, direct, again):
                assert torch.equal(a, b) and torch.equal(a, c), kw
    finally:
        _lib.set_tuning(6, 0)
        MSDA.set_deterministic(False)
