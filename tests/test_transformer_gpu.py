"""GPU parity of the DeVISTransformer mirror (devis_b200/devis_transformer.py) against the reference's own
DeVISTransformer run in float64 in the build container (tests/golden/trunk_*.npz, make_golden.py): every output of
the forward, the gradients of the inputs and the gradients of the attention modules' parameters -- the hot path
exercised exactly the way DeVIS calls it: padded clip (valid ratios < 1), encoder + decoder stack, iterative box
refinement switching the decoder from 2-d reference points to 4-d boxes, all-frames and +-window temporal modes.
"""
import pytest
import torch

from conftest import load_golden, nmax
from test_host_logic import _trunk_from_fixture

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["all", "window"])
@pytest.mark.parametrize("dtype,tol_out,tol_grad", [(torch.float64, 1e-10, 1e-9), (torch.float32, 5e-5, 5e-4)])
def test_transformer_mirror_matches_reference_transformer(tag, dtype, tol_out, tol_grad):
    g = load_golden(f"trunk_{tag}")
    nl, q = int(g["cfg"][6]), int(g["cfg"][11])
    tr = _trunk_from_fixture(g)
    tr.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")})
    tr = tr.to("cuda", dtype)
    if dtype == torch.float64:
        # The valid ratios are float32 in the reference (`.float() / H`, deformable_transformer.py:60-67) and a float32
        # division by a Python scalar rounds differently on the GPU (multiplication by the reciprocal) than on the CPU
        # that produced the fixture: 1 ulp, which the float64 comparison would see as 1e-6.  Feed the fixture's ratios
        # so that everything downstream can be held to float64 accuracy.
        ratios = iter(torch.from_numpy(g["valid_ratios"]).cuda().unbind(1))
        tr.get_valid_ratio = lambda mask: next(ratios)

    def dev(name, grad=False):
        x = torch.from_numpy(g[name]).cuda()
        x = x.to(dtype) if x.is_floating_point() else x
        return x.requires_grad_(True) if grad else x

    srcs = [dev(f"src{i}", True) for i in range(nl)]
    query_embed = dev("query_embed", True)
    hs, qe, memories, init_ref, inter_refs, lsi, valid, shapes = tr(
        srcs, [dev(f"mask{i}") for i in range(nl)], [dev(f"pos{i}") for i in range(nl)], query_embed)
    assert hs.shape == g["hs"].shape and qe.shape == (1, query_embed.shape[0], query_embed.shape[1] // 2)
    assert inter_refs.shape[-1] == 4                     # box refinement turned the points into boxes
    assert nmax(hs.detach().cpu().numpy(), g["hs"]) < tol_out
    assert nmax(init_ref.detach().cpu().numpy(), g["init_ref"]) < tol_out
    assert nmax(inter_refs.detach().cpu().numpy(), g["inter_refs"]) < tol_out
    for i in range(nl):
        assert memories[i].shape == g[f"memory{i}"].shape
        assert nmax(memories[i].detach().cpu().numpy(), g[f"memory{i}"]) < tol_out
    assert torch.equal(lsi.cpu(), torch.from_numpy(g["lsi"])) and torch.equal(shapes.cpu(), torch.from_numpy(g["shapes"]))

    loss = (hs * dev("g_hs")).sum() + sum((m * dev(f"g_memory{i}")).sum() for i, m in enumerate(memories))
    loss.backward()
    for i in range(nl):
        assert nmax(srcs[i].grad.cpu().numpy(), g[f"g_src{i}"]) < tol_grad
    assert nmax(query_embed.grad.cpu().numpy(), g["g_query_embed"]) < tol_grad
    params = dict(tr.named_parameters())
    checked = 0
    for key, want in g.items():
        if key.startswith("pg."):
            got = params[key[3:]].grad
            got = torch.zeros_like(params[key[3:]]) if got is None else got
            assert nmax(got.cpu().numpy(), want) < tol_grad, key
            checked += 1
    assert checked >= 20


def test_transformer_mirror_runs_the_devis_r50_shape_without_host_syncs_in_the_layers():
    """full-size smoke: T=6, 360x640 pyramid, 6+6 layers, 10 queries per frame, forward + backward; the attention
    modules must find the pyramid shape on the host copy attached by prepare_data (no .tolist() on device tensors)"""
    from devis_b200 import DeVISTransformer, clip_geometry, synthetic
    torch.manual_seed(0)
    t_frames, c = 6, 256
    tr = DeVISTransformer(d_model=c, num_frames=t_frames, enc_n_temporal_points=4, dec_n_temporal_points=4, dropout=0.0).cuda()
    srcs = [torch.randn(t_frames, c, h, w, device="cuda", requires_grad=True) for h, w in synthetic.DEVIS_SHAPES]
    masks = [torch.zeros(t_frames, h, w, dtype=torch.bool, device="cuda") for h, w in synthetic.DEVIS_SHAPES]
    pos = [torch.randn(t_frames, c, h, w, device="cuda") for h, w in synthetic.DEVIS_SHAPES]
    query_embed = torch.randn(t_frames * 10, 2 * c, device="cuda", requires_grad=True)
    calls = {"n": 0}
    real = torch.Tensor.tolist

    def counting(self):
        calls["n"] += int(self.is_cuda)
        return real(self)

    torch.Tensor.tolist = counting
    try:
        hs, *_rest = tr(srcs, masks, pos, query_embed)
    finally:
        torch.Tensor.tolist = real
    assert calls["n"] == 0
    assert hs.shape == (6, 1, t_frames * 10, c) and torch.isfinite(hs).all()
    hs.square().mean().backward()
    assert torch.isfinite(srcs[0].grad).all() and float(srcs[0].grad.abs().max()) > 0
    assert clip_geometry.host_list(_rest[-1]) == [list(s) for s in synthetic.DEVIS_SHAPES]


def test_transformer_forward_replays_as_one_cuda_graph():
    """no layer uploads or reads back anything after the first call: the whole trunk forward can be captured"""
    from devis_b200 import DeVISTransformer
    torch.manual_seed(1)
    t_frames, c, shapes_l = 3, 256, ((12, 20), (6, 10))
    tr = DeVISTransformer(d_model=c, num_frames=t_frames, num_encoder_layers=2, num_decoder_layers=2, num_feature_levels=2,
                          enc_n_temporal_points=4, dec_n_temporal_points=4).cuda().eval()
    with torch.no_grad():      # non-zero offset / weight projections like a trained model
        for name, prm in tr.named_parameters():
            if "sampling_offsets.weight" in name or "attention_weights.weight" in name:
                prm.normal_(0, 0.02)
    srcs = [torch.randn(t_frames, c, h, w, device="cuda") for h, w in shapes_l]
    masks = [torch.zeros(t_frames, h, w, dtype=torch.bool, device="cuda") for h, w in shapes_l]
    pos = [torch.randn(t_frames, c, h, w, device="cuda") for h, w in shapes_l]
    query_embed = torch.randn(t_frames * 7, 2 * c, device="cuda")

    def run():
        with torch.no_grad():
            hs, _, memories, *_ = tr(srcs, masks, pos, query_embed)
        return hs, memories[0]

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        hs_g, mem_g = run()
    for trial in range(2):
        for x in srcs:
            x.normal_()
        graph.replay()
        hs_e, mem_e = run()
        assert torch.equal(hs_g, hs_e) and torch.equal(mem_g, mem_e), trial


def test_temporal_encoder_module_under_bf16_autocast():
    """torch.autocast(bf16): the Linear layers emit bf16, so the op sees bf16 value and bf16 offsets / logits (the
    reference's op has no bf16 instantiation, ms_deform_attn_cuda.cu:64).  Output and input gradients must stay within
    bf16 rounding of the fp32 run, on both the fused-prologue and the unfused path."""
    from devis_b200 import TemporalMSDeformAttnEncoder, synthetic
    torch.manual_seed(0)
    T, shapes_l = 3, ((18, 30), (9, 15))
    S = sum(h * w for h, w in shapes_l)
    enc = TemporalMSDeformAttnEncoder(n_frames=T, d_model=256, n_levels=2, t_window=T - 1, n_heads=8, n_curr_points=4,
                                      n_temporal_points=4).cuda()
    with torch.no_grad():
        for lin in (enc.attention_weights, enc.temporal_attention_weights):
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

    def run(amp, fused):
        enc.fuse_prologue = fused
        enc.zero_grad(set_to_none=True)
        q = query.clone().requires_grad_(True)
        x = inp.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            out, _ = enc(q, ref, x, (shapes, tshapes), (lsi, tlsi), offsets)
        assert out.dtype == (torch.bfloat16 if amp else torch.float32)
        out.float().backward(gout)
        return out.detach().float(), q.grad, x.grad

    want = run(False, True)
    for fused in (True, False):
        got = run(True, fused)
        assert got[1].dtype == torch.float32 and got[2].dtype == torch.float32
        for g_, w_ in zip(got, want):
            assert nmax(g_.cpu().numpy(), w_.cpu().numpy()) < 4e-2
