"""devis_b200.GraphedLayer: CUDA-graph replay (forward + backward) of the decoder's temporal cross-attention and of a
whole decoder layer for fixed shapes (reference call site deformable_transformer.py:263).
Replays must give what the eager module gives -- outputs, input gradients, parameter gradients -- for NEW input values,
and must refuse calls whose static arguments differ from the captured ones."""
import copy

import pytest
import torch

from conftest import nmax

pytestmark = pytest.mark.gpu

T, Q, C = 6, 10, 256


def _setup(ref_dim):
    from devis_b200 import synthetic
    torch.manual_seed(0)
    shapes_l = synthetic.DEVIS_SHAPES
    S = sum(h * w for h, w in shapes_l)
    dev = "cuda"
    shapes = torch.tensor(shapes_l, device=dev)
    lsi = torch.tensor(synthetic.level_start_index(shapes_l), device=dev)
    tshapes = shapes.repeat(T - 1, 1)
    tlsi = torch.cat([tshapes.new_zeros(1), tshapes.prod(1).cumsum(0)[:-1]])
    offsets = [torch.tensor([d for d in range(-t, T - t) if d != 0], device=dev) for t in range(T)]

    def inputs(seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        q = torch.randn(1, T * Q, C, device=dev, generator=g).requires_grad_(True)
        ref = torch.rand(1, T * Q, 4, ref_dim, device=dev, generator=g) * 0.6 + 0.2
        if ref_dim == 4:
            ref[..., 2:] = ref[..., 2:] * 0.3
        ref.requires_grad_(True)
        src = torch.randn(T, S, C, device=dev, generator=g).requires_grad_(True)
        gout = torch.randn(1, T * Q, C, device=dev, generator=g)
        return q, ref, src, gout

    return (shapes, tshapes), (lsi, tlsi), offsets, inputs


def _randomize(mod):
    with torch.no_grad():
        for lin in (mod.sampling_offsets, mod.temporal_sampling_offsets, mod.attention_weights, mod.temporal_attention_weights):
            lin.weight.normal_(0, 0.02)


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_graphed_decoder_attention_matches_eager(ref_dim):
    from devis_b200 import GraphedLayer, TemporalMSDeformAttnDecoder, _lib
    shp, lsi, offsets, inputs = _setup(ref_dim)
    dec = TemporalMSDeformAttnDecoder(n_frames=T, d_model=C, n_levels=4, t_window=T - 1, n_heads=8, n_curr_points=4,
                                      n_temporal_points=4).cuda()
    _randomize(dec)
    q0, r0, s0, _ = inputs(1)
    fast = GraphedLayer(copy.deepcopy(dec), q0, r0, s0, shp, lsi, offsets)
    for seed in (2, 3):                                  # values the graph has never seen
        q, r, s, gout = inputs(seed)
        out, lc, lt, awc, awt = dec(q, r, s, shp, lsi, offsets)
        out.backward(gout)
        want = (out.detach(), q.grad, r.grad, s.grad, torch.stack(lc).detach(), awt.detach(),
                {n: p.grad.clone() for n, p in dec.named_parameters()})
        dec.zero_grad(set_to_none=True)
        q2, r2, s2 = (x.detach().clone().requires_grad_(True) for x in (q, r, s))
        fast.zero_grad(set_to_none=True)
        launches = _lib.launch_count()
        out2, lc2, lt2, awc2, awt2 = fast(q2, r2, s2, shp, lsi, offsets)
        out2.backward(gout)
        torch.cuda.synchronize()
        assert _lib.launch_count() == launches          # replays: the library issued no launch of its own
        assert len(lc2) == T and len(lt2) == T and lc2[0].shape == lc[0].shape
        assert nmax(out2.detach().cpu().numpy(), want[0].cpu().numpy()) < 1e-6
        assert nmax(torch.stack(lc2).detach().cpu().numpy(), want[4].cpu().numpy()) < 1e-6
        assert nmax(awt2.detach().cpu().numpy(), want[5].cpu().numpy()) < 1e-6
        for got, ref_g in ((q2.grad, want[1]), (r2.grad, want[2]), (s2.grad, want[3])):
            assert nmax(got.cpu().numpy(), ref_g.cpu().numpy()) < 1e-5     # float atomics: order differs run to run
        for n, p in fast.module.named_parameters():
            assert nmax(p.grad.cpu().numpy(), want[6][n].cpu().numpy()) < 1e-5, n


def test_graphed_layer_refuses_other_static_arguments_and_shapes():
    from devis_b200 import GraphedLayer, TemporalMSDeformAttnDecoder
    shp, lsi, offsets, inputs = _setup(2)
    dec = TemporalMSDeformAttnDecoder(n_frames=T, d_model=C, n_levels=4, t_window=T - 1, n_heads=8, n_curr_points=4,
                                      n_temporal_points=4).cuda()
    q, r, s, _ = inputs(1)
    fast = GraphedLayer(dec, q, r, s, shp, lsi, offsets)
    other = [o.clone() for o in offsets]
    with pytest.raises(RuntimeError, match="static argument"):
        fast(q, r, s, shp, lsi, other)
    with pytest.raises(RuntimeError, match="does not match the captured"):
        fast(q[:, :T * 5], r, s, shp, lsi, offsets)
    with pytest.raises(RuntimeError, match="positional"):
        fast(q, r, s, shp, lsi)


def test_graphed_whole_decoder_layer_matches_eager():
    """self-attention + temporal cross-attention + FFN (DeVISTransformerDecoderLayer), dropout off"""
    from devis_b200 import DeVISTransformerDecoderLayer, GraphedLayer
    shp, lsi, offsets, inputs = _setup(4)
    layer = DeVISTransformerDecoderLayer(d_model=C, d_ffn=1024, dropout=0.0, n_frames=T, t_window=T - 1, n_levels=4,
                                         n_heads=8, n_curr_points=4, n_temporal_points=4).cuda()
    _randomize(layer.cross_attn)
    q0, r0, s0, _ = inputs(1)
    pos0 = torch.randn_like(q0).requires_grad_(True)
    fast = GraphedLayer(copy.deepcopy(layer), q0, pos0, r0, s0, shp, lsi, temporal_offsets=offsets)
    q, r, s, gout = inputs(5)
    pos = torch.randn_like(q).requires_grad_(True)
    out = layer(q, pos, r, s, shp, lsi, temporal_offsets=offsets)
    out.backward(gout)
    q2, pos2, r2, s2 = (x.detach().clone().requires_grad_(True) for x in (q, pos, r, s))
    out2 = fast(q2, pos2, r2, s2, shp, lsi, temporal_offsets=offsets)
    out2.backward(gout)
    assert nmax(out2.detach().cpu().numpy(), out.detach().cpu().numpy()) < 1e-5
    for got, want in ((q2.grad, q.grad), (pos2.grad, pos.grad), (s2.grad, s.grad), (r2.grad, r.grad)):
        assert nmax(got.cpu().numpy(), want.cpu().numpy()) < 1e-4
    for (n, p), (_, p0) in zip(fast.module.named_parameters(), layer.named_parameters()):
        assert nmax(p.grad.cpu().numpy(), p0.grad.cpu().numpy()) < 1e-4, n
