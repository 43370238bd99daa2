"""oracle/msda_torch.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PyTorch (CPU or any device) restatement of the reference's pure-PyTorch path for
multi-scale deformable attention.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs may import this module; the
product package (devis_b200/) never does.

Follows /root/reference/src/models/ops/functions/ms_deform_attn_func.py:102-122
(``ms_deform_attn_core_pytorch``): per level, view the level's rows of ``value``
as an image batch of N*M maps with D channels, sample it with
``F.grid_sample(bilinear, zeros, align_corners=False)`` at ``2*loc-1`` and take
the attention-weighted sum over (level, point).  The backward is torch autograd
through that graph, exactly what the reference gets when it differentiates its
PyTorch path.

Parity pin: tests/test_oracle_golden.py compares this restatement with fixtures
produced by importing the reference file itself (tests/golden/make_golden.py).
"""
import torch
import torch.nn.functional as F


def msda_forward_torch(value, spatial_shapes, sampling_locations, attention_weights):
    """value (N,S,M,D); spatial_shapes (L,2) rows (H,W); sampling_locations
    (N,Lq,M,L,P,2) as (x,y) in [0,1]; attention_weights (N,Lq,M,L,P) -> (N,Lq,M*D).

    ``level_start_index`` is implied by the level order (the reference's PyTorch
    path ignores it too, ms_deform_attn_func.py:107)."""
    n, _, m, d = value.shape
    _, lq, _, nl, p, _ = sampling_locations.shape
    sizes = [int(h) * int(w) for h, w in spatial_shapes.tolist()]
    grids = sampling_locations * 2 - 1                        # grid_sample's [-1,1] frame
    per_level = []
    start = 0
    for lvl, (h, w) in enumerate(spatial_shapes.tolist()):
        rows = value[:, start:start + sizes[lvl]]             # (N, H*W, M, D)
        start += sizes[lvl]
        img = rows.permute(0, 2, 3, 1).reshape(n * m, d, int(h), int(w))
        grid = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(n * m, lq, p, 2)
        per_level.append(F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros",
                                       align_corners=False))  # (N*M, D, Lq, P)
    sampled = torch.stack(per_level, dim=3).reshape(n * m, d, lq, nl * p)
    weights = attention_weights.permute(0, 2, 1, 3, 4).reshape(n * m, 1, lq, nl * p)
    out = (sampled * weights).sum(-1)                         # (N*M, D, Lq)
    return out.reshape(n, m * d, lq).transpose(1, 2).contiguous()


def msda_forward_backward_torch(value, spatial_shapes, sampling_locations, attention_weights,
                                grad_output):
    """Forward plus autograd backward; returns (out, grad_value, grad_loc, grad_aw)."""
    v = value.detach().clone().requires_grad_(True)
    loc = sampling_locations.detach().clone().requires_grad_(True)
    aw = attention_weights.detach().clone().requires_grad_(True)
    out = msda_forward_torch(v, spatial_shapes, loc, aw)
    gv, gl, ga = torch.autograd.grad(out, (v, loc, aw), grad_output)
    return out.detach(), gv, gl, ga


def level_start_index_of(spatial_shapes):
    """ms_deform_attn/test.py:22 and deformable_transformer.py prepare_data: exclusive cumsum of H*W."""
    areas = spatial_shapes[:, 0] * spatial_shapes[:, 1]
    return torch.cat([areas.new_zeros(1), areas.cumsum(0)[:-1]])
