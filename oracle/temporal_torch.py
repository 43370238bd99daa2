"""oracle/temporal_torch.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PyTorch restatement of the reference's *temporal* multi-scale deformable attention
(/root/reference/src/models/ops/modules/ms_deform_attn.py), written functionally:
parameters come in as a ``state_dict``-style mapping with the reference's names
(``sampling_offsets.weight`` ... ``output_proj.bias``).  Every op call goes through
oracle.msda_torch.msda_forward_torch, i.e. the grid_sample path.  Only tests/,
smoke() and bench.py's CPU-baseline legs may import this.

Reference lines restated:
  projections + joint softmax      ms_deform_attn.py:225-266  (_compute_deformable_attention)
  encoder per-frame loop           ms_deform_attn.py:419-464  (TemporalMSDeformAttnEncoder.forward)
  decoder per-frame loop           ms_deform_attn.py:297-414  (TemporalMSDeformAttnDecoder.forward)
  temporal frame table             devis_transformer.py:94-112,147-151

The reference ships no test for these modules ("parity unpinned" by the reference
itself); tests/golden/make_golden.py pins this restatement against the reference
modules run in the build container with their op routed to the PyTorch core.
"""
import torch
import torch.nn.functional as F

from .msda_torch import msda_forward_torch


def all_frames_offsets(n_frames, device=None):
    """devis_transformer.py:96-100 / :147-151 -- every other frame, in frame order."""
    return [torch.tensor([d for d in range(-t, n_frames - t) if d != 0], device=device)
            for t in range(n_frames)]


def window_offsets(n_frames, t_window, device=None):
    """devis_transformer.py:102-112 -- +-window/2 with reflection at the clip ends."""
    deltas = [d for d in range(-t_window // 2, t_window // 2 + 1) if d != 0]
    table = []
    for t in range(n_frames):
        table.append(torch.tensor([(-d if (t + d < 0 or t + d > n_frames - 1) else d)
                                   for d in deltas], device=device))
    return table


def temporal_level_start_index(temporal_shapes):
    """devis_transformer.py:118,154."""
    areas = temporal_shapes.prod(1)
    return torch.cat([areas.new_zeros(1), areas.cumsum(0)[:-1]])


def temporal_projections(sd, query, input_flatten, n_heads, n_levels, t_window, pc, pt):
    """ms_deform_attn.py:225-266.  Returns value (T,S,M,D), curr offsets
    (T,Lq,M,L,Pc,2), temporal offsets (T,Lq,M,Wt*L,Pt,2) [frame-major, level-minor],
    curr weights (T,Lq,M,L,Pc), temporal weights (T,Lq,M,Wt*L,Pt)."""
    t, lq, c = query.shape
    s = input_flatten.shape[1]
    value = F.linear(input_flatten, sd["value_proj.weight"], sd["value_proj.bias"])
    value = value.view(t, s, n_heads, c // n_heads)
    off_t = F.linear(query, sd["temporal_sampling_offsets.weight"], sd["temporal_sampling_offsets.bias"])
    off_t = off_t.view(t, lq, n_heads, t_window, n_levels, pt, 2).flatten(3, 4)
    logit_t = F.linear(query, sd["temporal_attention_weights.weight"], sd["temporal_attention_weights.bias"])
    logit_t = logit_t.view(t, lq, n_heads, t_window * n_levels * pt)
    logit_c = F.linear(query, sd["attention_weights.weight"], sd["attention_weights.bias"])
    logit_c = logit_c.view(t, lq, n_heads, n_levels * pc)
    joint = F.softmax(torch.cat([logit_c, logit_t], dim=3), -1)          # one softmax over all K taps
    aw_c = joint[..., :n_levels * pc].reshape(t, lq, n_heads, n_levels, pc).contiguous()
    aw_t = joint[..., n_levels * pc:].reshape(t, lq, n_heads, t_window * n_levels, pt).contiguous()
    off_c = F.linear(query, sd["sampling_offsets.weight"], sd["sampling_offsets.bias"])
    off_c = off_c.view(t, lq, n_heads, n_levels, pc, 2)
    return value, off_c, off_t, aw_c, aw_t


def temporal_core_per_frame(value, loc_curr, aw_curr, loc_temporal, aw_temporal, shapes,
                            temporal_offsets):
    """The reference's unit of work for one layer-clip: per query frame one op call on
    the frame's own value, one gather-copy of the other frames' values and one op
    call on that copy (ms_deform_attn.py:435-460).  value (T,S,M,D); loc_curr
    (T,Lq,M,L,P,2); aw_curr (T,Lq,M,L,P); loc_temporal (T,Lq,M,Wt*L,P,2); aw_temporal
    (T,Lq,M,Wt*L,P); returns (T,Lq,M*D)."""
    n_frames = value.shape[0]
    wt = temporal_offsets[0].numel()
    t_shapes = shapes.repeat(wt, 1)
    frames_out = []
    for t in range(n_frames):
        cur = msda_forward_torch(value[t][None], shapes, loc_curr[t][None], aw_curr[t][None])
        others = temporal_offsets[t] + t
        stacked = value[others].flatten(0, 1)[None]                      # the gather copy
        tmp = msda_forward_torch(stacked, t_shapes, loc_temporal[t][None], aw_temporal[t][None])
        frames_out.append(cur + tmp)
    return torch.cat(frames_out, dim=0)


def encoder_locations(reference_points, off_c, off_t, shapes, t_window):
    """ms_deform_attn.py:431-452: curr loc = ref (per level) + off/(W,H); temporal loc =
    level-0 reference point broadcast over every temporal level + off/(W,H)."""
    norm = torch.stack([shapes[:, 1], shapes[:, 0]], -1).to(off_c.dtype)        # (L,2) as (W,H)
    loc_c = reference_points[:, :, None, :, None, :] + off_c / norm[None, None, None, :, None, :]
    norm_t = norm.repeat(t_window, 1)
    loc_t = reference_points[:, :, 0][:, :, None, None, None, :] + off_t / norm_t[None, None, None, :, None, :]
    return loc_c, loc_t


def decoder_locations(reference_points, off_c, off_t, shapes, t_window, temporal_offsets,
                      pc, pt, instance_aware):
    """ms_deform_attn.py:320-404.  reference_points (T,q,L,2|4).  Instance-aware: the
    temporal reference of query i in frame t for temporal slot j is query i's own
    reference in frame temporal_frames[t][j] (:342-344,:383-385); otherwise its
    frame-t reference repeated (:346-347,:387-388)."""
    n_frames = reference_points.shape[0]
    rows = []
    for t in range(n_frames):
        others = temporal_offsets[t] + t
        if instance_aware:
            rows.append(reference_points[others].transpose(0, 1).flatten(1, 2))  # (q, Wt*L, 2|4)
        else:
            rows.append(reference_points[t].repeat(1, t_window, 1))
    ref_t = torch.stack(rows, 0)                                                  # (T,q,Wt*L,2|4)
    if reference_points.shape[-1] == 2:
        norm = torch.stack([shapes[:, 1], shapes[:, 0]], -1).to(off_c.dtype)
        loc_c = reference_points[:, :, None, :, None, :] + off_c / norm[None, None, None, :, None, :]
        loc_t = ref_t[:, :, None, :, None, :] + off_t / norm.repeat(t_window, 1)[None, None, None, :, None, :]
    elif reference_points.shape[-1] == 4:
        loc_c = reference_points[:, :, None, :, None, :2] \
            + (off_c / pc) * reference_points[:, :, None, :, None, 2:] * 0.5
        loc_t = ref_t[:, :, None, :, None, :2] + (off_t / pt) * ref_t[:, :, None, :, None, 2:] * 0.5
    else:
        raise ValueError("Last dim of reference_points must be 2 or 4")
    return loc_c, loc_t


def temporal_encoder_forward(sd, query, reference_points, input_flatten, shapes, temporal_offsets,
                             n_heads, n_levels, t_window, pc, pt):
    """TemporalMSDeformAttnEncoder.forward, ms_deform_attn.py:419-464 -> (T,S,C)."""
    value, off_c, off_t, aw_c, aw_t = temporal_projections(sd, query, input_flatten, n_heads,
                                                           n_levels, t_window, pc, pt)
    loc_c, loc_t = encoder_locations(reference_points, off_c, off_t, shapes, t_window)
    core = temporal_core_per_frame(value, loc_c, aw_c, loc_t, aw_t, shapes, temporal_offsets)
    return F.linear(core, sd["output_proj.weight"], sd["output_proj.bias"])


def temporal_decoder_forward(sd, query, reference_points, input_flatten, shapes, temporal_offsets,
                             n_heads, n_levels, t_window, pc, pt, instance_aware=True):
    """TemporalMSDeformAttnDecoder.forward, ms_deform_attn.py:297-414.  query (1,T*q,C);
    returns the 5-tuple (out (1,T*q,C), [T x loc_curr (1,q,M,L,Pc,2)],
    [T x loc_temporal (1,q,M,Wt*L,Pt,2)], aw_curr, aw_temporal)."""
    n_frames = input_flatten.shape[0]
    q = query.shape[1] // n_frames
    query = query.reshape(n_frames, q, query.shape[-1])
    if reference_points.shape[0] != n_frames:
        reference_points = reference_points.reshape((n_frames, q) + tuple(reference_points.shape[-2:]))
    value, off_c, off_t, aw_c, aw_t = temporal_projections(sd, query, input_flatten, n_heads,
                                                           n_levels, t_window, pc, pt)
    loc_c, loc_t = decoder_locations(reference_points, off_c, off_t, shapes, t_window,
                                     temporal_offsets, pc, pt, instance_aware)
    core = temporal_core_per_frame(value, loc_c, aw_c, loc_t, aw_t, shapes, temporal_offsets)
    out = F.linear(core.flatten(0, 1)[None], sd["output_proj.weight"], sd["output_proj.bias"])
    return (out, [loc_c[t][None] for t in range(n_frames)],
            [loc_t[t][None] for t in range(n_frames)], aw_c, aw_t)
