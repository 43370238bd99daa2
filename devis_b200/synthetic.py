"""Synthetic workloads at the shapes DeVIS runs: used by bench.py and tests.

R50 / Swin-L at 360x640: 4 levels (45,80),(23,40),(12,20),(6,10) -> S = 4820 rows per frame, T = 6
frames, 8 heads x 32 channels, 4 current + 4 temporal points, t_window = T-1 = 5 -> K = 96 taps per
(query, head).  Encoder: one query per pixel (Lq = S); decoder: 10 / 30 / 300 object queries per frame.
"""
import math

import torch

DEVIS_SHAPES = ((45, 80), (23, 40), (12, 20), (6, 10))


def level_start_index(shapes):
    out, acc = [], 0
    for h, w in shapes:
        out.append(acc)
        acc += h * w
    return out


def pixel_reference_points(shapes, n_frames, device, dtype=torch.float32):
    """Encoder reference points with valid ratio 1: pixel centres of every level in normalised
    coordinates, same point repeated for all levels (deformable_transformer.py:184-198) -> (T,S,L,2)."""
    pts = []
    for h, w in shapes:
        ys = (torch.arange(h, device=device, dtype=dtype) + 0.5) / h
        xs = (torch.arange(w, device=device, dtype=dtype) + 0.5) / w
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack([gx.reshape(-1), gy.reshape(-1)], -1))
    ref = torch.cat(pts, 0)                                  # (S,2)
    return ref[None, :, None, :].expand(n_frames, -1, len(shapes), -1).contiguous()


def make_boundary_safe(loc, sizes_wh, margin=0.02):
    """Move every tap so that its pixel coordinate loc*size-0.5 keeps `margin` px from cell borders:
    floor() then picks the same cell in every arithmetic (CUDA fp32, grid_sample, fp64) -- SURVEY.md
    section 7 'floor discontinuity'.  loc (...,L,P,2), sizes_wh (L,2)."""
    size = sizes_wh.to(loc.dtype)[:, None, :]
    pix = loc * size - 0.5
    cell = torch.floor(pix)
    frac = (pix - cell).clamp(margin, 1 - margin)
    return (cell + frac + 0.5) / size


def _head_dirs(m, device):
    th = torch.arange(m, device=device, dtype=torch.float32) * (2 * math.pi / m)
    d = torch.stack([th.cos(), th.sin()], -1)
    return d / d.abs().max(-1, keepdim=True)[0]


def make_clip(n_frames=6, shapes=DEVIS_SHAPES, heads=8, channels=32, pc=4, pt=4, queries=None, dist="local",
              dtype=torch.float32, device="cuda", seed=0, sigma_px=2.0, t_window=None):
    """One layer-clip of op inputs.  queries=None -> encoder (one query per pixel); else decoder with
    `queries` object queries per frame.  dist: 'local' = DeVIS-like (reference point + head ray
    (i+1) px + N(0, sigma_px) px, the module's initial pattern ms_deform_attn.py:189-212 plus learned
    jitter) or 'uniform' = U(0,1) like the reference test (test.py:32).  All taps boundary-safe.
    Returns dict with value, loc_curr, aw_curr, loc_temporal, aw_temporal, grad_out, frame_table."""
    g = torch.Generator(device=device).manual_seed(seed)
    nl = len(shapes)
    s = sum(h * w for h, w in shapes)
    wt = (n_frames - 1) if t_window is None else t_window
    sizes = torch.tensor([[w, h] for h, w in shapes], device=device, dtype=torch.float32)      # (L,2) (W,H)
    if queries is None:
        lq = s
        ref = pixel_reference_points(shapes, n_frames, device)                                  # (T,S,L,2)
    else:
        lq = queries
        ref = torch.sigmoid(torch.randn(n_frames, lq, 1, 2, generator=g, device=device)).expand(-1, -1, nl, -1)

    def taps(n_slots, p, sizes_slots, ref_slots):
        if dist == "uniform":
            loc = torch.rand(n_frames, lq, heads, n_slots, p, 2, generator=g, device=device)
        else:
            ray = _head_dirs(heads, device).view(1, 1, heads, 1, 1, 2) \
                * torch.arange(1, p + 1, device=device, dtype=torch.float32).view(1, 1, 1, 1, p, 1)
            off = ray + sigma_px * torch.randn(n_frames, lq, heads, n_slots, p, 2, generator=g, device=device)
            loc = ref_slots[:, :, None, :, None, :] + off / sizes_slots[None, None, None, :, None, :]
        return make_boundary_safe(loc, sizes_slots)

    loc_c = taps(nl, pc, sizes, ref)
    ref_t = ref[:, :, :1].expand(-1, -1, wt * nl, -1)       # level-0 reference on every temporal level (:447)
    loc_t = taps(wt * nl, pt, sizes.repeat(wt, 1), ref_t)
    logits = torch.randn(n_frames, lq, heads, nl * pc + wt * nl * pt, generator=g, device=device)
    aw = torch.softmax(logits, -1)
    out = {
        "value": torch.randn(n_frames, s, heads, channels, generator=g, device=device).to(dtype),
        "loc_curr": loc_c.contiguous(),
        "aw_curr": aw[..., :nl * pc].reshape(n_frames, lq, heads, nl, pc).contiguous(),
        "loc_temporal": loc_t.contiguous(),
        "aw_temporal": aw[..., nl * pc:].reshape(n_frames, lq, heads, wt * nl, pt).contiguous(),
        "grad_out": torch.randn(n_frames, lq, heads * channels, generator=g, device=device).to(dtype),
        "shapes": shapes,
        "frame_table": [[f for f in range(n_frames) if f != t][:wt] for t in range(n_frames)],
    }
    if dtype == torch.float64:
        for k in ("loc_curr", "aw_curr", "loc_temporal", "aw_temporal"):
            out[k] = out[k].double()
    return out


def algorithmic_bytes(n_frames, s, heads, channels, lq, k_taps, elem=4, aux=4):
    """every distinct input element read once, every output element written once.
    elem = bytes of value/out/grad_out/grad_value elements, aux = bytes of location / weight elements."""
    c = heads * channels
    value = n_frames * s * c * elem
    out = n_frames * lq * c * elem
    loc = n_frames * lq * heads * k_taps * 2 * aux
    aw = n_frames * lq * heads * k_taps * aux
    fwd = value + out + loc + aw
    bwd = (value + loc + aw + out) + (value + loc + aw)
    return fwd, bwd
