"""GPU baseline: the reference's original CUDA op (oracle/_ref, compiled unmodified for sm_100a) against this
repository's kernels, same process, same inputs, CUDA events (>= 20 warm-up, >= 100 timed iterations, median of
per-iteration events).  BASELINE.md section 4, item 2.

  (i)   per call, the four DeVIS call shapes of SURVEY.md section 3.4
  (ii)  the reference's whole layer-clip sequence: 12 calls + 6 gather copies (+ 6 adds), forward and backward
  (iii) the reference kernel in the single-call whole-clip form (24 "levels")
  ours: per-call drop-in and the whole-clip op

    python benchmarks/ref_cuda_baseline.py [--dist local] [--out profiles/r1_ref_cuda_baseline.json]
"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks.sweep import RawClip  # noqa: E402
from devis_b200 import MultiScaleDeformableAttention as ours, clip_geometry, synthetic  # noqa: E402
from oracle import ref_cuda_build  # noqa: E402


def med_us(fn, iters=100, warmup=20):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs) * 1e3


def lsi_of(shapes):
    areas = shapes.prod(1)
    return torch.cat([areas.new_zeros(1), areas.cumsum(0)[:-1]])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dist", default="local")
    ap.add_argument("--out", default="")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--only-batched", action="store_true", help="only the batched MSDeformAttn shape (section iv)")
    a = ap.parse_args()
    ref = ref_cuda_build.load()
    assert ref is not None, "oracle/_ref not built"
    res = {"dist": a.dist, "iters": a.iters, "gpu": torch.cuda.get_device_name(0)}

    def calls_for(clip, t):
        shapes = torch.tensor(clip["shapes"], device="cuda")
        wt = len(clip["frame_table"][t])
        tshapes = shapes.repeat(wt, 1)
        cur = (clip["value"][t][None].contiguous(), shapes, lsi_of(shapes), clip["loc_curr"][t][None].contiguous(),
               clip["aw_curr"][t][None].contiguous())
        tmp = (clip["value"][clip["frame_table"][t]].flatten(0, 1)[None].contiguous(), tshapes, lsi_of(tshapes),
               clip["loc_temporal"][t][None].contiguous(), clip["aw_temporal"][t][None].contiguous())
        return cur, tmp

    # (iv) the plain batched op of MSDeformAttn (ms_deform_attn.py:84-132): COCO pre-training encoder
    # shape, 800 x 1333 input -> levels /8 .. /64, batch 2, one query per pixel, 4 levels x 4 points
    def batched():
        coco = ((100, 167), (50, 84), (25, 42), (13, 21))
        bclip = synthetic.make_clip(n_frames=2, shapes=coco, dist=a.dist, seed=4, device="cuda", t_window=1)
        bshapes = torch.tensor(coco, device="cuda")
        args = (bclip["value"], bshapes, lsi_of(bshapes), bclip["loc_curr"], bclip["aw_curr"])
        gout = bclip["grad_out"]
        row = {"batch": 2, "S": int(bclip["value"].shape[1]), "Lq": int(bclip["loc_curr"].shape[1])}
        outs = {}
        for impl, mod in (("ref", ref), ("ours", ours)):
            row[impl + "_fwd_us"] = med_us(lambda: mod.ms_deform_attn_forward(*args, 64), a.iters)
            row[impl + "_bwd_us"] = med_us(lambda: mod.ms_deform_attn_backward(*args, gout, 64), a.iters)
            outs[impl] = mod.ms_deform_attn_forward(*args, 64)
        row["max_abs_diff"] = float((outs["ref"] - outs["ours"]).abs().max())
        print("batched MSDeformAttn", row, flush=True)
        return row

    if a.only_batched:
        res["batched_msdeformattn_coco_encoder"] = batched()
        if a.out:
            with open(a.out, "w") as fh:
                json.dump(res, fh, indent=1)
        return

    # (i) per call
    per_call = {}
    for name, lq in (("enc", None), ("dec10", 10), ("dec30", 30), ("dec300", 300)):
        clip = synthetic.make_clip(queries=lq, dist=a.dist, seed=1, device="cuda")
        cur, tmp = calls_for(clip, 2)
        gout = clip["grad_out"][2][None].contiguous()
        for kind, args in (("curr", cur), ("temporal", tmp)):
            row = {}
            for impl, mod in (("ref", ref), ("ours", ours)):
                row[impl + "_fwd_us"] = med_us(lambda: mod.ms_deform_attn_forward(*args, 64), a.iters)
                row[impl + "_bwd_us"] = med_us(lambda: mod.ms_deform_attn_backward(*args, gout, 64), a.iters)
            per_call[f"{name}_{kind}"] = row
            print(name, kind, row, flush=True)
    res["per_call"] = per_call

    # (ii) the reference's layer-clip sequence, encoder shape
    clip = synthetic.make_clip(dist=a.dist, seed=1, device="cuda")
    T = clip["value"].shape[0]
    shapes = torch.tensor(clip["shapes"], device="cuda")
    lsi = lsi_of(shapes)
    tshapes = shapes.repeat(T - 1, 1)
    tlsi = lsi_of(tshapes)
    lc = [clip["loc_curr"][t][None].contiguous() for t in range(T)]
    ac = [clip["aw_curr"][t][None].contiguous() for t in range(T)]
    lt = [clip["loc_temporal"][t][None].contiguous() for t in range(T)]
    at = [clip["aw_temporal"][t][None].contiguous() for t in range(T)]
    go = [clip["grad_out"][t][None].contiguous() for t in range(T)]
    idx = [torch.tensor(clip["frame_table"][t], device="cuda") for t in range(T)]

    def seq_fwd(mod):
        outs = []
        for t in range(T):
            cur = mod.ms_deform_attn_forward(clip["value"][t][None], shapes, lsi, lc[t], ac[t], 64)
            stacked = clip["value"][idx[t]].flatten(0, 1)[None]
            tmp = mod.ms_deform_attn_forward(stacked, tshapes, tlsi, lt[t], at[t], 64)
            outs.append(cur + tmp)
        return torch.cat(outs, 0)

    def seq_bwd(mod):
        gv = torch.zeros_like(clip["value"])
        for t in range(T):
            g1 = mod.ms_deform_attn_backward(clip["value"][t][None], shapes, lsi, lc[t], ac[t], go[t], 64)
            stacked = clip["value"][idx[t]].flatten(0, 1)[None]
            g2 = mod.ms_deform_attn_backward(stacked, tshapes, tlsi, lt[t], at[t], go[t], 64)
            gv[t] += g1[0][0]
            gv.index_add_(0, idx[t], g2[0][0].view(T - 1, *clip["value"].shape[1:]))   # autograd of value[temporal_frames]
        return gv

    seq = {}
    for impl, mod in (("ref", ref), ("ours_per_call", ours)):
        seq[impl + "_fwd_us"] = med_us(lambda: seq_fwd(mod), max(20, a.iters // 4), 5)
        seq[impl + "_bwd_us"] = med_us(lambda: seq_bwd(mod), max(20, a.iters // 4), 5)
    res["layer_clip_sequence_12_calls_6_copies"] = seq
    print("sequence", seq, flush=True)

    # (iii) single-call whole-clip form with the reference kernel: slot (f,l) = frame f, level l
    L, P = len(clip["shapes"]), clip["loc_curr"].shape[4]
    S, M = clip["value"].shape[1], clip["value"].shape[2]
    Lq = clip["loc_curr"].shape[1]
    big_shapes = shapes.repeat(T, 1)
    big_lsi = torch.cat([lsi + f * S for f in range(T)])
    loc = torch.zeros(1, T * Lq, M, T * L, P, 2, device="cuda")
    aw = torch.zeros(1, T * Lq, M, T * L, P, device="cuda")
    for t in range(T):
        q0 = slice(t * Lq, (t + 1) * Lq)
        loc[0, q0, :, t * L:(t + 1) * L] = clip["loc_curr"][t]
        aw[0, q0, :, t * L:(t + 1) * L] = clip["aw_curr"][t]
        for j, f in enumerate(clip["frame_table"][t]):
            loc[0, q0, :, f * L:(f + 1) * L] = clip["loc_temporal"][t][:, :, j * L:(j + 1) * L]
            aw[0, q0, :, f * L:(f + 1) * L] = clip["aw_temporal"][t][:, :, j * L:(j + 1) * L]
    big_value = clip["value"].reshape(1, T * S, M, -1)
    big_go = clip["grad_out"].reshape(1, T * Lq, -1)
    single = {}
    for impl, mod in (("ref", ref), ("ours_per_call", ours)):
        single[impl + "_fwd_us"] = med_us(lambda: mod.ms_deform_attn_forward(big_value, big_shapes, big_lsi, loc, aw, 64), max(20, a.iters // 4), 5)
        single[impl + "_bwd_us"] = med_us(lambda: mod.ms_deform_attn_backward(big_value, big_shapes, big_lsi, loc, aw, big_go, 64), max(20, a.iters // 4), 5)
    res["single_call_whole_clip_form"] = single
    print("single-call", single, flush=True)
    want = ref.ms_deform_attn_forward(big_value, big_shapes, big_lsi, loc, aw, 64).view(T, Lq, -1)
    del loc, aw

    # ours: whole-clip op
    geom = clip_geometry.ClipGeometry(clip["shapes"], T, clip["frame_table"])
    rc = RawClip(clip, geom.tile_order("cuda"))
    res["ours_whole_clip"] = {"fwd_us": med_us(rc.fwd, a.iters), "bwd_us": med_us(rc.bwd, a.iters)}
    rc.fwd()
    res["ours_whole_clip"]["max_abs_diff_vs_ref_single_call"] = float((rc.out - want).abs().max())
    print("ours whole-clip", res["ours_whole_clip"], flush=True)
    # decoder layer-clips (launch-bound regime): the reference's 12-call sequence vs ONE whole-clip launch
    dec = {}
    for lq in (10, 30, 300):
        dclip = synthetic.make_clip(queries=lq, dist=a.dist, seed=2, device="cuda")
        d_lc = [dclip["loc_curr"][t][None].contiguous() for t in range(T)]
        d_ac = [dclip["aw_curr"][t][None].contiguous() for t in range(T)]
        d_lt = [dclip["loc_temporal"][t][None].contiguous() for t in range(T)]
        d_at = [dclip["aw_temporal"][t][None].contiguous() for t in range(T)]

        def dec_seq_fwd(mod):
            outs = []
            for t in range(T):
                cur = mod.ms_deform_attn_forward(dclip["value"][t][None], shapes, lsi, d_lc[t], d_ac[t], 64)
                stacked = dclip["value"][idx[t]].flatten(0, 1)[None]
                outs.append(cur + mod.ms_deform_attn_forward(stacked, tshapes, tlsi, d_lt[t], d_at[t], 64))
            return torch.cat(outs, 0)

        drc = RawClip(dclip, None)
        dec[f"q{lq}"] = {"ref_sequence_fwd_us": med_us(lambda: dec_seq_fwd(ref), max(20, a.iters // 4), 5),
                         "ours_whole_clip_fwd_us": med_us(drc.fwd, a.iters), "ours_whole_clip_bwd_us": med_us(drc.bwd, a.iters)}
        print("decoder", lq, dec[f"q{lq}"], flush=True)
    res["decoder_layer_clip"] = dec
    res["batched_msdeformattn_coco_encoder"] = batched()
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
