"""Module-level timing at the DeVIS R50 T=6 shape: TemporalMSDeformAttnEncoder forward+backward (projections,
softmax, location arithmetic, the op, output projection), ours vs the reference's per-frame loop driven with the
reference's own CUDA op (oracle/_ref) -- i.e. the layer a DeVIS user actually runs.

    python benchmarks/module_bench.py [--profile] [--out file.json]
"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devis_b200 import synthetic  # noqa: E402
from devis_b200.modules import TemporalMSDeformAttnDecoder, TemporalMSDeformAttnEncoder  # noqa: E402


def med_ms(fn, iters=30, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs)


class RefFunction(torch.autograd.Function):
    """the reference's MSDeformAttnFunction (ms_deform_attn_func.py:21-38) bound to ITS compiled op"""
    mod = None

    @staticmethod
    def forward(ctx, value, shapes, lsi, loc, aw, step):
        ctx.step = step
        ctx.save_for_backward(value, shapes, lsi, loc, aw)
        return RefFunction.mod.ms_deform_attn_forward(value, shapes, lsi, loc, aw, step)

    @staticmethod
    def backward(ctx, g):
        value, shapes, lsi, loc, aw = ctx.saved_tensors
        gv, gl, ga = RefFunction.mod.ms_deform_attn_backward(value, shapes, lsi, loc, aw, g.contiguous(), ctx.step)
        return gv, None, None, gl, ga, None


def reference_style_encoder(mod, fn, query, ref, inp, shapes, lsi, tshapes, tlsi, offsets):
    """ms_deform_attn.py:419-464 restated around an MSDeformAttnFunction-like `fn` (bench infrastructure)."""
    value, off_c, off_t, aw_c, aw_t = mod._compute_deformable_attention(query, inp)
    aw_c, aw_t = aw_c.contiguous(), aw_t.contiguous()
    norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
    tnorm = norm.repeat(mod.t_window, 1)
    outs = []
    for t in range(query.shape[0]):
        loc = ref[t][None, :, None, :, None] + off_c[t][None] / norm[None, None, None, :, None, :]
        cur = fn(value[t][None], shapes, lsi, loc, aw_c[t][None], 64)
        frames = offsets[t] + t if isinstance(offsets[t], torch.Tensor) else [o + t for o in offsets[t]]
        stacked = value[frames].flatten(0, 1)[None]
        tref = ref[t, :, 0][None, :, None, None, None]
        tloc = tref + off_t[t][None] / tnorm[None, None, None, :, None, :]
        tmp = fn(stacked, tshapes, tlsi, tloc, aw_t[t][None], 64)
        outs.append(cur + tmp)
    return mod.output_proj(torch.cat(outs, 0))


def reference_style_decoder(mod, fn, query, ref, inp, shapes, lsi, tshapes, tlsi, offsets):
    """The call structure of the reference decoder layer (ms_deform_attn.py:299-414: per frame one current call, one
    gather copy of the other frames' value, one temporal call; instance-aware temporal reference points; 2-d or box
    reference points) around an MSDeformAttnFunction-like `fn` (bench infrastructure)."""
    n_frames = inp.shape[0]
    q = query.shape[1] // n_frames
    query = query.reshape(n_frames, q, query.shape[-1])
    ref = ref.reshape((n_frames, q) + tuple(ref.shape[-2:]))
    value, off_c, off_t, aw_c, aw_t = mod._compute_deformable_attention(query, inp)
    aw_c, aw_t = aw_c.contiguous(), aw_t.contiguous()
    norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
    tnorm = norm.repeat(mod.t_window, 1)
    outs = []
    for t in range(n_frames):
        frames = offsets[t] + t
        tref = ref[frames].transpose(0, 1).flatten(1, 2)[None, :, None, :, None]
        if ref.shape[-1] == 2:
            loc = ref[t][None, :, None, :, None] + off_c[t][None] / norm[None, None, None, :, None, :]
            tloc = tref + off_t[t][None] / tnorm[None, None, None, :, None, :]
        else:
            loc = ref[t][None, :, None, :, None, :2] + (off_c[t][None] / mod.n_curr_points) * ref[t][None, :, None, :, None, 2:] * 0.5
            tloc = tref[..., :2] + (off_t[t][None] / mod.n_temporal_points) * tref[..., 2:] * 0.5
        cur = fn(value[t][None], shapes, lsi, loc, aw_c[t][None], 64)
        stacked = value[frames].flatten(0, 1)[None]
        tmp = fn(stacked, tshapes, tlsi, tloc, aw_t[t][None], 64)
        outs.append(cur + tmp)
    return mod.output_proj(torch.cat(outs, 0).flatten(0, 1)[None])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from devis_b200 import MSDeformAttnFunction
    from oracle import ref_cuda_build
    RefFunction.mod = ref_cuda_build.load()
    dev = "cuda"
    T, shapes_l = 6, synthetic.DEVIS_SHAPES
    S = sum(h * w for h, w in shapes_l)
    torch.manual_seed(0)
    enc = TemporalMSDeformAttnEncoder(n_frames=T, d_model=256, n_levels=4, t_window=T - 1, n_heads=8, n_curr_points=4,
                                      n_temporal_points=4).to(dev)
    with torch.no_grad():     # non-trivial offsets / logits like a trained layer
        for lin in (enc.sampling_offsets, enc.temporal_sampling_offsets, enc.attention_weights, enc.temporal_attention_weights):
            lin.weight.normal_(0, 0.02)
    query = torch.randn(T, S, 256, device=dev, requires_grad=True)
    inp = torch.randn(T, S, 256, device=dev, requires_grad=True)
    ref = synthetic.pixel_reference_points(shapes_l, T, dev)
    shapes = torch.tensor(shapes_l, device=dev)
    lsi = torch.tensor(synthetic.level_start_index(shapes_l), device=dev)
    tshapes = shapes.repeat(T - 1, 1)
    tlsi = torch.cat([tshapes.new_zeros(1), tshapes.prod(1).cumsum(0)[:-1]])
    offsets = [torch.tensor([d for d in range(-t, T - t) if d != 0], device=dev) for t in range(T)]
    gout = torch.randn(T, S, 256, device=dev)

    def ours():
        out, _ = enc(query, ref, inp, (shapes, tshapes), (lsi, tlsi), offsets)
        out.backward(gout)

    def ours_fwd():
        with torch.no_grad():
            enc(query, ref, inp, (shapes, tshapes), (lsi, tlsi), offsets)

    def refstyle(fn):
        def run():
            out = reference_style_encoder(enc, fn, query, ref, inp, shapes, lsi, tshapes, tlsi, offsets)
            out.backward(gout)
        return run

    def refstyle_fwd(fn):
        def run():
            with torch.no_grad():
                reference_style_encoder(enc, fn, query, ref, inp, shapes, lsi, tshapes, tlsi, offsets)
        return run

    key = "encoder_layer_attention_T6_S4820"
    res = {key: {
        "ours_fwd_bwd_ms": med_ms(ours), "ours_fwd_ms": med_ms(ours_fwd),
        "per_frame_loop_with_our_dropin_op_fwd_bwd_ms": med_ms(refstyle(MSDeformAttnFunction.apply)),
    }}
    if RefFunction.mod is not None:
        res[key]["reference_loop_with_reference_cuda_op_fwd_bwd_ms"] = med_ms(refstyle(RefFunction.apply))
        res[key]["reference_loop_with_reference_cuda_op_fwd_ms"] = med_ms(refstyle_fwd(RefFunction.apply))
    print(json.dumps(res, indent=1), flush=True)

    # ---- decoder cross-attention layer (A7): q object queries per frame, 2-d reference points (layer 0) and boxes
    # (layers 1-5 under box refinement); eager and CUDA-graph replay of forward+backward
    for q in (10, 30, 300):
        for ref_dim in (2, 4):
            torch.manual_seed(1)
            dec = TemporalMSDeformAttnDecoder(n_frames=T, d_model=256, n_levels=4, t_window=T - 1, n_heads=8,
                                              n_curr_points=4, n_temporal_points=4).to(dev)
            with torch.no_grad():
                for lin in (dec.sampling_offsets, dec.temporal_sampling_offsets, dec.attention_weights,
                            dec.temporal_attention_weights):
                    lin.weight.normal_(0, 0.02)
            dq = torch.randn(1, T * q, 256, device=dev, requires_grad=True)
            dref = torch.rand(1, T * q, 4, ref_dim, device=dev) * 0.6 + 0.2
            if ref_dim == 4:
                dref[..., 2:] = dref[..., 2:] * 0.3
            dgout = torch.randn(1, T * q, 256, device=dev)
            args = (dq, dref, inp, (shapes, tshapes), (lsi, tlsi), offsets)

            def d_ours():
                dec(*args)[0].backward(dgout)

            def d_ours_fwd():
                with torch.no_grad():
                    dec(*args)

            def d_ref(fn, bwd=True):
                def run():
                    if bwd:
                        reference_style_decoder(dec, fn, dq, dref, inp, shapes, lsi, tshapes, tlsi, offsets).backward(dgout)
                    else:
                        with torch.no_grad():
                            reference_style_decoder(dec, fn, dq, dref, inp, shapes, lsi, tshapes, tlsi, offsets)
                return run

            key = f"decoder_layer_attention_T6_q{q}_ref{ref_dim}d"
            row = {"ours_fwd_bwd_ms": med_ms(d_ours), "ours_fwd_ms": med_ms(d_ours_fwd)}
            dec.fuse_prologue = False
            row["ours_unfused_prologue_fwd_bwd_ms"] = med_ms(d_ours)
            dec.fuse_prologue = True
            row["per_frame_loop_with_our_dropin_op_fwd_bwd_ms"] = med_ms(d_ref(MSDeformAttnFunction.apply))
            # product path for fixed shapes: devis_b200.GraphedLayer (forward and backward graphs, autograd-integrated)
            try:
                from devis_b200 import GraphedLayer
                import copy
                gl = GraphedLayer(copy.deepcopy(dec), *args)
                g_q = dq.detach().clone().requires_grad_(True)

                def d_graphed():
                    gl(g_q, dref, inp, (shapes, tshapes), (lsi, tlsi), offsets)[0].backward(dgout)

                row["ours_graphed_layer_fwd_bwd_ms"] = med_ms(d_graphed)
                del gl
            except Exception as exc:   # noqa: BLE001
                row["ours_graphed_layer_fwd_bwd_ms"] = f"failed: {str(exc)[:160]}"
            if RefFunction.mod is not None:
                row["reference_loop_with_reference_cuda_op_fwd_bwd_ms"] = med_ms(d_ref(RefFunction.apply))
                row["reference_loop_with_reference_cuda_op_fwd_ms"] = med_ms(d_ref(RefFunction.apply, False))
            # CUDA-graph replay of our forward+backward (the layer is launch-bound at these sizes)
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(3):
                        d_ours()
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                for prm in list(dec.parameters()) + [dq, inp]:
                    prm.grad = None
                with torch.cuda.graph(graph):
                    d_ours()
                row["ours_fwd_bwd_cuda_graph_ms"] = med_ms(graph.replay)
            except Exception as exc:   # noqa: BLE001
                row["ours_fwd_bwd_cuda_graph_ms"] = f"capture failed: {str(exc)[:120]}"
            res[key] = row
            print(json.dumps({key: row}), flush=True)

    if a.profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(3):
                ours()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
