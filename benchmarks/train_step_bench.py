"""Training-step benchmark of the DeVIS transformer trunk (BASELINE.json configs[4] without backbone / mask head /
criterion, which are out of scope): 6 temporal-deformable encoder layers + 6 decoder layers on synthetic R50 T=6
features, forward + backward + gradient all-reduce (DDP over NCCL, one clip per rank like main.py:85,131) +
clip_grad_norm_(0.1) + AdamW -- the body of train_one_epoch (engine.py:48-77).

    python benchmarks/train_step_bench.py [--steps 10] [--attn ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/train_step_bench.py

The layer wrappers below restate deformable_transformer.py:143-173 (encoder layer) and :215-281 (decoder layer) as
benchmark scaffolding; the attention modules are this repository's.  `--attn reference` swaps the temporal attention
for the reference's per-frame loop around the reference's own CUDA op (oracle/_ref), single GPU only.
"""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devis_b200 import synthetic  # noqa: E402
from devis_b200.modules import TemporalMSDeformAttnDecoder, TemporalMSDeformAttnEncoder  # noqa: E402


class RefLoopEncoderAttn(TemporalMSDeformAttnEncoder):
    """same parameters, evaluated the reference's way: per-frame loop + gather copies + the reference CUDA op"""

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                temporal_offsets):
        from benchmarks.module_bench import RefFunction, reference_style_encoder
        return reference_style_encoder(self, RefFunction.apply, query, reference_points, input_flatten,
                                       input_spatial_shapes[0], input_level_start_index[0], input_spatial_shapes[1],
                                       input_level_start_index[1], temporal_offsets), None


class EncoderLayer(nn.Module):
    def __init__(self, attn_cls, t, d=256, ffn=1024, drop=0.1):
        super().__init__()
        self.self_attn = attn_cls(t, d, 4, t - 1, 8, 4, 4)
        self.drop1, self.norm1 = nn.Dropout(drop), nn.LayerNorm(d)
        self.lin1, self.lin2 = nn.Linear(d, ffn), nn.Linear(ffn, d)
        self.drop2, self.drop3, self.norm2 = nn.Dropout(drop), nn.Dropout(drop), nn.LayerNorm(d)

    def forward(self, src, pos, ref, shapes, lsi, offsets):
        a, _ = self.self_attn(src + pos, ref, src, shapes, lsi, offsets)
        src = self.norm1(src + self.drop1(a))
        f = self.lin2(self.drop2(torch.relu(self.lin1(src))))
        return self.norm2(src + self.drop3(f))


class DecoderLayer(nn.Module):
    def __init__(self, t, d=256, ffn=1024, drop=0.1):
        super().__init__()
        self.cross_attn = TemporalMSDeformAttnDecoder(t, d, 4, t - 1, 8, 4, 4, True)
        self.self_attn = nn.MultiheadAttention(d, 8, dropout=drop)
        self.n1, self.n2, self.n3 = nn.LayerNorm(d), nn.LayerNorm(d), nn.LayerNorm(d)
        self.d1, self.d2, self.d3, self.d4 = (nn.Dropout(drop) for _ in range(4))
        self.lin1, self.lin2 = nn.Linear(d, ffn), nn.Linear(ffn, d)

    def forward(self, tgt, qpos, ref, memory, shapes, lsi, offsets):
        q = k = tgt + qpos
        sa = self.self_attn(q.transpose(0, 1), k.transpose(0, 1), tgt.transpose(0, 1))[0].transpose(0, 1)
        tgt = self.n2(tgt + self.d2(sa))
        ca = self.cross_attn(tgt + qpos, ref, memory, shapes, lsi, offsets)[0]
        tgt = self.n1(tgt + self.d1(ca))
        f = self.lin2(self.d3(torch.relu(self.lin1(tgt))))
        return self.n3(tgt + self.d4(f))


class Trunk(nn.Module):
    def __init__(self, t, queries_per_frame, attn_cls, layers=6):
        super().__init__()
        self.enc = nn.ModuleList([EncoderLayer(attn_cls, t) for _ in range(layers)])
        self.dec = nn.ModuleList([DecoderLayer(t) for _ in range(layers)])
        self.query_embed = nn.Embedding(t * queries_per_frame, 512)
        self.ref_proj = nn.Linear(256, 2)

    def forward(self, src, pos, enc_ref, shapes, lsi, offsets):
        mem = src
        for layer in self.enc:
            mem = layer(mem, pos, enc_ref, shapes, lsi, offsets)
        qpos, tgt = torch.split(self.query_embed.weight, 256, dim=1)
        qpos, tgt = qpos[None], tgt[None]
        ref = self.ref_proj(qpos).sigmoid()[:, :, None].expand(-1, -1, 4, -1)      # (1, T*q, L, 2), valid ratio 1
        for layer in self.dec:
            tgt = layer(tgt, qpos, ref, mem, shapes, lsi, offsets)
        return tgt, mem


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--attn", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=10)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if a.attn == "reference":
        from benchmarks.module_bench import RefFunction
        from oracle import ref_cuda_build
        RefFunction.mod = ref_cuda_build.load()
        assert RefFunction.mod is not None and world == 1
    T, shapes_l = 6, synthetic.DEVIS_SHAPES
    S = sum(h * w for h, w in shapes_l)
    torch.manual_seed(0)
    model = Trunk(T, a.queries, TemporalMSDeformAttnEncoder if a.attn == "ours" else RefLoopEncoderAttn).to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4, weight_decay=1e-4)
    g = torch.Generator(device=dev).manual_seed(1 + rank)          # every rank its own clip
    src = torch.randn(T, S, 256, device=dev, generator=g)
    pos = torch.randn(T, S, 256, device=dev, generator=g)
    enc_ref = synthetic.pixel_reference_points(shapes_l, T, dev)
    shapes = torch.tensor(shapes_l, device=dev)
    lsi = torch.tensor(synthetic.level_start_index(shapes_l), device=dev)
    tshapes = shapes.repeat(T - 1, 1)
    tlsi = torch.cat([tshapes.new_zeros(1), tshapes.prod(1).cumsum(0)[:-1]])
    offsets = [torch.tensor([d for d in range(-t, T - t) if d != 0], device=dev) for t in range(T)]

    def step():
        opt.zero_grad(set_to_none=True)
        tgt, mem = model(src, pos, enc_ref, (shapes, tshapes), (lsi, tlsi), offsets)
        loss = tgt.square().mean() + 1e-3 * mem.square().mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    times = []
    for _ in range(a.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = step()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = statistics.median(times)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        res = {"workload": "DeVIS transformer trunk training step: 6 enc + 6 dec layers, T=6, S=4820, 10 queries/frame, "
                           "fwd+bwd+grad all-reduce+clip+AdamW, fp32, synthetic features",
               "attention": a.attn, "n_gpus": world, "ms_per_step": ms, "clips_per_sec": world / (ms * 1e-3),
               "params": n_params, "grad_allreduce_bytes_per_step": n_params * 4 if world > 1 else 0,
               "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
        print(json.dumps(res), flush=True)
        if a.out:
            with open(a.out, "w") as fh:
                json.dump(res, fh, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
