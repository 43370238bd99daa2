"""world_size-2 gloo tests (CPU) of the clip-sharding / gradient all-reduce helpers used by the multi-GPU paths."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    from devis_b200 import distributed as d
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                       # same init on every rank, like DDP's broadcast
        lin = torch.nn.Linear(8, 4)
        unused = torch.nn.Parameter(torch.ones(3))
        clips = d.shard_clips(5, rank, world)
        x = torch.stack([torch.full((8,), float(c + 1)) for c in clips])
        lin(x).sum().backward()
        n_buckets = d.allreduce_gradients(list(lin.parameters()) + [unused], bucket_bytes=64)
        slow = d.max_over_ranks(10.0 + rank, "cpu")
        out.put((rank, clips, lin.weight.grad.clone(), unused.grad.clone(), n_buckets, slow))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_gradient_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([out.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, clips0, g0, u0, nb0, slow0), (r1, clips1, g1, u1, nb1, slow1) = res
    assert clips0 == [0, 2, 4] and clips1 == [1, 3, 0]           # clip i -> rank i mod 2, padded by wrapping
    assert torch.allclose(g0, g1)                                 # identical after the all-reduce
    # d(sum(Wx+b))/dW = sum of inputs per rank; average over the two ranks
    want = (torch.full((8,), float(1 + 3 + 5)) + torch.full((8,), float(2 + 4 + 1))) / 2
    assert torch.allclose(g0, want.expand(4, 8))
    assert torch.equal(u0, torch.zeros(3)) and torch.equal(u1, torch.zeros(3))   # unused parameter -> zero grad
    assert nb0 == nb1 and nb0 >= 2                                # 64-byte buckets force several buckets
    assert slow0 == slow1 == 11.0


def test_shard_clips_covers_every_clip_once_when_divisible():
    from devis_b200 import distributed as d
    got = sorted(sum((d.shard_clips(8, r, 4) for r in range(4)), []))
    assert got == list(range(8))
    assert d.shard_clips(0, 0, 2) == []
    a = d.shard_clips(7, 1, 2, epoch=3, shuffle=True, seed=5)
    b = d.shard_clips(7, 1, 2, epoch=3, shuffle=True, seed=5)
    assert a == b and len(a) == 4
