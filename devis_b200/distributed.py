"""Clip-level data parallelism, the only parallelism DeVIS has.

The reference wraps the model in DistributedDataParallel (main.py:131) with one clip per rank per step
(main.py:85) and a DistributedSampler over clips (main.py:141-147).  The attention op itself is intra-GPU, so the
op path needs no collective; the training configuration needs the parameter-gradient all-reduce.  These helpers are
backend-agnostic (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_clips(n_clips, rank, world_size, epoch=0, shuffle=False, seed=0):
    """Clip indices this rank processes: clip i -> rank i mod world (DistributedSampler semantics, padded by
    wrapping so every rank gets the same count -- main.py:141-147)."""
    order = list(range(n_clips))
    if shuffle:
        g = torch.Generator().manual_seed(seed + epoch)
        order = torch.randperm(n_clips, generator=g).tolist()
    per_rank = (n_clips + world_size - 1) // world_size
    total = per_rank * world_size
    order = (order * ((total + max(n_clips, 1) - 1) // max(n_clips, 1) + 1))[:total] if n_clips else []
    return order[rank:total:world_size]


def allreduce_gradients(parameters, world_size=None, bucket_bytes=25 * 1024 * 1024):
    """Average the gradients of `parameters` over the process group in flat buckets (what DDP's reducer does with its
    25 MB default buckets, main.py:131).  Parameters without a gradient contribute zeros
    (find_unused_parameters=True semantics)."""
    if not dist.is_available() or not dist.is_initialized():
        return 0
    world_size = world_size or dist.get_world_size()
    params = [p for p in parameters if p.requires_grad]
    buckets, cur, size = [], [], 0
    for p in params:
        n = p.numel() * p.element_size()
        if cur and size + n > bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
        cur.append(p)
        size += n
    if cur:
        buckets.append(cur)
    for bucket in buckets:
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world_size)
        off = 0
        for p in bucket:
            n = p.numel()
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(flat[off:off + n].view_as(p))
            off += n
    return len(buckets)


def max_over_ranks(value, device):
    """Timing reduction the benchmarks use: a step is as slow as its slowest rank."""
    if not dist.is_available() or not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
