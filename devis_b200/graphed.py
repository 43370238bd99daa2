"""CUDA-graph replay of a layer for fixed shapes.

At decoder sizes (10 - 300 object queries per frame) the temporal attention kernels take tens of microseconds while
the layer around them is a few dozen small launches: the layer is bound by launch latency, not by the GPU
(reference call site: the decoder's cross-attention, deformable_transformer.py:263, inside
DeVISTransformerDecoderLayer.forward).  ``GraphedLayer`` captures the forward AND the backward of a module once, for
the shapes of a sample call, and replays them afterwards: one graph launch each way instead of ~100 kernel launches.

    layer = DeVISTransformerDecoderLayer(...).cuda()
    fast = GraphedLayer(layer, tgt, query_pos, reference_points, src, (shapes, tshapes), (lsi, tlsi),
                        temporal_offsets=offsets)
    out = fast(tgt, query_pos, reference_points, src, (shapes, tshapes), (lsi, tlsi), temporal_offsets=offsets)
    out.sum().backward()                       # replays the captured backward; parameter .grad is accumulated as usual

Floating-point tensors passed positionally are the graph's INPUTS (their values may change from call to call; shapes,
dtypes and requires_grad must stay those of the sample).  Everything else -- the (current, temporal) pairs of shape
and start-index tensors, the temporal offset lists, keyword arguments -- is STATIC: bound at capture time, and a call
that passes something different is refused rather than silently replayed with the captured value.

Built on ``torch.cuda.make_graphed_callables`` (capture of forward and backward into one memory pool, autograd
integration); what this module adds is the argument split, the static-argument check and support for the decoder
module's 5-tuple (lists of per-frame location tensors).  The kernels of this package never synchronise and read no
device tensor back once the clip geometry is memoised (clip_geometry.from_reference_args), which is what makes the
capture possible; the warm-up calls that ``make_graphed_callables`` issues before capturing do that memoisation.
"""
import gc

import torch
from torch import nn


def _is_dynamic(x):
    return isinstance(x, torch.Tensor) and x.is_floating_point()


def _same_static(a, b):
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        return a is b or (a.shape == b.shape and a.dtype == b.dtype and a.device == b.device and a.data_ptr() == b.data_ptr())
    if isinstance(a, (tuple, list)) and isinstance(b, (tuple, list)):
        return len(a) == len(b) and all(_same_static(x, y) for x, y in zip(a, b))
    return a is b or a == b


class _Flat(nn.Module):
    """the wrapped module with its static arguments bound; returns a flat tuple of tensors"""

    def __init__(self, module, template, static_kwargs):
        super().__init__()
        self.module = module
        self.template = template          # positional args with None where a dynamic tensor goes
        self.static_kwargs = static_kwargs
        self.out_spec = None

    def forward(self, *dynamic):
        it = iter(dynamic)
        args = [next(it) if slot is None else slot for slot in self.template]
        out = self.module(*args, **self.static_kwargs)
        flat, spec = [], []
        for o in (out if isinstance(out, tuple) else (out,)):
            if isinstance(o, torch.Tensor):
                flat.append(o)
                spec.append(("t", 1))
            elif isinstance(o, (list, tuple)) and o and all(isinstance(x, torch.Tensor) for x in o):
                flat.extend(o)
                spec.append(("l", len(o)))
            else:
                spec.append(("c", o))       # constants (e.g. the None of the encoder module's (out, None))
        self.out_spec = (spec, isinstance(out, tuple))
        return tuple(flat)


class GraphedLayer(nn.Module):
    def __init__(self, module, *sample_args, num_warmup_iters=3, **static_kwargs):
        super().__init__()
        if not any(_is_dynamic(a) for a in sample_args):
            raise ValueError("GraphedLayer needs at least one floating-point tensor among the positional arguments")
        dev = next(a for a in sample_args if _is_dynamic(a)).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedLayer captures CUDA graphs: the sample arguments must be CUDA tensors")
        self.template = [None if _is_dynamic(a) else a for a in sample_args]
        self.static_kwargs = dict(static_kwargs)
        self.sample_meta = [(tuple(a.shape), a.dtype, a.requires_grad) for a in sample_args if _is_dynamic(a)]
        self.flat = _Flat(module, self.template, self.static_kwargs)
        samples = tuple(a.detach().clone().requires_grad_(a.requires_grad) for a in sample_args if _is_dynamic(a))
        # make_graphed_callables swaps the module's forward for the graphed one and returns the module
        # No cyclic garbage collection while the graphs are captured: a collection that happens to run mid-capture
        # and frees an object owning CUDA resources (an older CUDA graph and its memory pool, for one) issues calls
        # that invalidate every capture in progress ("operation failed due to a previous error during capture").
        gc.collect()
        was_enabled = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.device(dev):
                self.graphed = torch.cuda.make_graphed_callables(self.flat, samples, num_warmup_iters=num_warmup_iters,
                                                                 allow_unused_input=True)
        finally:
            if was_enabled:
                gc.enable()

    @property
    def module(self):
        return self.flat.module

    def forward(self, *args, **kwargs):
        if len(args) != len(self.template):
            raise RuntimeError(f"GraphedLayer was captured with {len(self.template)} positional arguments, got {len(args)}")
        # (a train/eval mode other than the one at capture runs the module eagerly: torch's graphed forward checks it)
        dynamic = []
        for slot, a in zip(self.template, args):
            if slot is None:
                if not _is_dynamic(a):
                    raise RuntimeError("GraphedLayer: a graph input must be a floating-point tensor")
                dynamic.append(a)
            elif not _same_static(slot, a):
                raise RuntimeError("GraphedLayer: a static argument differs from the one bound at capture "
                                   "(shapes / start indices / temporal offsets are part of the captured graph)")
        if set(kwargs) != set(self.static_kwargs) or not all(_same_static(self.static_kwargs[k], v) for k, v in kwargs.items()):
            raise RuntimeError("GraphedLayer: keyword arguments differ from the ones bound at capture")
        for a, (shape, dtype, _) in zip(dynamic, self.sample_meta):
            if tuple(a.shape) != shape or a.dtype != dtype:
                raise RuntimeError(f"GraphedLayer: input {tuple(a.shape)} {a.dtype} does not match the captured "
                                   f"{shape} {dtype}")
        flat = self.graphed(*dynamic)
        flat = flat if isinstance(flat, tuple) else (flat,)
        spec, was_tuple = self.flat.out_spec
        out, i = [], 0
        for kind, n in spec:
            if kind == "t":
                out.append(flat[i])
                i += 1
            elif kind == "l":
                out.append(list(flat[i:i + n]))
                i += n
            else:
                out.append(n)
        return tuple(out) if was_tuple else out[0]
