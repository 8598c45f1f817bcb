"""Data-parallel gradient exchange for the HAMT hot path: one process per GPU, NCCL over NVLink.

The reference wraps the model in ``DistributedDataParallel(find_unused_parameters=True)``
(pretrain_src/utils/misc.py:52-65): per step it walks the autograd graph for unused parameters and
all-reduces ~27 buckets of 25 MiB.  Here the gradients already live in ONE flat fp32 buffer (arena.py)
in layer order, every rank runs the same task in a step (same touched set), so the exchange is a
handful of ``all_reduce(AVG)`` calls on contiguous slices:

  * ``sync_grads``      -- after backward: coalesce the touched parameters into contiguous ranges and
                           all-reduce each (async on the NCCL stream, one wait at the end);
  * ``LayerOverlap``    -- optional: ranges are reduced as soon as the backward of their layer has
                           finished (hook called from the fused layer functions), overlapping the
                           exchange with the rest of the backward pass.

This is the only collective in the timed loop; there is no data-path collective.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .arena import ParamArena


def _ranges(arena: ParamArena, params) -> List[Tuple[int, int]]:
    spans = sorted((arena.offsets[id(p)], arena.offsets[id(p)] + (p.numel() + arena.ALIGN - 1) // arena.ALIGN * arena.ALIGN) for p in params)
    out: List[Tuple[int, int]] = []
    for a, b in spans:
        if out and a <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], b))
        else:
            out.append((a, b))
    return out


def _reduce(t: torch.Tensor, group, wire_dtype=None):
    """Average across ranks: NCCL has a native AVG; gloo (CPU tests) sums and the caller scales.  wire_dtype (opt-in, e.g.
    torch.bfloat16) sends a down-cast copy and writes the result back into the fp32 slice when the collective has finished --
    half the bytes on the wire for two extra elementwise passes.  Returns (work, scale, wire_buffer)."""
    buf = t if wire_dtype is None or wire_dtype == t.dtype else t.to(wire_dtype)
    if dist.get_backend(group) == "nccl":
        return dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group, async_op=True), None, buf
    return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=True), 1.0 / dist.get_world_size(group), buf


def _finish(work, scale, buf, t):
    work.wait()
    if buf is not t:
        t.copy_(buf)
    if scale is not None:
        t.mul_(scale)


def sync_grads(arena: ParamArena, group=None, max_chunk: int = 64 << 20, wire_dtype=None) -> int:
    """All-reduce (average) the gradients touched in this step.  Returns the number of elements exchanged."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    works, total = [], 0
    for a, b in _ranges(arena, arena._touched):
        for s in range(a, b, max_chunk):
            e = min(b, s + max_chunk)
            works.append((_reduce(arena.flat_grad[s:e], group, wire_dtype), arena.flat_grad[s:e]))
            total += e - s
    for (w, scale, buf), t in works:
        _finish(w, scale, buf, t)
    return total


class LayerOverlap:
    """Reduce each layer's gradient slice as soon as that layer's backward is done."""

    def __init__(self, arena: ParamArena, group=None, wire_dtype=None):
        self.arena, self.group, self.wire_dtype = arena, group, wire_dtype
        self.pending = []
        self.done_ids = set()

    def layer_done(self, module):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        params = [p for p in module.parameters() if p.grad is not None and id(p) not in self.done_ids]
        if not params:
            return
        for p in params:
            self.done_ids.add(id(p))
        for a, b in _ranges(self.arena, params):
            self.pending.append((_reduce(self.arena.flat_grad[a:b], self.group, self.wire_dtype), self.arena.flat_grad[a:b]))

    def finish(self):
        """After backward: exchange whatever was not covered by a layer hook, then wait for everything."""
        rest = [p for p in self.arena._touched if id(p) not in self.done_ids]
        if rest and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            for a, b in _ranges(self.arena, rest):
                self.pending.append((_reduce(self.arena.flat_grad[a:b], self.group, self.wire_dtype), self.arena.flat_grad[a:b]))
        for (w, scale, buf), t in self.pending:
            _finish(w, scale, buf, t)
        self.pending, self.done_ids = [], set()
