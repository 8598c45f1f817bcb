"""Host -> device batch transport for the hot path: the role of the reference's PrefetchLoader
(pretrain_src/data/loader.py:90-125: a side CUDA stream copies batch i+1 while batch i computes; `move_to_cuda` walks the
collated dict tensor by tensor, ~20 cudaMemcpyAsync per batch).

B200 version: a collated batch is PACKED -- every tensor of the dict (and of the ITM negative plan) lives at a fixed,
256-byte-aligned offset of ONE pinned host blob, mirrored by one device blob; a step's transfer is a single
cudaMemcpyAsync (~103 MB at batch 64, PCIe-bound) on the copy stream and the dict the model sees is a set of views into
the device blob, created once.  A collate function writes straight into `host_views` (no intermediate copy); the
captured step copies blob -> its static inputs with one device-to-device copy (GraphedStep.__call__).

`LossReader` is the device -> host side: the step's scalar result goes to a pinned slot through an async copy + event, so
reading step i-1's loss waits for step i-1 only (Tensor.item() synchronises the whole stream, i.e. also step i that was
just enqueued, and serialises host and device).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

_ALIGN = 256


def flatten(batch: Dict) -> List[Tuple[tuple, torch.Tensor]]:
    """(path, tensor) for every tensor the step consumes, in sorted key order.  Keys starting with '_' are host-side
    side-information and are not transported."""
    out = []
    for k in sorted(batch):
        v = batch[k]
        if k.startswith("_"):
            continue
        if torch.is_tensor(v):
            out.append(((k,), v))
        elif k == "itm_plan" and v is not None:
            if v[0] is not None:
                out.append(((k, 0), v[0]))
            for i, t in enumerate(v[1]):
                out.append(((k, 1, i), t))
    return out


class Layout:
    """Offsets of a batch signature inside a blob."""

    def __init__(self, batch: Dict):
        self.entries = []        # (path, offset, nbytes, shape, dtype)
        off = 0
        for path, t in flatten(batch):
            nb = t.numel() * t.element_size()
            self.entries.append((path, off, nb, tuple(t.shape), t.dtype))
            off += (nb + _ALIGN - 1) // _ALIGN * _ALIGN
        self.nbytes = max(off, _ALIGN)
        self.payload_bytes = sum(e[2] for e in self.entries)
        self.other = {k: v for k, v in batch.items() if not torch.is_tensor(v) and k != "itm_plan" and not k.startswith("_")}
        self.has_plan = batch.get("itm_plan") is not None
        self.plan_has_neg = self.has_plan and batch["itm_plan"][0] is not None
        self.key = tuple((p, s, str(d)) for p, _, _, s, d in self.entries)

    def views(self, blob: torch.Tensor) -> Dict:
        """The batch dict as views into `blob` (uint8, >= nbytes)."""
        d = dict(self.other)
        plan_list = {}
        neg = None
        for path, off, nb, shape, dtype in self.entries:
            v = blob[off:off + nb].view(dtype).view(shape) if nb else torch.empty(shape, dtype=dtype, device=blob.device)
            if len(path) == 1:
                d[path[0]] = v
            elif path[1] == 0:
                neg = v
            else:
                plan_list[path[2]] = v
        if self.has_plan:
            d["itm_plan"] = (neg, [plan_list[i] for i in sorted(plan_list)])
        return d


class PackedBatch:
    """One batch slot: pinned host blob + device blob + dict views of both."""

    def __init__(self, example: Dict, device):
        self.layout = Layout(example)
        self.host = torch.empty(self.layout.nbytes, dtype=torch.uint8).pin_memory()
        self.dev = torch.empty(self.layout.nbytes, dtype=torch.uint8, device=device)
        self.host_views = self.layout.views(self.host)
        self.dev_views = self.layout.views(self.dev)
        self.dev_views["_packed"] = self
        self.ready = torch.cuda.Event()
        self.fill(example)

    def fill(self, batch: Dict, only_plan: bool = False):
        """Collate into the pinned blob (what a dataloader worker does once per batch)."""
        dst = dict(flatten(self.host_views))
        for path, t in flatten(batch):
            if only_plan and path[0] != "itm_plan":
                continue
            dst[path].copy_(t)

    def to_device(self, stream: torch.cuda.Stream):
        """One cudaMemcpyAsync of the whole blob on `stream`; `self.ready` fires when it has landed."""
        with torch.cuda.stream(stream):
            self.dev.copy_(self.host, non_blocking=True)
            self.ready.record(stream)
        return self.dev_views


class LossReader:
    """Pinned ring of scalar results: push(loss_tensor) enqueues an async device -> host copy behind the step and records
    an event; pop() waits for the OLDEST outstanding event only."""

    def __init__(self, depth: int = 4):
        self.buf = torch.empty(depth, dtype=torch.float32).pin_memory()
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self.depth, self.head, self.tail = depth, 0, 0

    def push(self, value: torch.Tensor):
        assert self.head - self.tail < self.depth, "LossReader ring full: pop() before pushing more"
        s = self.head % self.depth
        self.buf[s:s + 1].copy_(value.detach().reshape(1), non_blocking=True)     # detach: the ring must not join the autograd graph
        self.events[s].record()
        self.head += 1

    def pending(self) -> int:
        return self.head - self.tail

    def pop(self) -> float:
        s = self.tail % self.depth
        self.events[s].synchronize()
        self.tail += 1
        return float(self.buf[s])
