"""CUDA-graph capture of one pretraining step (forward + backward [+ gradient exchange]) per task.

The eager path issues ~470 kernel launches per step from Python (ctypes + autograd), ~17 ms of host work at batch 64
-- as long as the GPU work itself.  A captured step replays the same launches with one ``cudaGraphLaunch``:

    step = GraphedStep(model, "sap", example_batch)         # captures after 2 warm-up runs
    loss = step(batch)                                      # copies the batch into the static buffers, replays

What makes the step capturable (host-sync removal):
  * dropout seeds live in device memory and are advanced by a captured in-place add;
  * the boolean-mask gathers of MLM / MRC (`hidden[mask]`, pretrain_cmt.py:161-165) use row indices supplied with the
    batch (computed on the host from the labels the collate function built there anyway) instead of a device `nonzero`;
    the number of masked rows is part of the graph key;
  * the ITM negatives (vilmodel.py:676-704) are drawn on the host with the reference's RNG call order and copied into
    static index buffers before the replay;
  * weight-gradient accumulation targets (the arena's flat gradient buffer) have fixed addresses.
All graphs of one model share one memory pool (they never run concurrently).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .loader import Layout, flatten
from .vilmodel import itm_negative_plan

_INDEX_KEYS = {"mlm": ("txt_label_rows", "txt_labels", -1), "mrc": ("hist_mrc_rows", "hist_mrc_masks", None)}


def add_sync_free_extras(task: str, batch: Dict, device=None) -> Dict:
    """Augment a (host or device) batch with what the captured step needs: masked-row indices for MLM / MRC and the ITM
    negative plan.  Uses the global numpy / torch RNGs for ITM exactly like the reference's forward_itm."""
    b = dict(batch)
    if task in _INDEX_KEYS:
        key, src, skip = _INDEX_KEYS[task]
        m = b[src]
        mask = (m != skip) if skip is not None else m
        b[key] = torch.nonzero(mask.reshape(-1), as_tuple=False).squeeze(1)
    if task == "itm" and not b.get("_itm_device_negatives", False):       # (config.itm_device_negatives: the step draws them on the device)
        hm = b.get("_hist_masks_host")           # host copy kept next to a device-resident batch: no device->host sync
        if hm is None:
            hm = b["hist_masks"].cpu()
        neg, shuf = itm_negative_plan(hm.shape[0], hm, hm.shape[1] - 1, 4)        # hist_masks = CLS slot + T steps
        b["itm_plan"] = (neg, shuf)
    if device is not None:
        for k, v in list(b.items()):
            if torch.is_tensor(v):
                b[k] = v.to(device, non_blocking=True)
            elif k == "itm_plan":
                b[k] = (None if v[0] is None else v[0].to(device, non_blocking=True), [t.to(device, non_blocking=True) for t in v[1]])
    return b


def _signature(task, batch):
    sig = [task]
    for k in sorted(batch):
        v = batch[k]
        if k.startswith("_"):
            continue
        if torch.is_tensor(v):
            sig.append((k, tuple(v.shape), str(v.dtype)))
        elif k == "itm_plan" and v is not None:
            sig.append((k, None if v[0] is None else tuple(v[0].shape), len(v[1])))
        elif v is None:
            sig.append((k, None))
    return tuple(sig)


class GraphedStep:
    """One captured `loss = model(batch, task); loss.mean().backward()` for a fixed batch signature.

    The captured body holds ONLY the forward + backward (+ gradient exchange) launches.  What depends on the training
    loop's state runs eagerly in `__call__` before the replay: the bf16 weight-shadow cast (skipped when the fused
    optimizer refreshed it), the gradient reset (skipped with `accumulate=True`: gradient_accumulation_steps > 1 in the
    reference loop, main_r2r.py:243-249) and the dropout-seed advance."""

    _pools: Dict[int, object] = {}

    def __init__(self, model, task: str, batch: Dict, post_backward=None, warmup: int = 2, bwd_sm_limit: int = 0):
        self.model, self.task = model, task
        dev = next(model.parameters()).device
        arena = model.arena()
        arena.ensure()
        # static inputs of the captured step: one packed device blob (loader.Layout), so a packed batch arrives with ONE copy
        self.layout = Layout(batch)
        self.blob = torch.empty(self.layout.nbytes, dtype=torch.uint8, device=dev)
        self.static = self.layout.views(self.blob)
        self._static_flat = dict(flatten(self.static))
        for path, t in flatten(batch):
            self._static_flat[path].copy_(t)
        self.signature = _signature(task, batch)

        def body():
            loss = model(self.static, task, compute_loss=True)
            # data-parallel: the backward GEMMs leave SMs free for the NCCL kernels that overlap them (grid sizes are baked in at capture)
            if bwd_sm_limit:
                _lib.load().hamt_gemm_set_sm_limit(bwd_sm_limit)
            try:
                loss.mean().backward()
            finally:
                if bwd_sm_limit:
                    _lib.load().hamt_gemm_set_sm_limit(0)
            if post_backward is not None:
                post_backward()
            return loss

        # A capture may happen in the middle of training (first batch of a new signature): the warm-up passes must not disturb
        # gradients that are being accumulated, nor the dropout-seed sequence.
        prev_touched, prev_sentinel = list(arena._touched), arena._sentinel
        grad_backup = arena.flat_grad.clone() if prev_touched else None
        arena.step_begin(True, zero_grads=False)       # a valid bf16 shadow for the warm-up passes
        for p in prev_touched:                         # so that what is attached below is exactly this task's touched set
            p.grad = None
        arena._touched, arena._sentinel = [], None
        arena.external_prologue = True
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            pool = GraphedStep._pools.get(id(model))
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            # thread_local: the NCCL watchdog thread (data-parallel runs) polls CUDA events while we capture
            with torch.cuda.graph(self.graph, pool=pool, capture_error_mode="thread_local"):
                self.loss = body()
            self.native_launches = _lib.launch_count() - n0
        finally:
            arena.external_prologue = False
        if pool is None:
            GraphedStep._pools[id(model)] = self.graph.pool()
        # after a replay the gradients of exactly the parameters this task touches are valid
        self.touched = list(arena._touched)
        self.touched_ids = {id(p) for p in self.touched}
        # restore the pre-capture gradient state
        for p in arena.params:
            p.grad = None
        if grad_backup is not None:
            arena.flat_grad.copy_(grad_backup)
            for p in prev_touched:
                o = arena.offsets[id(p)]
                p.grad = arena.flat_grad[o:o + p.numel()].view(p.shape)
            arena._touched, arena._sentinel = prev_touched, prev_sentinel
        else:
            arena.flat_grad.zero_()
            arena._touched, arena._sentinel = [], None

    def matches(self, task, batch) -> bool:
        return _signature(task, batch) == self.signature

    def __call__(self, batch: Dict, accumulate: bool = False) -> torch.Tensor:
        arena = self.model.arena()
        # ---- eager prologue (never captured): shadow cast, gradient reset, seed advance ----
        arena.step_begin(True, zero_grads=False)
        keep = []
        if accumulate:
            keep = [p for p in arena._touched if p.grad is not None]
        elif arena._touched:
            arena.flat_grad.zero_()
        arena.next_seed()
        packed = batch.get("_packed")
        if packed is not None and packed.layout.key == self.layout.key:
            self.blob.copy_(packed.dev, non_blocking=True)           # one device-to-device copy of the whole batch
        else:
            for path, t in flatten(batch):
                self._static_flat[path].copy_(t, non_blocking=True)
        self.graph.replay()
        # host-side gradient bookkeeping of this task (graphs of other tasks may have run in between)
        keep_ids = {id(p) for p in keep}
        for p in arena._touched:
            if id(p) not in self.touched_ids and id(p) not in keep_ids:
                p.grad = None
        now = list(self.touched) + [p for p in keep if id(p) not in self.touched_ids]
        for p in now:
            if p.grad is None:
                o = arena.offsets[id(p)]
                p.grad = arena.flat_grad[o:o + p.numel()].view(p.shape)
        arena._touched = now
        arena._sentinel = now[0] if now else None
        return self.loss


class GraphedTrainer:
    """Cache of captured steps keyed by (task, batch signature); falls back to capture-on-first-use."""

    def __init__(self, model, post_backward=None, bwd_sm_limit: int = 0):
        self.model, self.post_backward, self.steps, self.bwd_sm_limit = model, post_backward, {}, bwd_sm_limit

    def step(self, task: str, batch: Dict, accumulate: bool = False) -> torch.Tensor:
        sig = _signature(task, batch)
        st = self.steps.get(sig)
        if st is None:
            st = GraphedStep(self.model, task, batch, self.post_backward, bwd_sm_limit=self.bwd_sm_limit)
            self.steps[sig] = st
        return st(batch, accumulate=accumulate)
