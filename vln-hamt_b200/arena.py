"""Flat parameter / gradient / bf16-shadow arena.

HBM layout (B200-first, explicit): all parameters of a model live in ONE contiguous fp32 buffer, with
a parallel fp32 gradient buffer and a bf16 shadow of the same length.

  * ``nn.Parameter.data`` of every parameter is a view into ``flat_param`` (state_dict keys / shapes
    are untouched, ``load_state_dict`` / in-place optimizers keep working);
  * the CUDA kernels write weight gradients straight into ``flat_grad`` views (wgrad GEMM epilogue,
    LN / embedding column sums) -- autograd never sees parameter gradients, so there is no
    per-parameter AccumulateGrad traffic; ``param.grad`` is attached to its view the first time a
    kernel touches it in a step, parameters not used by the step keep ``grad is None`` exactly like the
    reference (whose AdamW skips them, pretrain_src/optim/adamw.py:64-66);
  * the bf16 shadow (GEMM operands) is refreshed by ONE cast kernel per optimizer step;
  * data-parallel gradient exchange works on contiguous slices of ``flat_grad`` (dp.py);
  * query/key/value weights of each attention module are adjacent, so the fused [3H,H] QKV operand,
    its bias and its gradient are plain slices.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops


class ParamArena:
    ALIGN = 64  # elements; keeps every view 256-byte aligned (TMA needs 16 B)

    def __init__(self, module: nn.Module):
        self.module = module
        self.device = None
        self.flat_param: Optional[torch.Tensor] = None
        self.flat_grad: Optional[torch.Tensor] = None
        self.flat_bf16: Optional[torch.Tensor] = None
        self.offsets: Dict[int, int] = {}
        self.params: List[nn.Parameter] = []
        self._shadow_version = None
        self._fresh_version = None
        self._shadow_valid = False          # True only after a FULL cast of flat_param into flat_bf16 (see shadow_is_fresh)
        self._valid_version = None
        self.external_prologue = False      # graph.py: the cast / gradient zeroing / seed advance run eagerly OUTSIDE the captured body
        self._touched: List[nn.Parameter] = []
        self._sentinel: Optional[nn.Parameter] = None
        self.anchor = None
        self.seed = None
        self.used_ranges = []

    # ---------------------------------------------------------------- construction
    def _ordered_params(self) -> List[nn.Parameter]:
        """Registration order, except that the six tensors of every attention projection are emitted as
        [q.weight, k.weight, v.weight, q.bias, k.bias, v.bias] so fused QKV operands are contiguous slices."""
        group_of = {}
        for m in self.module.modules():
            q, k, v = (getattr(m, n, None) for n in ("query", "key", "value"))
            if all(isinstance(t, nn.Linear) and t.bias is not None for t in (q, k, v)):
                grp = [q.weight, k.weight, v.weight, q.bias, k.bias, v.bias]
                for t in grp:
                    group_of[id(t)] = grp
        seen, out = set(), []
        for _, p in self.module.named_parameters():
            for t in group_of.get(id(p), [p]):
                if id(t) not in seen:
                    seen.add(id(t))
                    out.append(t)
        return out

    def build(self):
        params = self._ordered_params()
        dev = params[0].device
        total, offsets = 0, {}
        for p in params:
            if p.dtype != torch.float32:
                raise RuntimeError("hamt_b200: parameters must be fp32 (bf16 shadows are derived)")
            offsets[id(p)] = total
            total += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        for p in params:
            o = offsets[id(p)]
            view = flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = None
        self.flat_param, self.offsets, self.params, self.device = flat, offsets, params, dev
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_bf16 = torch.empty(total, dtype=torch.bfloat16, device=dev)
        self._shadow_version = None
        self._fresh_version = None
        self._shadow_valid, self._valid_version = False, None     # flat_bf16 is uninitialised memory until the first full cast
        self._touched, self._sentinel = [], None
        self.anchor = torch.zeros((), dtype=torch.float32, device=dev, requires_grad=True)
        if self.seed is None or self.seed.device != dev:
            self.seed = torch.zeros(1, dtype=torch.int64, device=dev)

    def valid(self) -> bool:
        if self.flat_param is None:
            return False
        p = self.params[0]
        return p.device == self.device and p.data_ptr() == self.flat_param.data_ptr() + 4 * self.offsets[id(p)] and \
            self.params[-1].data_ptr() == self.flat_param.data_ptr() + 4 * self.offsets[id(self.params[-1])]

    def ensure(self):
        if not self.valid():
            self.build()

    # ---------------------------------------------------------------- per-step protocol
    def step_begin(self, training: bool, zero_grads: bool = True):
        """Call at the start of every model forward (graph.py calls it eagerly before a replay and sets `external_prologue`
        so that the captured body does not repeat it)."""
        if next(self.module.parameters()).device.type != "cuda":
            raise RuntimeError("hamt_b200: the compute path needs the model on a CUDA device (no CPU fallback)")
        self.ensure()
        if self.external_prologue:
            return
        # bf16 shadow refresh: ONE cast launch for every GEMM operand.  Training: every step (weights move every
        # optimizer step; in-place updates through `.data` bump no version counter, so this is unconditional --
        # 1 GB of traffic, ~0.2 ms) unless the fused optimizer refreshed a fully valid shadow in its own pass.
        # Eval: only when a parameter's version counter moved or after training.
        ver = sum(p._version for p in self.params)
        if training and self._fresh_version is not None and ver == self._fresh_version and self._shadow_valid:
            pass        # the fused optimizer wrote the shadow together with the weights (optim.AdamW.step) and nothing changed since
        elif training or not self._shadow_valid or ver != self._shadow_version:
            ops.cast_bf16(self.flat_param, out=self.flat_bf16)
            self._shadow_valid, self._valid_version = True, ver
            self._shadow_version = None if training else ver
        self._fresh_version = None
        if zero_grads and training and self._sentinel is not None and self._sentinel.grad is None:
            # the caller cleared the gradients (zero_grad(set_to_none=True)): start a fresh accumulation
            self.flat_grad.zero_()
            for p in self._touched:
                p.grad = None
            self._touched, self._sentinel = [], None

    def mark_dirty(self):
        """Parameters were modified without bumping flat_param's version counter (e.g. through .data)."""
        self._shadow_version = None
        self._fresh_version = None
        self._shadow_valid = False

    def shadow_is_fresh(self):
        """Called by the fused optimizer: it wrote the bf16 shadow of every ACTIVE segment in the same pass as the fp32 weights.
        The next training step may skip its cast launch only if the shadow was completely valid before that pass (a full cast
        happened since build / mark_dirty) and no parameter's version counter moved since that cast (load_state_dict, manual
        edits): otherwise inactive segments -- or everything, when the optimizer steps before the first forward like the
        reference loop does (main_r2r.py:229-230) -- would still hold stale or uninitialised values."""
        ver = sum(p._version for p in self.params)
        if self._shadow_valid and ver == self._valid_version:
            self._fresh_version = ver
        else:
            self._fresh_version = None
            self._shadow_valid = False
        self._shadow_version = None

    def grads_are_zero(self):
        """Called by the fused optimizer after it zeroed the gradients of every touched parameter in its update pass: the next
        step starts a fresh accumulation without the flat memset."""
        for p in self._touched:
            p.grad = None
        self._touched, self._sentinel = [], None

    # ---------------------------------------------------------------- views
    def w16(self, p: nn.Parameter) -> torch.Tensor:
        o = self.offsets[id(p)]
        return self.flat_bf16[o:o + p.numel()].view(p.shape)

    def grad(self, p: nn.Parameter) -> torch.Tensor:
        """fp32 gradient view of p (attaches it as p.grad on first use in the step)."""
        o = self.offsets[id(p)]
        g = self.flat_grad[o:o + p.numel()].view(p.shape)
        if p.grad is None:
            p.grad = g
            self._touched.append(p)
            if self._sentinel is None:
                self._sentinel = p
        elif p.grad.data_ptr() != g.data_ptr():
            raise RuntimeError("hamt_b200: a parameter's .grad was replaced by a foreign tensor; use zero_grad() or grad=None")
        return g

    def _span(self, plist, rows_each, attr):
        """Contiguity check + merged 2-D view for adjacent parameters (fused QKV)."""
        o0 = self.offsets[id(plist[0])]
        n = plist[0].numel()
        for i, p in enumerate(plist):
            if self.offsets[id(p)] != o0 + i * n or p.numel() != n:
                return None
        tail = plist[0].shape[1:]
        buf = getattr(self, attr)
        return buf[o0:o0 + n * len(plist)].view(rows_each * len(plist), *tail)

    def fused_w16(self, plist) -> torch.Tensor:
        v = self._span(plist, plist[0].shape[0], "flat_bf16")
        if v is None:
            raise RuntimeError("hamt_b200: parameters to fuse are not adjacent in the arena")
        return v

    def fused_param(self, plist) -> torch.Tensor:
        return self._span(plist, plist[0].shape[0], "flat_param")

    def fused_grad(self, plist) -> torch.Tensor:
        for p in plist:
            self.grad(p)
        return self._span(plist, plist[0].shape[0], "flat_grad")

    def next_seed(self):
        """Advance the device-resident dropout seed (one tiny in-place add; CUDA-graph friendly)."""
        if not self.external_prologue:
            self.seed.add_(0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF)

    def run_seed(self) -> torch.Tensor:
        """A private copy of the current seed for ONE forward (Run): the backward kernels regenerate the dropout masks from the
        pointer they are handed, so a forward that is followed by other forwards before its backward (the finetune agent runs
        dozens of 'language' / 'history' / 'visual' calls and then one loss.backward(), agent_cmt.py:562) must not share a
        seed cell that those later forwards advance.  Inside a captured graph the copy lives in the graph's memory pool and is
        re-made from the live seed on every replay."""
        return self.seed.clone()
