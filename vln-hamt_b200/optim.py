"""Fused optimizer step for the HAMT hot path behind the reference's optimizer API.

Mirrors pretrain_src/optim: `AdamW(params, lr, betas, eps, weight_decay, correct_bias)` (adamw.py:13-110, the HF variant),
`build_optimizer(model, opts)` (misc.py:12-37: two groups, no decay for names containing bias / LayerNorm.bias / LayerNorm.weight),
`warmup_linear` / `get_lr_sched` (sched.py:17-30), and the training loop's `clip_grad_norm_` + `optimizer.step()` +
`optimizer.zero_grad()` sequence (main_r2r.py:252-281).

The reference loops over ~400 parameters in python (~8 eager kernels each).  Here all parameters, gradients and both Adam moments
are flat fp32 buffers (arena.py), a step is THREE launches (hamt_optim.cu): squared-norm partials, a one-block prepare (norm, clip
coefficient, per-parameter step counters and bias-corrected step sizes) and one update pass that also writes the bf16 weight shadow
the GEMMs read and zeroes the gradients.  Per-parameter semantics are kept: a parameter whose `.grad is None` in this step (its task
did not run) is skipped and keeps its own `state["step"]` (adamw.py:64-66).  The learning rate lives in a device scalar, so the
step is CUDA-graph capturable.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import torch

from . import _lib
from .arena import ParamArena


def warmup_linear(step: int, warmup_step: int, tot_step: int) -> float:
    """BERT schedule (sched.py:17-21)."""
    if step < warmup_step:
        return step / warmup_step
    return max(0, (tot_step - step) / (tot_step - warmup_step))


def get_lr_sched(global_step: int, opts) -> float:
    """sched.py:24-30."""
    lr_this_step = opts.learning_rate * warmup_linear(global_step, opts.warmup_steps, opts.num_train_steps)
    if lr_this_step <= 0:
        lr_this_step = 1e-8
    return lr_this_step


class AdamW:
    """HF AdamW over a ParamArena.  `params`: iterable of parameters or of dicts {'params': [...], 'weight_decay': w} exactly like
    torch.optim; every parameter must live in `arena`."""

    def __init__(self, arena: ParamArena, params: Iterable, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-6,
                 weight_decay: float = 0.0, correct_bias: bool = True):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[1]))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        arena.ensure()
        self.arena = arena
        params = list(params)
        if params and not isinstance(params[0], dict):
            params = [{"params": params}]
        self.param_groups: List[Dict] = []
        for g in params:
            self.param_groups.append(dict(params=list(g["params"]), lr=g.get("lr", lr), betas=g.get("betas", betas), eps=g.get("eps", eps),
                                          weight_decay=g.get("weight_decay", weight_decay), correct_bias=g.get("correct_bias", correct_bias)))
        b, e, c = self.param_groups[0]["betas"], self.param_groups[0]["eps"], self.param_groups[0]["correct_bias"]
        if any(g["betas"] != b or g["eps"] != e or g["correct_bias"] != c for g in self.param_groups):
            raise ValueError("hamt_b200.optim.AdamW: betas / eps / correct_bias must be the same for all groups (one fused pass)")
        dev = arena.flat_param.device
        total = arena.flat_param.numel()
        # segment tables: one segment per optimised parameter
        self.seg_params: List[torch.nn.Parameter] = []
        chunk_seg = torch.full((total // 64,), -1, dtype=torch.int32)
        wd, ends = [], []
        seen = set()
        for g in self.param_groups:
            for p in g["params"]:
                if id(p) in seen:
                    raise ValueError("some parameters appear in more than one parameter group")
                if id(p) not in arena.offsets:
                    raise ValueError("hamt_b200.optim.AdamW: parameter is not part of the arena")
                seen.add(id(p))
                s = len(self.seg_params)
                o = arena.offsets[id(p)]
                chunk_seg[o // 64:(o + p.numel() + 63) // 64] = s
                self.seg_params.append(p)
                wd.append(float(g["weight_decay"]))
                ends.append(o + p.numel())
        self.nseg = len(self.seg_params)
        self.chunk_seg = chunk_seg.to(dev)
        self.seg_wd = torch.tensor(wd, dtype=torch.float32, device=dev)
        self.seg_end = torch.tensor(ends, dtype=torch.int64, device=dev)
        self.seg_step = torch.zeros(self.nseg, dtype=torch.int32, device=dev)
        self.seg_step_size = torch.zeros(self.nseg, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.lr_dev = torch.tensor([self.param_groups[0]["lr"]], dtype=torch.float32, device=dev)
        self._lr_last = None
        self.workspace = torch.zeros(_lib.load().hamt_adamw_workspace_floats(), dtype=torch.float32, device=dev)
        self._active_cache: Dict[Tuple[int, ...], torch.Tensor] = {}

    # -------------------------------------------------------------------------------------------------------------
    def _active(self) -> torch.Tensor:
        """uint8 [nseg]: 1 where the parameter has a gradient this step.  The touched set is static per task, so the device
        copies are cached by pattern (no host -> device traffic in steady state)."""
        key = tuple(i for i, p in enumerate(self.seg_params) if p.grad is not None)
        t = self._active_cache.get(key)
        if t is None:
            m = torch.zeros(self.nseg, dtype=torch.uint8)
            if key:
                m[list(key)] = 1
            t = m.to(self.seg_wd.device)
            self._active_cache[key] = t
        return t

    def set_lr(self, lr: float):
        """`for g in optimizer.param_groups: g['lr'] = lr_this_step` (main_r2r.py:257-259)."""
        for g in self.param_groups:
            g["lr"] = lr

    def step(self, max_grad_norm: float = -1.0, zero_grad: bool = False, want_norm: bool = False) -> Optional[torch.Tensor]:
        """clip_grad_norm_(parameters, max_grad_norm) if max_grad_norm > 0, then the AdamW update; with zero_grad the gradients
        of the updated parameters are zeroed in the same pass (optimizer.zero_grad()).  Returns the global gradient norm
        (device scalar, before clipping) when clipping or want_norm, else None."""
        lr = self.param_groups[0]["lr"]
        if any(g["lr"] != lr for g in self.param_groups):
            raise ValueError("hamt_b200.optim.AdamW: one learning rate for all groups (the reference sets them together)")
        if lr != self._lr_last:
            self.lr_dev.fill_(lr)          # scalar travels as a kernel argument: no pinned staging to race with
            self._lr_last = lr
        a = self.arena
        g0 = self.param_groups[0]
        active = self._active()
        rc = _lib.load().hamt_adamw_step(a.flat_param.data_ptr(), a.flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                         a.flat_bf16.data_ptr(), a.flat_param.numel(), self.chunk_seg.data_ptr(), self.seg_end.data_ptr(), self.nseg, active.data_ptr(),
                                         self.seg_wd.data_ptr(), self.seg_step.data_ptr(), self.seg_step_size.data_ptr(), self.lr_dev.data_ptr(),
                                         float(g0["betas"][0]), float(g0["betas"][1]), float(g0["eps"]), int(bool(g0["correct_bias"])),
                                         float(max_grad_norm), int(want_norm), int(zero_grad), self.workspace.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "adamw_step")
        a.shadow_is_fresh()
        if zero_grad:
            seg_ids = {id(p) for p in self.seg_params}
            if all(id(p) in seg_ids for p in a._touched):
                a.grads_are_zero()          # every touched gradient was zeroed by the update pass
            self.zero_grad()
        return self.workspace[0] if (max_grad_norm > 0 or want_norm) else None

    def zero_grad(self, set_to_none: bool = True):
        """optimizer.zero_grad(): drops the .grad views; the arena zeroes the flat gradient buffer when the next step begins
        (or step(zero_grad=True) already did it in the update pass)."""
        for p in self.seg_params:
            p.grad = None

    # -------------------------------------------------------------------------------------------------------------
    @property
    def state(self) -> Dict:
        """Per-parameter state views like torch.optim.Optimizer.state (reads the step counters back: host sync, debugging / checkpoint)."""
        steps = self.seg_step.cpu().tolist()
        out = {}
        for s, p in enumerate(self.seg_params):
            if steps[s] == 0:
                continue
            o = self.arena.offsets[id(p)]
            out[p] = dict(step=steps[s], exp_avg=self.exp_avg[o:o + p.numel()].view(p.shape), exp_avg_sq=self.exp_avg_sq[o:o + p.numel()].view(p.shape))
        return out


    # -------------------------------------------------------------------------------------------------------------
    def state_dict(self) -> Dict:
        """torch.optim.Optimizer.state_dict() layout ({'state': {index: {...}}, 'param_groups': [{..., 'params': [indices]}]}), what
        the reference's checkpointing stores (pretrain_src/utils/save.py:42 `optimizer.state_dict()`; finetune agent_cmt.py:616).
        Parameters that were never updated (their task never ran) have no entry, like the reference's lazily created state."""
        steps = self.seg_step.cpu().tolist()
        index, groups, i0 = {}, [], 0
        for g in self.param_groups:
            ids = list(range(i0, i0 + len(g["params"])))
            for i, p in zip(ids, g["params"]):
                index[id(p)] = i
            i0 += len(g["params"])
            groups.append({**{k: v for k, v in g.items() if k != "params"}, "params": ids})
        state = {}
        for s, p in enumerate(self.seg_params):
            if steps[s] == 0:
                continue
            o = self.arena.offsets[id(p)]
            state[index[id(p)]] = dict(step=steps[s], exp_avg=self.exp_avg[o:o + p.numel()].view(p.shape).clone(),
                                       exp_avg_sq=self.exp_avg_sq[o:o + p.numel()].view(p.shape).clone())
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: Dict):
        """Inverse of state_dict(); also accepts a state_dict written by the reference's own AdamW (same layout)."""
        groups = sd["param_groups"]
        if len(groups) != len(self.param_groups) or any(len(a["params"]) != len(b["params"]) for a, b in zip(groups, self.param_groups)):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        by_index = {}
        for g_saved, g in zip(groups, self.param_groups):
            for k, v in g_saved.items():
                if k != "params":
                    g[k] = v
            for i, p in zip(g_saved["params"], g["params"]):
                by_index[i] = p
        seg_of = {id(p): s for s, p in enumerate(self.seg_params)}
        steps = torch.zeros(self.nseg, dtype=torch.int32)
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for i, st in sd["state"].items():
            p = by_index[int(i)]
            o = self.arena.offsets[id(p)]
            steps[seg_of[id(p)]] = int(st["step"])
            self.exp_avg[o:o + p.numel()].view(p.shape).copy_(st["exp_avg"])
            self.exp_avg_sq[o:o + p.numel()].view(p.shape).copy_(st["exp_avg_sq"])
        self.seg_step.copy_(steps)
        self._lr_last = None


def build_optimizer(model, opts) -> AdamW:
    """misc.py:12-37 for opts.optim == 'adamw' (the shipped recipe, pretrain_r2r.json:24): decay everything except names containing
    'bias', 'LayerNorm.bias', 'LayerNorm.weight'."""
    if getattr(opts, "optim", "adamw") != "adamw":
        raise ValueError("invalid optimizer")       # same message as the reference's fall-through
    param_optimizer = list(model.named_parameters())
    no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]
    groups = [
        {"params": [p for n, p in param_optimizer if not any(nd in n for nd in no_decay)], "weight_decay": opts.weight_decay},
        {"params": [p for n, p in param_optimizer if any(nd in n for nd in no_decay)], "weight_decay": 0.0},
    ]
    arena = model.arena()
    arena.ensure()
    return AdamW(arena, groups, lr=opts.learning_rate, betas=tuple(opts.betas))
