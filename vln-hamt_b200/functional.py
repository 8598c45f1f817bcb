"""Fused layer blocks (hand-derived forward + backward) and their autograd wrappers.

Activations are bf16 2-D tensors ``[tokens, hidden]``.  Parameter gradients are written by the CUDA
kernels directly into the arena (arena.py) -- the autograd graph only carries activation gradients
between ~20 coarse nodes per step (one per transformer layer / embedder / head op).

Block structure follows the reference modules:
  attn_block  = BertAttention   (vilmodel.py:146-157)  = fused-QKV GEMM -> fused attention -> dense -> dropout+residual+LN
  ffn_block   = BertIntermediate + BertOutput (:159-186) = GEMM(+bias+GELU) -> GEMM -> dropout+residual+LN
  cross_block = BertXAttention in BOTH directions with its shared weights (:351-383)
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import ops
from .arena import ParamArena

BF16 = torch.bfloat16
F32 = torch.float32


def _new32(x: torch.Tensor) -> torch.Tensor:
    """fp32 twin of a bf16 activation buffer: the residual stream is carried in full precision between LayerNorms (ln_fwd)."""
    return torch.empty(x.shape, dtype=F32, device=x.device)


def _as32(x16: torch.Tensor, x32: Optional[torch.Tensor]) -> torch.Tensor:
    return x32 if x32 is not None else x16.detach().float()


# Weight-gradient GEMMs on a second stream.  dW = dY^T X only feeds the gradient buffer: nothing later in the backward pass reads
# it, while the dgrad GEMM that consumes the same dY is on the critical path.  Every GEMM is a persistent grid holding all 148 SMs
# (one ~200 KB CTA each), so two dependent launches never overlap: the measured cost of a launch is ~8 us on top of its
# tensor-pipe time (pipeline fill, last-tile epilogue, grid drain; profiles/r02_bench_signatures_final.txt, intercept of time against
# K at M = 5120) and there are ~260 GEMM launches per step.  With the wgrads forked onto a side stream the block scheduler places
# their CTAs on the SMs the critical-path kernel frees while it drains (and vice versa), in eager mode and -- as parallel
# branches -- inside the captured step graph.  The side stream is joined where a layer's gradients must be final (Run.done:
# data-parallel hook) and at the end of the backward pass.  HAMT_WGRAD_STREAM=0 restores the single-stream order.
WGRAD_SIDE_STREAM = os.environ.get("HAMT_WGRAD_STREAM", "1") != "0"
_side_streams = {}
# (Measured and removed in round 2: running the text branch -- embedder + 9 text layers, M = 5 120 -- on its own stream next to the
# panorama encoder, forward and backward.  The CTAs of the two branches interleave, but every kernel of the path is a persistent grid
# that wants all 148 SMs: 11.18 ms/step against 10.95 without; profiles/r02_ab_streams.txt.)
# FFN blocks: save gelu'(pre) in the forward epilogue and multiply by it in the dgrad epilogue (1) or save the pre-activation and evaluate
# the derivative in the dgrad epilogue (0, the round-1 pair) -- A/B switch, default by measurement (profiles/r02_ab_streams.txt, r02_kbench_ffn_epilogues.txt)
GELU_DERIVATIVE = os.environ.get("HAMT_GELU_DER", "1") != "0"


def _side_stream(device) -> "torch.cuda.Stream":
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=device)
    return st


class Run:
    """Per-forward context: arena, mode, dropout probabilities and the dropout call-site counter."""

    def __init__(self, arena: ParamArena, training: bool, heads: int, eps: float, seed=None):
        self.arena = arena
        self.seed = seed            # this forward's private dropout seed cell (arena.run_seed()); None in eval
        self.training = training
        self.heads = heads
        self.eps = eps
        self.save = torch.is_grad_enabled()
        self._site = 0
        self.uses = {}
        self._side = None           # side stream with wgrad work in flight (None: nothing to join)
        self._side_refs = []        # operands of in-flight side-stream GEMMs (kept alive until retired: they were allocated on the main stream)
        self._refs_total = 0        # operands ever appended / already retired (absolute counters)
        self._refs_base = 0
        self._marks = []            # (event on the side stream, _refs_total at that point, layer whose backward ended there)
        self._join_queued = False

    def fork_wgrad(self, fn, *keep):
        """Run fn() (a weight-gradient GEMM) on the side stream, ordered after everything enqueued so far on the current stream."""
        cur = torch.cuda.current_stream()
        if not WGRAD_SIDE_STREAM:
            fn()
            return
        side = _side_stream(cur.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            fn()
        self._side = side
        self._fork_stream = cur     # the operands kept below were allocated on this stream: it is the one that must wait before they go
        self._side_refs.append(keep)
        self._refs_total += 1
        self._ensure_final_join()

    def _ensure_final_join(self):
        """(backward pass only) final join at the end of this backward pass (heads / embedders have no layer hook); inside a captured
        step this is also what re-joins the forked streams before the capture ends."""
        if not self._join_queued:
            self._join_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._final_join)

    def _hook(self, layer):
        hook = getattr(self.arena, "layer_hook", None)
        if hook is not None and layer is not None:
            hook(layer)

    def _wait_targets(self):
        cur = torch.cuda.current_stream()
        fs = getattr(self, "_fork_stream", None)
        return [cur] if (fs is None or fs == cur) else [cur, fs]

    def _retire(self, mark):
        ev, total, layer = mark
        for st in self._wait_targets():
            st.wait_event(ev)
        del self._side_refs[:total - self._refs_base]
        self._refs_base = total
        self._hook(layer)           # the layer's weight gradients are final: data-parallel exchange may start

    def join_side(self):
        """Main stream waits for all forked weight-gradient work; pending layer hooks fire."""
        if self._side is not None:
            for st in self._wait_targets():
                st.wait_stream(self._side)
            self._side = None
            self._side_refs.clear()
            self._refs_base = self._refs_total
        marks, self._marks = self._marks, []
        for _, _, layer in marks:
            self._hook(layer)

    def _final_join(self):
        self._join_queued = False
        self.join_side()

    def used(self, layer):
        self.uses[id(layer)] = self.uses.get(id(layer), 0) + 1

    def done(self, layer):
        """Called at the end of a layer's backward.  Its weight gradients come from the side stream, so the layer is retired
        (operands released, data-parallel hook fired) one layer later, when waiting for them costs nothing."""
        n = self.uses.get(id(layer), 1) - 1
        self.uses[id(layer)] = n
        if self._side is None:
            if n == 0:
                self._hook(layer)
            return
        ev = torch.cuda.Event()
        ev.record(self._side)
        self._marks.append((ev, self._refs_total, layer if n == 0 else None))
        if len(self._marks) > 1:
            self._retire(self._marks.pop(0))

    def drop(self, module_or_p) -> ops.Drop:
        p = module_or_p if isinstance(module_or_p, float) else float(module_or_p.p)
        if not self.training or p <= 0.0:
            return ops.NO_DROP
        self._site += 1
        return ops.Drop(self.seed if self.seed is not None else self.arena.seed, self._site, p)


def _wgrad(run: "Run", dy: torch.Tensor, x: torch.Tensor, weight):
    """weight.grad[N,K] += dy[M,N]^T x[M,K]  (both operands MN-major, fp32 split-K accumulation), off the critical path."""
    if weight.requires_grad:
        g = run.arena.grad(weight)
        run.fork_wgrad(lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, out=g, accumulate=True), dy, x)


def _wgrad_fused(run: "Run", dy: torch.Tensor, x: torch.Tensor, ws):
    """Same for the fused q/k/v weight block (adjacent in the arena)."""
    g = run.arena.fused_grad(ws)
    run.fork_wgrad(lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, out=g, accumulate=True), dy, x)


def _bgrad(A: ParamArena, dy: torch.Tensor, bias):
    if bias is not None and bias.requires_grad:
        ops.colsum(dy, A.grad(bias))


# ------------------------------------------------------------------------------------------------
# attention block (self attention)
# ------------------------------------------------------------------------------------------------
def _qkv_params(att):
    return [att.query.weight, att.key.weight, att.value.weight], [att.query.bias, att.key.bias, att.value.bias]


def attn_block_fwd(run: Run, x, B: int, S: int, mask, att, out_mod, y_out=None, x32=None, y32_out=None):
    """x32: fp32 twin of x (residual), or None -> the bf16 x is the residual.  Returns (y, y32, saved)."""
    A, H = run.arena, x.shape[1]
    ws, bs = _qkv_params(att)
    qkv = ops.gemm(x, A.fused_w16(ws), bias=A.fused_param(bs))
    d_attn = run.drop(att.dropout)
    ctx, lse = ops.attn_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, S, S, run.heads, mask, d_attn, need_lse=run.save)
    t = ops.gemm(ctx, A.w16(out_mod.dense.weight), bias=out_mod.dense.bias)
    d_hid = run.drop(out_mod.dropout)
    ln = out_mod.LayerNorm
    y32 = _new32(x) if y32_out is None else y32_out
    y, z, mean, rstd = ops.ln_fwd(t, x if x32 is None else x32, ln.weight, ln.bias, run.eps, d_hid, save_z=run.save, out=y_out, out32=y32)
    saved = (x, qkv, ctx, lse, z, mean, rstd, d_attn, d_hid, mask, B, S) if run.save else None
    return y, y32, saved


def attn_block_bwd(run: Run, dy, saved, att, out_mod, dx_out=None):
    x, qkv, ctx, lse, z, mean, rstd, d_attn, d_hid, mask, B, S = saved
    A, H = run.arena, x.shape[1]
    ws, bs = _qkv_params(att)
    ln, dense = out_mod.LayerNorm, out_mod.dense
    dt, dx = ops.ln_bwd(dy, z, mean, rstd, ln.weight, A.grad(ln.weight), A.grad(ln.bias), A.grad(dense.bias), drop=d_hid, dres_out=dx_out)
    _wgrad(run, dt, ctx, dense.weight)
    dctx = ops.gemm(dt, A.w16(dense.weight), b_mn=True)
    dqkv = torch.empty_like(qkv)
    db = A.fused_grad(bs) if ws[0].requires_grad else None      # q/k/v bias gradients: column sums fused into the attention backward
    ops.attn_bwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], ctx, lse, dctx, dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:], B, S, S, run.heads,
                 mask, d_attn, dbias=db)
    if ws[0].requires_grad:
        _wgrad_fused(run, dqkv, x, ws)
    ops.gemm(dqkv, A.fused_w16(ws), b_mn=True, out=dx, accumulate=True)       # dx (residual path) += dqkv @ Wqkv
    return dx


# ------------------------------------------------------------------------------------------------
# feed-forward block
# ------------------------------------------------------------------------------------------------
def ffn_block_fwd(run: Run, x, inter, out_mod, y_out=None, x32=None, y32_out=None):
    A = run.arena
    w1, w2 = inter.dense.weight, out_mod.dense.weight
    if run.save:
        # h = gelu'(pre-activation), evaluated by the forward epilogue next to gelu itself: the backward epilogue is one multiply
        h = torch.empty((x.shape[0], w1.shape[0]), dtype=BF16, device=x.device)
        a = ops.gemm(x, A.w16(w1), bias=inter.dense.bias, act=ops.ACT_GELU, aux_mode=ops.AUX_STORE_DGELU if GELU_DERIVATIVE else ops.AUX_STORE_PRE, aux=h)
    else:
        h = None
        a = ops.gemm(x, A.w16(w1), bias=inter.dense.bias, act=ops.ACT_GELU)
    t = ops.gemm(a, A.w16(w2), bias=out_mod.dense.bias)
    d_hid = run.drop(out_mod.dropout)
    ln = out_mod.LayerNorm
    y32 = _new32(x) if y32_out is None else y32_out
    y, z, mean, rstd = ops.ln_fwd(t, x if x32 is None else x32, ln.weight, ln.bias, run.eps, d_hid, save_z=run.save, out=y_out, out32=y32)
    saved = (x, h, a, z, mean, rstd, d_hid) if run.save else None
    return y, y32, saved


def ffn_block_bwd(run: Run, dy, saved, inter, out_mod, dx_out=None):
    x, h, a, z, mean, rstd, d_hid = saved
    A = run.arena
    ln = out_mod.LayerNorm
    w1, w2 = inter.dense.weight, out_mod.dense.weight
    dt, dx = ops.ln_bwd(dy, z, mean, rstd, ln.weight, A.grad(ln.weight), A.grad(ln.bias), A.grad(out_mod.dense.bias), drop=d_hid, dres_out=dx_out)
    _wgrad(run, dt, a, w2)
    b1 = inter.dense.bias
    dh = ops.gemm(dt, A.w16(w2), b_mn=True, aux_mode=ops.AUX_MUL if GELU_DERIVATIVE else ops.AUX_MUL_DGELU, aux=h,
                  colsum=A.grad(b1) if (b1 is not None and b1.requires_grad) else None)     # bias gradient fused into the dgrad epilogue
    _wgrad(run, dh, x, w1)
    ops.gemm(dh, A.w16(w1), b_mn=True, out=dx, accumulate=True)
    return dx


# ------------------------------------------------------------------------------------------------
# cross attention, both directions, shared weights (rows [0:ML] = language, [ML:] = vision)
# ------------------------------------------------------------------------------------------------
def cross_block_fwd(run: Run, xcat, B: int, L: int, V: int, lang_mask, visn_mask, xatt, lang_ca: bool = True, xcat32=None):
    A, H = run.arena, xcat.shape[1]
    ML = B * L
    ws, bs = _qkv_params(xatt.att)
    qkv = ops.gemm(xcat, A.fused_w16(ws), bias=A.fused_param(bs))
    ctx = torch.empty_like(xcat)
    d_att_l, d_att_v = run.drop(xatt.att.dropout), run.drop(xatt.att.dropout)
    lse_l = None
    if lang_ca:
        _, lse_l = ops.attn_fwd(qkv[:ML, :H], qkv[ML:, H:2 * H], qkv[ML:, 2 * H:], B, L, V, run.heads, visn_mask, d_att_l, need_lse=run.save, out=ctx[:ML])
    _, lse_v = ops.attn_fwd(qkv[ML:, :H], qkv[:ML, H:2 * H], qkv[:ML, 2 * H:], B, V, L, run.heads, lang_mask, d_att_v, need_lse=run.save, out=ctx[ML:])
    ln, dense = xatt.output.LayerNorm, xatt.output.dense
    d_hid = run.drop(xatt.output.dropout)
    y32 = _new32(xcat)
    res = xcat if xcat32 is None else xcat32
    if lang_ca:
        t = ops.gemm(ctx, A.w16(dense.weight), bias=dense.bias)
        y, z, mean, rstd = ops.ln_fwd(t, res, ln.weight, ln.bias, run.eps, d_hid, save_z=run.save, out32=y32)
    else:
        # finetune no_lang_ca (vilmodel_cmt.py:379-384): language rows pass through unchanged
        y = torch.empty_like(xcat)
        y[:ML].copy_(xcat[:ML])
        y32[:ML].copy_(res[:ML])
        t = ops.gemm(ctx[ML:], A.w16(dense.weight), bias=dense.bias)
        _, z, mean, rstd = ops.ln_fwd(t, res[ML:], ln.weight, ln.bias, run.eps, d_hid, save_z=run.save, out=y[ML:], out32=y32[ML:])
    saved = (xcat, qkv, ctx, lse_l, lse_v, z, mean, rstd, d_att_l, d_att_v, d_hid, lang_mask, visn_mask, B, L, V, lang_ca) if run.save else None
    return y, y32, saved


def cross_block_bwd(run: Run, dy, saved, xatt):
    xcat, qkv, ctx, lse_l, lse_v, z, mean, rstd, d_att_l, d_att_v, d_hid, lang_mask, visn_mask, B, L, V, lang_ca = saved
    A, H = run.arena, xcat.shape[1]
    ML = B * L
    ws, bs = _qkv_params(xatt.att)
    ln, dense = xatt.output.LayerNorm, xatt.output.dense
    db = A.fused_grad(bs) if ws[0].requires_grad else None      # q/k/v bias gradients accumulate from both directions' backward kernels
    if lang_ca:
        dt, dx = ops.ln_bwd(dy, z, mean, rstd, ln.weight, A.grad(ln.weight), A.grad(ln.bias), A.grad(dense.bias), drop=d_hid)
        _wgrad(run, dt, ctx, dense.weight)
        dctx = ops.gemm(dt, A.w16(dense.weight), b_mn=True)
        dqkv = torch.empty_like(qkv)
        ops.attn_bwd(qkv[:ML, :H], qkv[ML:, H:2 * H], qkv[ML:, 2 * H:], ctx[:ML], lse_l, dctx[:ML], dqkv[:ML, :H], dqkv[ML:, H:2 * H],
                     dqkv[ML:, 2 * H:], B, L, V, run.heads, visn_mask, d_att_l, dbias=db)
        ops.attn_bwd(qkv[ML:, :H], qkv[:ML, H:2 * H], qkv[:ML, 2 * H:], ctx[ML:], lse_v, dctx[ML:], dqkv[ML:, :H], dqkv[:ML, H:2 * H],
                     dqkv[:ML, 2 * H:], B, V, L, run.heads, lang_mask, d_att_v, dbias=db)
    else:
        dx = torch.empty_like(xcat)
        dx[:ML].copy_(dy[:ML])
        dt, _ = ops.ln_bwd(dy[ML:], z, mean, rstd, ln.weight, A.grad(ln.weight), A.grad(ln.bias), A.grad(dense.bias), drop=d_hid, dres_out=dx[ML:])
        _wgrad(run, dt, ctx[ML:], dense.weight)
        dctx_v = ops.gemm(dt, A.w16(dense.weight), b_mn=True)
        dqkv = torch.zeros_like(qkv)
        ops.attn_bwd(qkv[ML:, :H], qkv[:ML, H:2 * H], qkv[:ML, 2 * H:], ctx[ML:], lse_v, dctx_v, dqkv[ML:, :H], dqkv[:ML, H:2 * H],
                     dqkv[:ML, 2 * H:], B, V, L, run.heads, lang_mask, d_att_v, dbias=db)
    if ws[0].requires_grad:
        _wgrad_fused(run, dqkv, xcat, ws)
    ops.gemm(dqkv, A.fused_w16(ws), b_mn=True, out=dx, accumulate=True)
    return dx


# ------------------------------------------------------------------------------------------------
# autograd wrappers
# ------------------------------------------------------------------------------------------------
def _as_bf16_2d(g: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    g = g.reshape(like.shape)
    if g.dtype != BF16:
        g = g.to(BF16)
    return g.contiguous()


class BertLayerFn(torch.autograd.Function):
    """BertLayer (vilmodel.py:188-201) on rows x [B*S, H]."""

    @staticmethod
    def forward(ctx, anchor, x, x32, run: Run, layer, B: int, S: int, mask):
        """Returns (y bf16, y32 fp32 twin -- not differentiable: the whole gradient travels through y)."""
        y1, y1_32, s1 = attn_block_fwd(run, x, B, S, mask, layer.attention.self, layer.attention.output, x32=x32)
        y2, y2_32, s2 = ffn_block_fwd(run, y1, layer.intermediate, layer.output, x32=y1_32)
        ctx.run, ctx.layer, ctx.s1, ctx.s2 = run, layer, s1, s2
        run.used(layer)
        ctx.mark_non_differentiable(y2_32)
        return y2, y2_32

    @staticmethod
    def backward(ctx, dy, _dy32=None):
        run, layer = ctx.run, ctx.layer
        dy = _as_bf16_2d(dy, ctx.s2[0])
        d1 = ffn_block_bwd(run, dy, ctx.s2, layer.intermediate, layer.output)
        dx = attn_block_bwd(run, d1, ctx.s1, layer.attention.self, layer.attention.output)
        ctx.s1 = ctx.s2 = None
        run.done(layer)
        return None, dx, None, None, None, None, None, None


class XLayerFn(torch.autograd.Function):
    """LXRTXLayer (vilmodel.py:362-412) on the joint buffer xcat = [language rows ; vision rows]."""

    @staticmethod
    def forward(ctx, anchor, xcat, xcat32, run: Run, layer, B: int, L: int, V: int, lang_mask, visn_mask, lang_ca: bool):
        ML = B * L
        y0, y0_32, s0 = cross_block_fwd(run, xcat, B, L, V, lang_mask, visn_mask, layer.visual_attention, lang_ca, xcat32=xcat32)
        y1, y1_32 = torch.empty_like(y0), _new32(y0)
        out, out32 = torch.empty_like(y0), _new32(y0)
        if lang_ca:
            _, _, sl = attn_block_fwd(run, y0[:ML], B, L, lang_mask, layer.lang_self_att.self, layer.lang_self_att.output, y_out=y1[:ML],
                                      x32=y0_32[:ML], y32_out=y1_32[:ML])
            _, _, fl = ffn_block_fwd(run, y1[:ML], layer.lang_inter, layer.lang_output, y_out=out[:ML], x32=y1_32[:ML], y32_out=out32[:ML])
        else:
            sl = fl = None
            out[:ML].copy_(y0[:ML])
            out32[:ML].copy_(y0_32[:ML])
        _, _, sv = attn_block_fwd(run, y0[ML:], B, V, visn_mask, layer.visn_self_att.self, layer.visn_self_att.output, y_out=y1[ML:],
                                  x32=y0_32[ML:], y32_out=y1_32[ML:])
        _, _, fv = ffn_block_fwd(run, y1[ML:], layer.visn_inter, layer.visn_output, y_out=out[ML:], x32=y1_32[ML:], y32_out=out32[ML:])
        ctx.run, ctx.layer, ctx.saved, ctx.ML, ctx.lang_ca = run, layer, (s0, sl, fl, sv, fv), ML, lang_ca
        run.used(layer)
        ctx.mark_non_differentiable(out32)
        return out, out32

    @staticmethod
    def backward(ctx, dout, _dout32=None):
        run, layer, ML = ctx.run, ctx.layer, ctx.ML
        s0, sl, fl, sv, fv = ctx.saved
        dout = _as_bf16_2d(dout, s0[0])
        d1 = torch.empty_like(dout)
        d0 = torch.empty_like(dout)
        if ctx.lang_ca:
            ffn_block_bwd(run, dout[:ML], fl, layer.lang_inter, layer.lang_output, dx_out=d1[:ML])
            attn_block_bwd(run, d1[:ML], sl, layer.lang_self_att.self, layer.lang_self_att.output, dx_out=d0[:ML])
        else:
            d0[:ML].copy_(dout[:ML])
        ffn_block_bwd(run, dout[ML:], fv, layer.visn_inter, layer.visn_output, dx_out=d1[ML:])
        attn_block_bwd(run, d1[ML:], sv, layer.visn_self_att.self, layer.visn_self_att.output, dx_out=d0[ML:])
        dx = cross_block_bwd(run, d0, s0, layer.visual_attention)
        ctx.saved = None
        run.done(layer)
        return (None, dx) + (None,) * 9


class LangSelfFn(torch.autograd.Function):
    """lang_self_att + lang_inter + lang_output of an x-layer, used by the finetune 'language' mode with
    no_lang_ca (vilmodel_cmt.py:645-652)."""

    @staticmethod
    def forward(ctx, anchor, x, x32, run: Run, layer, B: int, L: int, mask):
        y1, y1_32, s1 = attn_block_fwd(run, x, B, L, mask, layer.lang_self_att.self, layer.lang_self_att.output, x32=x32)
        y2, _, s2 = ffn_block_fwd(run, y1, layer.lang_inter, layer.lang_output, x32=y1_32)
        ctx.run, ctx.layer, ctx.s1, ctx.s2 = run, layer, s1, s2
        return y2

    @staticmethod
    def backward(ctx, dy):
        run, layer = ctx.run, ctx.layer
        dy = _as_bf16_2d(dy, ctx.s2[0])
        d1 = ffn_block_bwd(run, dy, ctx.s2, layer.lang_inter, layer.lang_output)
        dx = attn_block_bwd(run, d1, ctx.s1, layer.lang_self_att.self, layer.lang_self_att.output)
        return None, dx, None, None, None, None, None, None


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) for the heads / feature projections.  x bf16 [M,K]; out bf16 (or fp32 logits)."""

    @staticmethod
    def forward(ctx, anchor, x, run: Run, lin, act: int, out_f32: bool, need_dx: bool, w_override=None):
        A = run.arena
        weight = lin.weight if w_override is None else w_override
        bias = getattr(lin, "bias", None)
        w16 = A.w16(weight)
        pre = None
        if act != ops.ACT_NONE and run.save:
            pre = torch.empty((x.shape[0], weight.shape[0]), dtype=BF16, device=x.device)
            y = ops.gemm(x, w16, bias=bias, act=act, aux_mode=ops.AUX_STORE_PRE, aux=pre)
        else:
            out = None
            if out_f32 and weight.shape[0] % 4:
                # fp32 logits with an odd class count (30522): pad the row pitch so the epilogue can use 128-bit stores
                N = weight.shape[0]
                out = torch.empty((x.shape[0], (N + 7) // 8 * 8), dtype=torch.float32, device=x.device)[:, :N]
            y = ops.gemm(x, w16, bias=bias, act=act, out=out, out_dtype=torch.float32 if out_f32 else BF16)
        ctx.run, ctx.weight, ctx.bias, ctx.x, ctx.pre, ctx.act, ctx.need_dx = run, weight, bias, x, pre, act, need_dx
        return y

    @staticmethod
    def backward(ctx, dy):
        run, A = ctx.run, ctx.run.arena
        weight, bias, x, pre, act = ctx.weight, ctx.bias, ctx.x, ctx.pre, ctx.act
        if dy.dtype != BF16:
            dy = dy.to(BF16)
        if dy.stride(-1) != 1 or (dy.stride(0) * 2) % 16 != 0:
            # TMA operands need a 16-byte row pitch: pad the class dimension (zeros) once
            N = dy.shape[1]
            pad = torch.zeros((dy.shape[0], (N + 7) // 8 * 8), dtype=BF16, device=dy.device)
            pad[:, :N] = dy
            dy = pad[:, :N]
        if act == ops.ACT_GELU:
            dy = _mul_dact(dy, pre, ops.AUX_MUL_DGELU)
        elif act == ops.ACT_RELU:
            dy = _mul_dact(dy, pre, ops.AUX_MUL_DRELU)
        _wgrad(run, dy, x, weight)
        _bgrad(A, dy, bias)
        dx = None
        if ctx.need_dx:
            if weight.shape[0] >= 8192 and dy.shape[0] <= 2048:
                # tied MLM decoder (pretrain_cmt.py:96-99): reduction over 30 522 classes with only ~20 output tiles -- split-K into an
                # fp32 scratch (red.global.add) and one small cast instead of 21 CTAs walking 477 k-blocks each (122 us -> ~30 us)
                dx = ops.gemm(dy, A.w16(weight), b_mn=True, out_dtype=F32, accumulate=True).to(BF16)
            else:
                dx = ops.gemm(dy, A.w16(weight), b_mn=True)
        ctx.x = ctx.pre = None
        return None, dx, None, None, None, None, None, None


def _mul_dact(dy, pre, mode):
    """dy * act'(pre) as a torch elementwise op on small head tensors (M <= a few thousand rows)."""
    if mode == ops.AUX_MUL_DRELU:
        return (dy * (pre > 0)).contiguous()
    p = pre.float()
    cdf = 0.5 * (1.0 + torch.erf(p * 0.7071067811865476))
    pdf = 0.3989422804014327 * torch.exp(-0.5 * p * p)
    return (dy.float() * (cdf + p * pdf)).to(BF16)


class LayerNormFn(torch.autograd.Function):
    """Plain LayerNorm (+ optional dropout AFTER the norm is handled by the caller) for the heads."""

    @staticmethod
    def forward(ctx, anchor, x, run: Run, ln):
        y, _, mean, rstd = ops.ln_fwd(x, None, ln.weight, ln.bias, run.eps, save_z=True, inplace_z=True)
        ctx.run, ctx.ln, ctx.saved = run, ln, (x, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        A, ln = ctx.run.arena, ctx.ln
        x, mean, rstd = ctx.saved
        dy = _as_bf16_2d(dy, x)
        dx, _ = ops.ln_bwd(dy, x, mean, rstd, ln.weight, A.grad(ln.weight), A.grad(ln.bias), None, want_dres=False)
        return None, dx, None, None


class RowdotFn(torch.autograd.Function):
    """Final Linear(H -> N<=4) of a head, fp32 logits."""

    @staticmethod
    def forward(ctx, anchor, x, run: Run, lin):
        ctx.run, ctx.lin, ctx.x = run, lin, x
        return ops.rowdot_fwd(x, lin.weight, lin.bias)

    @staticmethod
    def backward(ctx, dy):
        A, lin = ctx.run.arena, ctx.lin
        dx = ops.rowdot_bwd(dy.float().contiguous(), ctx.x, lin.weight, A.grad(lin.weight), A.grad(lin.bias) if lin.bias is not None else None)
        return None, dx, None, None


class CrossEntropyFn(torch.autograd.Function):
    """F.cross_entropy(reduction='none') on fp32 logits (-inf entries allowed)."""

    @staticmethod
    def forward(ctx, logits, labels):
        if logits.stride(1) != 1:
            logits = logits.contiguous()
        loss, lse = ops.ce_fwd(logits, labels)
        ctx.saved = (logits, labels, lse)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        logits, labels, lse = ctx.saved
        wide = logits.shape[1] > 4096     # MLM: hand the decoder GEMMs a bf16, TMA-aligned gradient directly
        d = ops.ce_bwd(logits, labels, lse, gloss.float().contiguous(), bf16_padded=wide)
        return (d[:, :logits.shape[1]] if wide else d), None


class TextEmbedFn(torch.autograd.Function):
    """BertEmbeddings (vilmodel.py:40-69)."""

    @staticmethod
    def forward(ctx, anchor, run: Run, emb, ids):
        d = run.drop(emb.dropout)
        typ0 = emb.token_type_embeddings.weight[0]
        out = ops.embed_text_fwd(ids, emb.word_embeddings.weight, emb.position_embeddings.weight, typ0, emb.LayerNorm.weight, emb.LayerNorm.bias,
                                 run.eps, d)
        ctx.run, ctx.emb, ctx.ids, ctx.d = run, emb, ids, d
        return out

    @staticmethod
    def backward(ctx, dy):
        A, emb = ctx.run.arena, ctx.emb
        dy = dy.to(BF16).contiguous()
        ctx.run._ensure_final_join()
        ctx.run.join_side()         # the tied MLM decoder's weight gradient (side stream) lands in the word-embedding gradient too
        ops.embed_text_bwd(dy, ctx.ids, emb.word_embeddings.weight, emb.position_embeddings.weight, emb.token_type_embeddings.weight[0],
                           emb.LayerNorm.weight, A.grad(emb.word_embeddings.weight), A.grad(emb.position_embeddings.weight),
                           A.grad(emb.token_type_embeddings.weight)[0], A.grad(emb.LayerNorm.weight), A.grad(emb.LayerNorm.bias), ctx.run.eps,
                           ctx.d)
        ctx.run._hook(emb)          # data-parallel: the embedding tables' gradients (94 MB) are final -- exchange them now
        return None, None, None, None


class FeatEmbedFn(torch.autograd.Function):
    """LN(img_linear(x)) + LN(ang_linear(a)) [+ ...] [-> LN -> dropout]; the image linear runs on the tensor cores,
    everything else is one fused row kernel (csrc/hamt_embed.cu).  ``P`` = dict of parameter handles."""

    @staticmethod
    def forward(ctx, anchor, extra, run: Run, P: dict, x16, ang, nav_ids, pos_ids, pos_mod: int, drop_mod):
        A = run.arena
        lin = P["img_linear"]
        t = ops.gemm(x16, A.w16(lin.weight), bias=lin.bias)
        d = run.drop(drop_mod) if drop_mod is not None else ops.NO_DROP
        kw = dict(add_vec=P.get("add_vec"), nav_table=P["nav_table"].weight if nav_ids is not None else None, nav_ids=nav_ids, extra=extra,
                  pos_table=P["pos_table"].weight if P.get("pos_table") is not None else None, pos_ids=pos_ids, pos_mod=pos_mod,
                  g_f=P["ln_f"].weight if P.get("ln_f") is not None else None, b_f=P["ln_f"].bias if P.get("ln_f") is not None else None)
        base = (P["ang_linear"].weight, P["ang_linear"].bias, P["ln_img"].weight, P["ln_img"].bias, P["ln_ang"].weight, P["ln_ang"].bias)
        out = ops.embed_feat_fwd(t, ang, *base, eps=run.eps, drop=d, **kw)
        ctx.run, ctx.P, ctx.saved, ctx.kw, ctx.base, ctx.d, ctx.has_extra = run, P, (x16, t, ang), kw, base, d, extra is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        run, P, A = ctx.run, ctx.P, ctx.run.arena
        x16, t, ang = ctx.saved
        dy = _as_bf16_2d(dy, t)
        lin = P["img_linear"]
        grads = dict(dw_ang=A.grad(P["ang_linear"].weight), db_ang=A.grad(P["ang_linear"].bias), dg_img=A.grad(P["ln_img"].weight),
                     db_img=A.grad(P["ln_img"].bias), dg_ang=A.grad(P["ln_ang"].weight), dbe_ang=A.grad(P["ln_ang"].bias),
                     db_lin=A.grad(lin.bias))
        if P.get("add_vec_grad") is not None:
            grads["dadd_vec"] = P["add_vec_grad"](A)
        if ctx.kw["nav_ids"] is not None:
            grads["dnav_table"] = A.grad(P["nav_table"].weight)
        if ctx.kw["pos_table"] is not None:
            grads["dpos_table"] = A.grad(P["pos_table"].weight)
        if ctx.kw["g_f"] is not None:
            grads["dg_f"], grads["db_f"] = A.grad(P["ln_f"].weight), A.grad(P["ln_f"].bias)
        dt, dextra = ops.embed_feat_bwd(dy, t, ang, *ctx.base, grads, eps=run.eps, drop=ctx.d, want_dextra=ctx.has_extra, **ctx.kw)
        _wgrad(run, dt, x16, lin.weight)
        # end-to-end stage: the features come from the ViT backbone and want their gradient (image_vilmodel.py:50-55)
        dx16 = ops.gemm(dt, A.w16(lin.weight), b_mn=True) if ctx.needs_input_grad[4] else None
        ctx.saved = None
        return (None, dextra, None, None, dx16) + (None,) * 5


class MeanPoolFn(torch.autograd.Function):
    """torch.mean over the P views of each panorama (vilmodel.py:563-564); bf16 [N*P,H] -> fp32 [N,H]."""

    @staticmethod
    def forward(ctx, x, N: int, P: int):
        ctx.N, ctx.P = N, P
        return ops.mean_pool_fwd(x, N, P)

    @staticmethod
    def backward(ctx, dy):
        return ops.mean_pool_bwd(dy.float().contiguous(), ctx.N, ctx.P), None, None


class RowLNFn(torch.autograd.Function):
    """drop(LN(x + res)) on small fp32->bf16 row sets (history CLS token, ITM re-positioned history steps)."""

    @staticmethod
    def forward(ctx, anchor, x, run: Run, ln, drop_mod):
        d = run.drop(drop_mod)
        # fp32 input: the bf16 part goes through the kernel's x slot, the remainder through its fp32 residual slot (summed in fp32)
        x16 = x.to(BF16).contiguous()
        lo = (x - x16.float()).contiguous()
        y, z, mean, rstd = ops.ln_fwd(x16, lo, ln.weight, ln.bias, run.eps, save_z=True, inplace_z=True)
        # dropout AFTER the norm: applied as ln_fwd(dropout(.)) is wrong here, so do it with the streaming kernel trick:
        ctx.run, ctx.ln, ctx.saved, ctx.d = run, ln, (z, mean, rstd), d
        if d.p > 0:
            mask = _post_drop_mask(y, d)
            y = (y * mask).contiguous()
            ctx.mask = mask
        else:
            ctx.mask = None
        return y

    @staticmethod
    def backward(ctx, dy):
        A, ln = ctx.run.arena, ctx.ln
        z, mean, rstd = ctx.saved
        dy = _as_bf16_2d(dy, z)
        if ctx.mask is not None:
            dy = (dy * ctx.mask).contiguous()
        dx, _ = ops.ln_bwd(dy, z, mean, rstd, ln.weight, A.grad(ln.weight), A.grad(ln.bias), None, want_dres=False)
        return None, dx.float(), None, None, None


def _post_drop_mask(y: torch.Tensor, d: ops.Drop) -> torch.Tensor:
    """Keep-mask * 1/(1-p) generated by the same device hash as every other dropout site: run the LN kernel on a
    tensor of ones with gamma=1, beta=0 and read the pre-norm z it saves (z = dropout(1))."""
    ones = torch.ones_like(y)
    g = torch.ones(y.shape[1], dtype=torch.float32, device=y.device)
    b = torch.zeros_like(g)
    _, z, _, _ = ops.ln_fwd(ones, None, g, b, 1e-12, d, save_z=True, inplace_z=True)
    return z


class MulRowsFn(torch.autograd.Function):
    """ob_embeds * txt_embeds[:, :1] (pretrain_cmt.py:176)."""

    @staticmethod
    def forward(ctx, a, v, B: int, S: int):
        ctx.saved, ctx.B, ctx.S = (a, v), B, S
        return ops.mul_rows(a.contiguous(), v.contiguous(), B, S)

    @staticmethod
    def backward(ctx, dy):
        a, v = ctx.saved
        B, S = ctx.B, ctx.S
        dy = dy.to(BF16).contiguous()
        da = ops.mul_rows(dy, v.contiguous(), B, S)
        dv = (dy.float().view(B, S, -1) * a.float().view(B, S, -1)).sum(1).to(BF16)
        return da, dv, None, None


class DropoutFn(torch.autograd.Function):
    """Stand-alone dropout (head MLPs: ... LN -> Dropout -> Linear, pretrain_cmt.py:16-20) with the device-hash mask."""

    @staticmethod
    def forward(ctx, x, d: ops.Drop):
        mask = _post_drop_mask(x, d)
        ctx.mask = mask
        return (x * mask).contiguous()

    @staticmethod
    def backward(ctx, dy):
        return (dy.to(BF16) * ctx.mask).contiguous(), None


def dropout(run: Run, x: torch.Tensor, module) -> torch.Tensor:
    d = run.drop(module)
    return x if d.p <= 0 else DropoutFn.apply(x, d)


class GatherRowsFn(torch.autograd.Function):
    """hidden[mask] (_compute_masked_hidden, pretrain_cmt.py:161-165) with precomputed row indices."""

    @staticmethod
    def forward(ctx, x, idx):
        ctx.idx, ctx.rows = idx, x.shape[0]
        return ops.gather_rows(x.contiguous(), idx)

    @staticmethod
    def backward(ctx, dy):
        return ops.scatter_rows(dy.to(BF16).contiguous(), ctx.idx, ctx.rows), None


# ------------------------------------------------------------------------------------------------
# end-to-end stage: ViT-B/16 backbone on the same kernels
# ------------------------------------------------------------------------------------------------
class CastFn(torch.autograd.Function):
    """fp32 -> bf16 copy of the view features with a gradient (end-to-end stage: the features come out of the ViT)."""

    @staticmethod
    def forward(ctx, x):
        return ops.cast_bf16(x)

    @staticmethod
    def backward(ctx, dy):
        return dy.float()


class VitFn(torch.autograd.Function):
    """VisionTransformer.forward_features (pretrain_src/model/vision_transformer.py:335-348): images fp32 [N,3,H,W] -> class-token
    features fp32 [N,E].

    Pre-LN blocks (x = x + attn(norm1(x)); x = x + mlp(norm2(x)), :195-198) map onto the fused dropout + residual + LayerNorm kernel
    shifted by one sublayer: the kernel that adds a sublayer's output to the fp32 residual stream also applies the NEXT sublayer's
    LayerNorm (ops.ln_fwd_prenorm: z = drop(t) + x -> new stream in fp32, y = LN_next(z) in bf16), so a block is
      qkv GEMM -> attention -> proj GEMM -> [add + norm2] -> fc1 GEMM (+GELU) -> fc2 GEMM -> [add + next norm1 / final norm]
    with no stand-alone residual-add or LayerNorm pass.  The backward walks the sublayers in reverse with the same fused kernels
    (ln_bwd adds the incoming stream gradient through its residual-gradient input)."""

    @staticmethod
    def forward(ctx, anchor, images, run: Run, vit):
        A = run.arena
        N = images.shape[0]
        pe = vit.patch_embed
        ps = pe.patch_size[0]
        S, E, heads = pe.num_patches + 1, vit.embed_dim, vit.num_heads
        eps = float(vit.norm.eps)            # 1e-6 (vision_transformer.py:265), not the BERT-side 1e-12 of an enclosing model's Run
        patches = ops.patchify(images, ps)                                                         # [N*196, 3*16*16] bf16
        w_pe = pe.proj.weight
        t0 = ops.gemm(patches, A.w16(w_pe).view(E, -1), bias=pe.proj.bias)                          # Conv2d(k = s = 16) as a GEMM
        d_pos = run.drop(vit.pos_drop)
        x32, x16 = ops.vit_embed_fwd(t0, vit.cls_token.view(-1), vit.pos_embed.view(-1), N, S, d_pos)
        blocks = list(vit.blocks)
        n0 = blocks[0].norm1 if blocks else vit.norm
        y, y32, _, _, mean0, rstd0 = ops.ln_fwd_prenorm(None, x32, n0.weight, n0.bias, eps, save=run.save, want_y32=not blocks)
        saved = []
        for i, blk in enumerate(blocks):
            at, mlp = blk.attn, blk.mlp
            qkv = ops.gemm(y, A.w16(at.qkv.weight), bias=at.qkv.bias)
            d_att = run.drop(at.attn_drop)
            c, lse = ops.attn_fwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], N, S, S, heads, None, d_att, need_lse=run.save)
            t = ops.gemm(c, A.w16(at.proj.weight), bias=at.proj.bias)
            d1 = run.drop(at.proj_drop)
            y2, _, z1, x32, mean1, rstd1 = ops.ln_fwd_prenorm(t, x32, blk.norm2.weight, blk.norm2.bias, eps, d1, save=run.save)
            d_mid = run.drop(mlp.drop)
            h = torch.empty((N * S, mlp.fc1.weight.shape[0]), dtype=BF16, device=images.device) if run.save else None
            a = ops.gemm(y2, A.w16(mlp.fc1.weight), bias=mlp.fc1.bias, act=ops.ACT_GELU, aux_mode=ops.AUX_STORE_DGELU if run.save else ops.AUX_NONE, aux=h)
            m_mid = None
            if d_mid.p > 0:        # Mlp.drop between GELU and fc2 (:148): mask from the device hash, applied as an elementwise pass
                m_mid = _post_drop_mask(a.view(-1, E), d_mid).view_as(a)
                a = a * m_mid
            t2 = ops.gemm(a, A.w16(mlp.fc2.weight), bias=mlp.fc2.bias)
            d2 = run.drop(mlp.drop)
            last = i + 1 == len(blocks)
            nn_ = vit.norm if last else blocks[i + 1].norm1
            yn, y32, z2, x32, mean2, rstd2 = ops.ln_fwd_prenorm(t2, x32, nn_.weight, nn_.bias, eps, d2, save=run.save, want_z32=not last, want_y32=last)
            if run.save:
                saved.append((y, qkv, c, lse, d_att, z1, mean1, rstd1, d1, y2, h, a, m_mid, z2, mean2, rstd2, d2))
            y = yn
        feats = y32.view(N, S, E)[:, 0].contiguous()                                                # fp32 class token after the final norm
        ctx.run, ctx.vit, ctx.N, ctx.S = run, vit, N, S
        ctx.saved = (patches, x16, mean0, rstd0, d_pos, saved) if run.save else None
        return feats

    @staticmethod
    def backward(ctx, dfeat):
        run, vit, N, S = ctx.run, ctx.vit, ctx.N, ctx.S
        A = run.arena
        patches, x16, mean0, rstd0, d_pos, saved = ctx.saved
        E, heads = vit.embed_dim, vit.num_heads
        blocks = list(vit.blocks)
        g_y = torch.zeros((N, S, E), dtype=BF16, device=dfeat.device)          # only the class token carries gradient out of the final norm
        g_y[:, 0] = dfeat.to(BF16)
        g_y = g_y.view(N * S, E)
        g_x = None                                                            # gradient wrt the residual stream behind the current point
        for i in range(len(blocks) - 1, -1, -1):
            blk = blocks[i]
            at, mlp = blk.attn, blk.mlp
            y, qkv, c, lse, d_att, z1, mean1, rstd1, d1, y2, h, a, m_mid, z2, mean2, rstd2, d2 = saved[i]
            nn_ = vit.norm if i + 1 == len(blocks) else blocks[i + 1].norm1
            # ---- mlp sublayer (+ the LayerNorm that followed it)
            dt2, g_x = ops.ln_bwd(g_y, z2, mean2, rstd2, nn_.weight, A.grad(nn_.weight), A.grad(nn_.bias), A.grad(mlp.fc2.bias), dres_in=g_x, drop=d2, prenorm=True)
            _wgrad(run, dt2, a, mlp.fc2.weight)
            # h holds gelu'(pre-activation) (forward epilogue); with Mlp.drop active the keep-mask is folded into it
            hd = h if m_mid is None else h * m_mid
            dh = ops.gemm(dt2, A.w16(mlp.fc2.weight), b_mn=True, aux_mode=ops.AUX_MUL, aux=hd, colsum=A.grad(mlp.fc1.bias))
            _wgrad(run, dh, y2, mlp.fc1.weight)
            g_y = ops.gemm(dh, A.w16(mlp.fc1.weight), b_mn=True)
            # ---- attention sublayer (+ norm2)
            dt, g_x = ops.ln_bwd(g_y, z1, mean1, rstd1, blk.norm2.weight, A.grad(blk.norm2.weight), A.grad(blk.norm2.bias), A.grad(at.proj.bias),
                                 dres_in=g_x, drop=d1, prenorm=True)
            _wgrad(run, dt, c, at.proj.weight)
            dc = ops.gemm(dt, A.w16(at.proj.weight), b_mn=True)
            dqkv = torch.empty_like(qkv)
            ops.attn_bwd(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], c, lse, dc, dqkv[:, :E], dqkv[:, E:2 * E], dqkv[:, 2 * E:], N, S, S, heads, None,
                         d_att, dbias=A.grad(at.qkv.bias))
            _wgrad(run, dqkv, y, at.qkv.weight)
            g_y = ops.gemm(dqkv, A.w16(at.qkv.weight), b_mn=True)
            saved[i] = None
        n0 = blocks[0].norm1 if blocks else vit.norm
        _, g_x = ops.ln_bwd(g_y, x16, mean0, rstd0, n0.weight, A.grad(n0.weight), A.grad(n0.bias), None, dres_in=g_x, want_dx=False)
        dfull, dt0 = ops.vit_embed_bwd(g_x, N, S, d_pos)
        col = torch.zeros(S * E, dtype=F32, device=dfeat.device)
        ops.colsum(dfull.view(N, S * E), col)                                  # sum over the images: d pos_embed; its first row is d cls_token too
        A.grad(vit.pos_embed).view(-1).add_(col)
        A.grad(vit.cls_token).view(-1).add_(col[:E])
        pe = vit.patch_embed
        g_w = A.grad(pe.proj.weight).view(E, -1)
        run.fork_wgrad(lambda: ops.gemm(dt0, patches, a_mn=True, b_mn=True, out=g_w, accumulate=True), dt0, patches)
        ops.colsum(dt0, A.grad(pe.proj.bias))
        ctx.saved = None
        return None, None, None, None
