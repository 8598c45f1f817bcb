"""Raw (non-autograd) op wrappers: torch tensors in, hand-written sm_100a kernels out.

Every function enqueues on ``torch.cuda.current_stream()`` and returns torch tensors allocated by
the PyTorch caching allocator (torch = device memory + streams; the arithmetic is in
csrc/*.cu).  No function here has an eager / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class Drop:
    """Dropout descriptor: device seed pointer + call-site id + probability."""
    __slots__ = ("seed", "site", "p")

    def __init__(self, seed: Optional[torch.Tensor], site: int, p: float):
        self.seed, self.site, self.p = seed, int(site) & 0xFFFFFFFF, float(p)

    @property
    def args(self):
        if self.p <= 0.0 or self.seed is None:
            return (None, 0, 0.0)
        return (self.seed.data_ptr(), self.site, self.p)


NO_DROP = Drop(None, 0, 0.0)

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
AUX_NONE, AUX_STORE_PRE, AUX_MUL_DGELU, AUX_MUL_DRELU, AUX_STORE_DGELU, AUX_MUL = 0, 1, 2, 3, 4, 5


def _check_bf16_2d(t: torch.Tensor, name: str):
    if t.dtype != BF16 or t.dim() != 2 or t.stride(1) != 1 or not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA bf16 2-D tensor with unit inner stride, got {t.dtype} {tuple(t.shape)} {t.stride()}")


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, bias: Optional[torch.Tensor] = None,
         act: int = ACT_NONE, aux_mode: int = AUX_NONE, aux: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         out_dtype=BF16, accumulate: bool = False, alpha: float = 1.0, tile_n: int = 0, splits: int = 0,
         colsum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """D[M,N] = act(alpha * sum_k A(m,k) B(n,k) + bias).  a: [M,K] (or [K,M] when a_mn), b: [N,K] (or [K,N] when b_mn).
    colsum (fp32 [N]): += column sums of the stored bf16 D (fused bias gradient)."""
    _check_bf16_2d(a, "gemm.a")
    _check_bf16_2d(b, "gemm.b")
    M, K = (a.shape[1], a.shape[0]) if a_mn else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else b.shape
    if K != Kb:
        raise ValueError(f"gemm: reduction mismatch {K} vs {Kb}")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
        if accumulate:
            out.zero_()
    if out.shape != (M, N) or out.stride(1) != 1 or out.dtype not in (BF16, F32):
        raise ValueError("gemm: bad output tensor")
    out_f32 = 1 if out.dtype == F32 else 0
    if bias is not None and (bias.dtype != F32 or bias.numel() != N):
        raise ValueError("gemm: bias must be fp32 [N]")
    if aux_mode != AUX_NONE:
        _check_bf16_2d(aux, "gemm.aux")
    if colsum is not None and (colsum.dtype != F32 or colsum.numel() != N or not colsum.is_contiguous()):
        raise ValueError("gemm: colsum must be contiguous fp32 [N]")
    lib = _lib.load()
    rc = lib.hamt_gemm_bf16(a.data_ptr(), int(a_mn), a.stride(0), b.data_ptr(), int(b_mn), b.stride(0), out.data_ptr(), out.stride(0),
                            out_f32, (2 if out_f32 else 1) if accumulate else 0, M, N, K, _ptr(bias), act, aux_mode, _ptr(aux),
                            aux.stride(0) if aux is not None else 0, alpha, tile_n, splits, _ptr(colsum), _stream())
    _lib.check(rc, "gemm_bf16")
    return out


def ln_fwd(x, res, gamma, beta, eps: float, drop: Drop = NO_DROP, save_z: bool = True, inplace_z: bool = True, out=None, out32=None):
    """y = LN(dropout(x) + res).  Returns (y, z, mean, rstd); z aliases x when inplace_z.  `res` may be bf16 or fp32 (the
    full-precision residual stream); `out32` (fp32 [M,H]) additionally receives the un-rounded y for the next residual add."""
    M, H = x.shape
    y = torch.empty_like(x) if out is None else out
    res16 = res if (res is not None and res.dtype == BF16) else None
    res32 = res if (res is not None and res.dtype == F32) else None
    if res is not None and (res.shape != x.shape or res.stride(1) != 1 or res.stride(0) != H):
        raise ValueError("ln_fwd: the residual must be a contiguous [M, H] tensor")
    if out32 is not None and (out32.dtype != F32 or out32.shape != x.shape or out32.stride(0) != H):
        raise ValueError("ln_fwd: out32 must be contiguous fp32 [M, H]")
    z = (x if inplace_z else torch.empty_like(x)) if save_z else None
    mean = torch.empty(M, dtype=F32, device=x.device) if save_z else None
    rstd = torch.empty(M, dtype=F32, device=x.device) if save_z else None
    sp, site, p = drop.args
    rc = _lib.load().hamt_ln_fwd(x.data_ptr(), _ptr(res16), _ptr(res32), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _ptr(out32), _ptr(z),
                                 _ptr(mean), _ptr(rstd), M, H, eps, sp, site, p, _stream())
    _lib.check(rc, "ln_fwd")
    return y, z, mean, rstd


def ln_bwd(dy, z, mean, rstd, gamma, dgamma, dbeta, dbias=None, dres_in=None, want_dx: bool = True, want_dres: bool = True, drop: Drop = NO_DROP,
           dres_out=None, prenorm: bool = False):
    """Returns (dx, dres); column sums are accumulated into dgamma / dbeta / dbias (fp32).  prenorm (ViT blocks): z is the residual
    stream, dx = mask o (dz + dres_in)."""
    M, H = dy.shape
    dx = torch.empty_like(dy) if want_dx else None
    dres = (torch.empty_like(dy) if dres_out is None else dres_out) if want_dres else None
    sp, site, p = drop.args
    lib = _lib.load()
    rc = (lib.hamt_ln_bwd_prenorm if prenorm else lib.hamt_ln_bwd)(dy.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), _ptr(dres_in), _ptr(dx), _ptr(dres),
                                 _ptr(dgamma), _ptr(dbeta), _ptr(dbias), M, H, sp, site, p, _stream())
    _lib.check(rc, "ln_bwd")
    return dx, dres


def attn_fwd(q, k, v, B: int, Sq: int, Sk: int, heads: int, mask: Optional[torch.Tensor], drop: Drop = NO_DROP, need_lse: bool = True, out=None):
    """q: view [B*Sq, heads*64] (row pitch = stride(0)), k/v: views [B*Sk, heads*64].  Returns (ctx [B*Sq, heads*64], lse)."""
    Hd = heads * 64
    if out is None:
        out = torch.empty((B * Sq, Hd), dtype=BF16, device=q.device)
    lse = torch.empty((B, heads, Sq), dtype=F32, device=q.device) if need_lse else None
    sp, site, p = drop.args
    rc = _lib.load().hamt_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), Sq * q.stride(0), Sk * k.stride(0), q.stride(0), k.stride(0),
                                   _ptr(mask), out.data_ptr(), out.stride(0), Sq * out.stride(0), _ptr(lse), B, heads, Sq, Sk, 0.125, sp, site, p,
                                   _stream())
    _lib.check(rc, "attn_fwd")
    return out, lse


def attn_bwd(q, k, v, out, lse, dout, dq, dk, dv, B: int, Sq: int, Sk: int, heads: int, mask, drop: Drop = NO_DROP, dbias=None):
    """dq/dk/dv are written through views with the same pitches as q/k/v.  dbias: fp32 [3 * heads * 64] (q | k | v bias gradients, e.g.
    arena.fused_grad of the three biases): the column sums of dq/dk/dv are ACCUMULATED into it by the kernel."""
    assert dq.stride(0) == q.stride(0) and dk.stride(0) == k.stride(0) and dv.stride(0) == v.stride(0)
    Hd = heads * 64
    sp, site, p = drop.args
    db = (None, None, None)
    if dbias is not None:
        if dbias.dtype != F32 or dbias.numel() != 3 * Hd or not dbias.is_contiguous():
            raise ValueError("attn_bwd: dbias must be contiguous fp32 [3 * heads * 64]")
        db = (dbias.data_ptr(), dbias.data_ptr() + 4 * Hd, dbias.data_ptr() + 8 * Hd)
    rc = _lib.load().hamt_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), Sq * q.stride(0), Sk * k.stride(0), q.stride(0), k.stride(0), _ptr(mask),
                                   out.data_ptr(), out.stride(0), Sq * out.stride(0), lse.data_ptr(), dout.data_ptr(), dout.stride(0),
                                   Sq * dout.stride(0), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, heads, Sq, Sk, 0.125, sp, site, p, *db, _stream())
    _lib.check(rc, "attn_bwd")


def embed_text_fwd(ids, word, pos, type0, gamma, beta, eps, drop: Drop = NO_DROP):
    B, L = ids.shape
    H = word.shape[1]
    out = torch.empty((B * L, H), dtype=BF16, device=ids.device)
    sp, site, p = drop.args
    rc = _lib.load().hamt_embed_text_fwd(ids.data_ptr(), word.data_ptr(), pos.data_ptr(), type0.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                         out.data_ptr(), B, L, H, eps, sp, site, p, _stream())
    _lib.check(rc, "embed_text_fwd")
    return out


def embed_text_bwd(dy, ids, word, pos, type0, gamma, dword, dpos, dtype0, dgamma, dbeta, eps, drop: Drop = NO_DROP):
    B, L = ids.shape
    H = word.shape[1]
    sp, site, p = drop.args
    rc = _lib.load().hamt_embed_text_bwd(dy.data_ptr(), ids.data_ptr(), word.data_ptr(), pos.data_ptr(), type0.data_ptr(), gamma.data_ptr(),
                                         dword.data_ptr(), dpos.data_ptr(), dtype0.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), B, L, H, eps,
                                         sp, site, p, _stream())
    _lib.check(rc, "embed_text_bwd")


def _feat_desc(t, ang, w_ang, b_ang, g_img, b_img, g_ang, be_ang, add_vec, nav_table, nav_ids, extra, pos_table, pos_ids, pos_mod, g_f, b_f,
               out, eps, drop: Drop):
    sp, site, p = drop.args
    M, H = t.shape
    return _lib.EmbedFeatDesc(t.data_ptr(), ang.data_ptr(), ang.shape[-1], w_ang.data_ptr(), b_ang.data_ptr(), g_img.data_ptr(), b_img.data_ptr(),
                              g_ang.data_ptr(), be_ang.data_ptr(), _ptr(add_vec), _ptr(nav_table), _ptr(nav_ids), _ptr(extra), _ptr(pos_table),
                              _ptr(pos_ids), int(pos_mod), _ptr(g_f), _ptr(b_f), _ptr(out), M, H, eps, sp, site, p)


def embed_feat_fwd(t, ang, w_ang, b_ang, g_img, b_img, g_ang, be_ang, *, add_vec=None, nav_table=None, nav_ids=None, extra=None, pos_table=None,
                   pos_ids=None, pos_mod=1, g_f=None, b_f=None, eps=1e-12, drop: Drop = NO_DROP):
    out = torch.empty_like(t)
    d = _feat_desc(t, ang, w_ang, b_ang, g_img, b_img, g_ang, be_ang, add_vec, nav_table, nav_ids, extra, pos_table, pos_ids, pos_mod, g_f, b_f, out,
                   eps, drop)
    rc = _lib.load().hamt_embed_feat_fwd(C.byref(d), _stream())
    _lib.check(rc, "embed_feat_fwd")
    return out


def embed_feat_bwd(dy, t, ang, w_ang, b_ang, g_img, b_img, g_ang, be_ang, grads: dict, *, add_vec=None, nav_table=None, nav_ids=None, extra=None,
                   pos_table=None, pos_ids=None, pos_mod=1, g_f=None, b_f=None, eps=1e-12, drop: Drop = NO_DROP, want_dextra: bool = False):
    """grads: dict of fp32 accumulation buffers keyed dw_ang, db_ang, dg_img, db_img, dg_ang, dbe_ang[, dadd_vec, dnav_table, dpos_table,
    dg_f, db_f, db_lin].  Returns (dt bf16, dextra fp32 or None)."""
    dt = torch.empty_like(t)
    dextra = torch.empty(t.shape, dtype=F32, device=t.device) if want_dextra else None
    d = _feat_desc(t, ang, w_ang, b_ang, g_img, b_img, g_ang, be_ang, add_vec, nav_table, nav_ids, extra, pos_table, pos_ids, pos_mod, g_f, b_f, None,
                   eps, drop)
    g = _lib.EmbedFeatGrads(dy.data_ptr(), dt.data_ptr(), *[_ptr(grads.get(k)) for k in ("dw_ang", "db_ang", "dg_img", "db_img", "dg_ang", "dbe_ang",
                                                                                           "dadd_vec", "dnav_table")],
                            _ptr(dextra), *[_ptr(grads.get(k)) for k in ("dpos_table", "dg_f", "db_f", "db_lin")])
    rc = _lib.load().hamt_embed_feat_bwd(C.byref(d), C.byref(g), _stream())
    _lib.check(rc, "embed_feat_bwd")
    return dt, dextra


def cast_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 -> bf16 copy (contiguous)."""
    if x.dtype == BF16:
        return x
    x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=BF16, device=x.device)
    rc = _lib.load().hamt_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    _lib.check(rc, "cast_f32_to_bf16")
    return out


def colsum(x: torch.Tensor, out: torch.Tensor):
    """out[n] += sum_m x[m,n]  (x bf16 view with unit inner stride)."""
    M, N = x.shape
    rc = _lib.load().hamt_colsum_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), M, N, _stream())
    _lib.check(rc, "colsum_bf16")


def mean_pool_fwd(x: torch.Tensor, N: int, P: int) -> torch.Tensor:
    H = x.shape[-1]
    out = torch.empty((N, H), dtype=F32, device=x.device)
    _lib.check(_lib.load().hamt_mean_pool_fwd(x.data_ptr(), out.data_ptr(), N, P, H, _stream()), "mean_pool_fwd")
    return out


def mean_pool_bwd(dy: torch.Tensor, N: int, P: int) -> torch.Tensor:
    H = dy.shape[-1]
    dx = torch.empty((N * P, H), dtype=BF16, device=dy.device)
    _lib.check(_lib.load().hamt_mean_pool_bwd(dy.data_ptr(), dx.data_ptr(), N, P, H, _stream()), "mean_pool_bwd")
    return dx


def add(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is None:
        out = torch.empty_like(a)
    _lib.check(_lib.load().hamt_add_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), "add_bf16")
    return out


def mul_rows(a: torch.Tensor, v: torch.Tensor, B: int, S: int) -> torch.Tensor:
    """out[b,s,:] = a[b,s,:] * v[b,:]; a: [B*S,H], v: [B,H] (contiguous bf16)."""
    out = torch.empty_like(a)
    _lib.check(_lib.load().hamt_mul_rows_bf16(a.data_ptr(), v.data_ptr(), out.data_ptr(), B, S, a.shape[-1], _stream()), "mul_rows_bf16")
    return out


def rowdot_fwd(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    M, H = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=F32, device=x.device)
    _lib.check(_lib.load().hamt_rowdot_fwd(x.data_ptr(), w.data_ptr(), _ptr(b), y.data_ptr(), M, N, H, _stream()), "rowdot_fwd")
    return y


def rowdot_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, dw: torch.Tensor, db: Optional[torch.Tensor], want_dx: bool = True):
    M, H = x.shape
    N = w.shape[0]
    dx = torch.empty_like(x) if want_dx else None
    _lib.check(_lib.load().hamt_rowdot_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), _ptr(dx), dw.data_ptr(), _ptr(db), M, N, H, _stream()), "rowdot_bwd")
    return dx


def ce_fwd(logits: torch.Tensor, labels: torch.Tensor):
    M, N = logits.shape
    loss = torch.empty(M, dtype=F32, device=logits.device)
    lse = torch.empty(M, dtype=F32, device=logits.device)
    _lib.check(_lib.load().hamt_ce_fwd(logits.data_ptr(), logits.stride(0), labels.data_ptr(), loss.data_ptr(), lse.data_ptr(), M, N, _stream()), "ce_fwd")
    return loss, lse


def ce_bwd(logits, labels, lse, gloss, bf16_padded: bool = False):
    M, N = logits.shape
    if bf16_padded:
        ld = (N + 7) // 8 * 8
        d = torch.empty((M, ld), dtype=BF16, device=logits.device)
        rc = _lib.load().hamt_ce_bwd(logits.data_ptr(), logits.stride(0), labels.data_ptr(), lse.data_ptr(), gloss.data_ptr(), None, d.data_ptr(), ld, M,
                                     N, _stream())
    else:
        d = torch.empty((M, N), dtype=F32, device=logits.device)
        rc = _lib.load().hamt_ce_bwd(logits.data_ptr(), logits.stride(0), labels.data_ptr(), lse.data_ptr(), gloss.data_ptr(), d.data_ptr(), None, N, M,
                                     N, _stream())
    _lib.check(rc, "ce_bwd")
    return d


def gather_rows(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    n, H = idx.numel(), x.shape[-1]
    out = torch.empty((n, H), dtype=BF16, device=x.device)
    _lib.check(_lib.load().hamt_gather_rows_bf16(x.data_ptr(), idx.data_ptr(), out.data_ptr(), n, H, _stream()), "gather_rows")
    return out


def scatter_rows(x: torch.Tensor, idx: torch.Tensor, rows: int) -> torch.Tensor:
    n, H = idx.numel(), x.shape[-1]
    out = torch.zeros((rows, H), dtype=BF16, device=x.device)
    _lib.check(_lib.load().hamt_scatter_rows_bf16(x.data_ptr(), idx.data_ptr(), out.data_ptr(), n, H, _stream()), "scatter_rows")
    return out


def gather_rows_pad(table: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[i] = table[idx[i]] if idx[i] >= 0 else 0.  table: bf16 [R, H] contiguous, idx: int64 (any shape) on the same device.
    The caller guarantees idx < R (feature_store checks on the host where the indices are built)."""
    if table.dtype != BF16 or table.dim() != 2 or not table.is_contiguous():
        raise ValueError("gather_rows_pad: table must be contiguous bf16 [R, H]")
    if idx.dtype != torch.int64 or idx.device != table.device or not idx.is_contiguous():
        raise ValueError("gather_rows_pad: idx must be a contiguous int64 tensor on the table's device")
    n, H = idx.numel(), table.shape[1]
    if out is None:
        out = torch.empty((n, H), dtype=BF16, device=table.device)
    elif out.dtype != BF16 or out.numel() != n * H or not out.is_contiguous():
        raise ValueError("gather_rows_pad: out must be contiguous bf16 with n * H elements")
    _lib.check(_lib.load().hamt_gather_rows_pad_bf16(table.data_ptr(), table.shape[0], idx.data_ptr(), out.data_ptr(), n, H, _stream()), "gather_rows_pad")
    return out


# ---- end-to-end ViT stage ------------------------------------------------------------------------------------------
def ln_fwd_prenorm(x, res32, gamma, beta, eps: float, drop: Drop = NO_DROP, save: bool = True, want_z32: bool = True, want_y32: bool = False):
    """Pre-LN residual step: z = dropout(x) + res32 (the new fp32 residual stream), y = LN(z).  x (bf16 [M,H]) may be None (z = res32).
    Returns (y bf16, y32 or None, z16 (aliases x; None when x is None or not saved), z32 or None, mean, rstd)."""
    M, H = res32.shape
    if res32.dtype != F32 or not res32.is_contiguous() or (x is not None and (x.shape != res32.shape or not x.is_contiguous())):
        raise ValueError("ln_fwd_prenorm: res32 must be contiguous fp32 [M, H] and x a contiguous bf16 tensor of the same shape")
    y = torch.empty((M, H), dtype=BF16, device=res32.device)
    y32 = torch.empty((M, H), dtype=F32, device=res32.device) if want_y32 else None
    z32 = torch.empty((M, H), dtype=F32, device=res32.device) if (want_z32 and x is not None) else None
    z16 = x if (save and x is not None) else None
    mean = torch.empty(M, dtype=F32, device=res32.device) if save else None
    rstd = torch.empty(M, dtype=F32, device=res32.device) if save else None
    sp, site, p = drop.args
    rc = _lib.load().hamt_ln_fwd_prenorm(_ptr(x), res32.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _ptr(y32), _ptr(z16), _ptr(z32),
                                         _ptr(mean), _ptr(rstd), M, H, eps, sp, site, p, _stream())
    _lib.check(rc, "ln_fwd_prenorm")
    return y, y32, z16, z32, mean, rstd


def patchify(images: torch.Tensor, patch: int) -> torch.Tensor:
    """fp32 [N, C, H, W] -> bf16 [N * (H/patch) * (W/patch), C * patch * patch] (columns = channel, row, column)."""
    if images.dtype != F32 or images.dim() != 4 or not images.is_contiguous() or not images.is_cuda:
        raise ValueError("patchify: expected a contiguous CUDA fp32 [N, C, H, W] tensor")
    N, C, Hh, Ww = images.shape
    out = torch.empty((N * (Hh // patch) * (Ww // patch), C * patch * patch), dtype=BF16, device=images.device)
    _lib.check(_lib.load().hamt_patchify_bf16(images.data_ptr(), out.data_ptr(), N, C, Hh, Ww, patch, _stream()), "patchify_bf16")
    return out


def vit_embed_fwd(t0: torch.Tensor, cls: torch.Tensor, pos: torch.Tensor, N: int, S: int, drop: Drop = NO_DROP):
    """x = pos_drop(cat(cls, tokens) + pos): t0 bf16 [N*(S-1), H], cls fp32 [H], pos fp32 [S*H] -> (x32 fp32, x16 bf16) [N*S, H]."""
    H = t0.shape[1]
    x32 = torch.empty((N * S, H), dtype=F32, device=t0.device)
    x16 = torch.empty((N * S, H), dtype=BF16, device=t0.device)
    sp, site, p = drop.args
    rc = _lib.load().hamt_vit_embed_fwd(t0.data_ptr(), cls.data_ptr(), pos.data_ptr(), x32.data_ptr(), x16.data_ptr(), N, S, H, sp, site, p, _stream())
    _lib.check(rc, "vit_embed_fwd")
    return x32, x16


def vit_embed_bwd(dx: torch.Tensor, N: int, S: int, drop: Drop = NO_DROP):
    """-> (dfull bf16 [N*S, H] = dx o mask, dt0 bf16 [N*(S-1), H] = its patch-token rows)."""
    H = dx.shape[1]
    dfull = torch.empty_like(dx)
    dt0 = torch.empty((N * (S - 1), H), dtype=BF16, device=dx.device)
    sp, site, p = drop.args
    _lib.check(_lib.load().hamt_vit_embed_bwd(dx.data_ptr(), dfull.data_ptr(), dt0.data_ptr(), N, S, H, sp, site, p, _stream()), "vit_embed_bwd")
    return dfull, dt0
