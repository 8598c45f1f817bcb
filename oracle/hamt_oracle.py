"""CPU oracle for the HAMT hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (fp32, CPU) *functional* restatement of the arithmetic of the reference's
hot path (cshizhe/VLN-HAMT @ c8b9ee1):

  * backbone         pretrain_src/model/vilmodel.py:591-638   (NavPreTrainedModel.forward)
  * ITM backbone     pretrain_src/model/vilmodel.py:640-724   (forward_itm)
  * proxy-task heads pretrain_src/model/pretrain_cmt.py:101-262
  * finetune facade  finetune_src/models/vilmodel_cmt.py:624-728 (NavCMT.forward modes)

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module; the product package (`vln-hamt_b200/`) never does.

Pinning: this restatement is checked (tests/test_oracle_vs_reference.py, runs wherever
/root/reference exists) against the *unmodified* reference modules imported through
oracle/ref_shim.py, and against the committed golden vectors in tests/golden/ that were produced
from the reference by oracle/make_golden.py.  The reference itself ships no tests / golden vectors
(SURVEY.md section 4), so those two are the pin.

The state is a flat ``dict[str, Tensor]`` using the reference's ``state_dict`` key names, so the same
seeded weights can be loaded into the reference, the oracle and the CUDA product.

``regime``:
  * ``"fp32"``  - exact fp32 restatement of the reference (the parity target).
  * ``"bf16"``  - same algebra, but every tensor the CUDA product stores in HBM as bf16 (GEMM
                  operands, GEMM outputs, LayerNorm outputs, attention probabilities fed to the PV
                  product) is rounded to bf16 at the same point; all reductions stay fp32.  This is
                  the "same dtype regime" comparator for the 1e-3 bar (hard parts).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

LN_EPS = 1e-12          # vilmodel.py:136 (layer_norm_eps) and the literal 1e-12 in embedders/heads
MASK_NEG = -10000.0     # vilmodel.py:599


class Regime:
    """Rounding policy.  q() marks a tensor that the CUDA path materialises in bf16."""

    def __init__(self, name: str = "fp32"):
        assert name in ("fp32", "bf16")
        self.name = name
        self.bf16 = name == "bf16"

    def q(self, x: Tensor) -> Tensor:
        if self.bf16:
            # straight-through so autograd of the oracle stays usable in the bf16 regime
            return x + (x.to(torch.bfloat16).to(torch.float32) - x).detach()
        return x


FP32 = Regime("fp32")
BF16 = Regime("bf16")


# --------------------------------------------------------------------------------------
# primitive ops
# --------------------------------------------------------------------------------------

def gelu_erf(x: Tensor) -> Tensor:
    """vilmodel.py:23-29 -- exact erf GELU."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def linear(sd: State, prefix: str, x: Tensor, rg: Regime, quant_out: bool = True) -> Tensor:
    """nn.Linear.  In the bf16 regime both operands are bf16, accumulation fp32, bias fp32."""
    w = sd[prefix + ".weight"]
    b = sd.get(prefix + ".bias")
    y = F.linear(rg.q(x), rg.q(w), None)
    if b is not None:
        y = y + b
    return rg.q(y) if quant_out else y


def layer_norm(sd: State, prefix: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], LN_EPS)


def dropout(x: Tensor, p: float, mask: Optional[Tensor]) -> Tensor:
    """Dropout with an *explicit* keep-mask (1 = keep) so the CUDA Philox stream can be replayed.
    mask None  ->  identity (eval mode)."""
    if mask is None or p == 0.0:
        return x
    return x * mask.to(x.dtype) / (1.0 - p)


class DropPlan:
    """Supplies keep-masks in call order.  ``None`` plan = eval mode."""

    def __init__(self, p_hidden: float = 0.0, p_attn: float = 0.0, masks: Optional[List[Tensor]] = None):
        self.p_hidden, self.p_attn = p_hidden, p_attn
        self.masks = list(masks) if masks is not None else None
        self.i = 0

    def next(self) -> Optional[Tensor]:
        if self.masks is None:
            return None
        m = self.masks[self.i]
        self.i += 1
        return m


EVAL = None  # drop-plan placeholder for eval mode


def _hidden_drop(x: Tensor, dp: Optional[DropPlan]) -> Tensor:
    if dp is None:
        return x
    return dropout(x, dp.p_hidden, dp.next())


def _attn_drop(x: Tensor, dp: Optional[DropPlan]) -> Tensor:
    if dp is None:
        return x
    return dropout(x, dp.p_attn, dp.next())


def ext_mask(mask: Tensor) -> Tensor:
    """bool/0-1 [B,S] -> additive [B,1,1,S]; vilmodel.py:597-599."""
    return (1.0 - mask.to(torch.float32))[:, None, None, :] * MASK_NEG


def attention_core(q: Tensor, k: Tensor, v: Tensor, add_mask: Optional[Tensor], n_heads: int,
                   rg: Regime, dp: Optional[DropPlan]) -> Tensor:
    """softmax(QK^T / sqrt(d) + mask) V -- divide THEN add (vilmodel.py:106-109, 332-336);
    dropout on the probabilities (:116, :343).  q:[B,Sq,H]  k,v:[B,Sk,H]."""
    B, Sq, H = q.shape
    Sk = k.shape[1]
    d = H // n_heads
    qh = q.view(B, Sq, n_heads, d).permute(0, 2, 1, 3)
    kh = k.view(B, Sk, n_heads, d).permute(0, 2, 1, 3)
    vh = v.view(B, Sk, n_heads, d).permute(0, 2, 1, 3)
    scores = torch.matmul(qh, kh.transpose(-1, -2)) / math.sqrt(d)
    if add_mask is not None:
        scores = scores + add_mask
    if not rg.bf16:
        probs = torch.softmax(scores, dim=-1)
        probs = _attn_drop(probs, dp)
        ctx = torch.matmul(probs, vh)
    else:
        # Mirror of csrc/hamt_attn.cu: keys are consumed in blocks of 64 with an online softmax; the UNNORMALISED
        # block probabilities exp(s - running_max) are rounded to bf16 for the P.V tensor-core product, the row sum
        # uses the unrounded values, and the 1/l normalisation is applied to the fp32 accumulator at the end.
        drop_mask = dp.next() if (dp is not None and dp.masks is not None) else None
        m = torch.full(scores.shape[:-1], -float("inf"))
        l = torch.zeros(scores.shape[:-1])
        acc = torch.zeros(B, n_heads, Sq, d)
        for k0 in range(0, Sk, 64):
            s = scores[..., k0:k0 + 64]
            m_new = torch.maximum(m, s.max(-1).values)
            corr = torch.exp(m - m_new)
            p = torch.exp(s - m_new[..., None])
            l = l * corr + p.sum(-1)
            if drop_mask is not None:
                p = p * drop_mask[..., k0:k0 + 64].to(p.dtype) / (1.0 - dp.p_attn)
            acc = acc * corr[..., None] + torch.matmul(rg.q(p), vh[:, :, k0:k0 + 64])
            m = m_new
        ctx = acc / l[..., None]
    ctx = ctx.permute(0, 2, 1, 3).contiguous().view(B, Sq, H)
    return rg.q(ctx)


def self_output(sd: State, prefix: str, hidden: Tensor, residual: Tensor, rg: Regime,
                dp: Optional[DropPlan]) -> Tensor:
    """BertSelfOutput / BertOutput: LN(dropout(dense(h)) + residual); vilmodel.py:139-143,181-185."""
    t = linear(sd, prefix + ".dense", hidden, rg)
    t = _hidden_drop(t, dp)
    # bf16 regime: the CUDA path keeps the LayerNorm output in fp32 for the next residual add (the residual stream) and hands a
    # bf16 copy to the GEMMs -- `linear` rounds its input itself, so the value returned here stays un-rounded
    return layer_norm(sd, prefix + ".LayerNorm", t + residual)


def bert_self_attention(sd: State, prefix: str, x: Tensor, add_mask: Tensor, n_heads: int, rg: Regime,
                        dp: Optional[DropPlan]) -> Tensor:
    """BertAttention = BertSelfAttention + BertSelfOutput; vilmodel.py:96-129,146-157."""
    q = linear(sd, prefix + ".self.query", x, rg)
    k = linear(sd, prefix + ".self.key", x, rg)
    v = linear(sd, prefix + ".self.value", x, rg)
    ctx = attention_core(q, k, v, add_mask, n_heads, rg, dp)
    return self_output(sd, prefix + ".output", ctx, x, rg, dp)


def ffn(sd: State, inter_prefix: str, out_prefix: str, x: Tensor, rg: Regime, dp: Optional[DropPlan]) -> Tensor:
    """BertIntermediate + BertOutput; vilmodel.py:168-171,181-185.  GELU is applied to the fp32
    accumulator (+bias) before the bf16 store in the CUDA epilogue."""
    h = linear(sd, inter_prefix + ".dense", x, rg, quant_out=False)
    a = rg.q(gelu_erf(h))
    return self_output(sd, out_prefix, a, x, rg, dp)


def bert_layer(sd: State, prefix: str, x: Tensor, add_mask: Tensor, n_heads: int, rg: Regime,
               dp: Optional[DropPlan]) -> Tensor:
    """BertLayer; vilmodel.py:195-201."""
    a = bert_self_attention(sd, prefix + ".attention", x, add_mask, n_heads, rg, dp)
    return ffn(sd, prefix + ".intermediate", prefix + ".output", a, rg, dp)


def cross_attention(sd: State, prefix: str, x: Tensor, ctx_in: Tensor, ctx_mask: Optional[Tensor], n_heads: int,
                    rg: Regime, dp: Optional[DropPlan]) -> Tensor:
    """BertXAttention (BertOutAttention + BertSelfOutput); vilmodel.py:322-360."""
    q = linear(sd, prefix + ".att.query", x, rg)
    k = linear(sd, prefix + ".att.key", ctx_in, rg)
    v = linear(sd, prefix + ".att.value", ctx_in, rg)
    ctx = attention_core(q, k, v, ctx_mask, n_heads, rg, dp)
    return self_output(sd, prefix + ".output", ctx, x, rg, dp)


def lxrt_x_layer(sd: State, prefix: str, lang: Tensor, lang_mask: Tensor, visn: Tensor, visn_mask: Tensor,
                 n_heads: int, rg: Regime, dp: Optional[DropPlan], no_lang_ca: bool = False) -> Tuple[Tensor, Tensor]:
    """LXRTXLayer.forward; vilmodel.py:401-412 (finetune variant with no_lang_ca:
    vilmodel_cmt.py:379-424).  One visual_attention module serves BOTH directions and both
    directions read the layer inputs (vilmodel.py:379-383)."""
    if no_lang_ca:
        lang_x = lang
    else:
        lang_x = cross_attention(sd, prefix + ".visual_attention", lang, visn, visn_mask, n_heads, rg, dp)
    visn_x = cross_attention(sd, prefix + ".visual_attention", visn, lang, lang_mask, n_heads, rg, dp)
    if no_lang_ca:
        lang_s = lang_x
    else:
        lang_s = bert_self_attention(sd, prefix + ".lang_self_att", lang_x, lang_mask, n_heads, rg, dp)
    visn_s = bert_self_attention(sd, prefix + ".visn_self_att", visn_x, visn_mask, n_heads, rg, dp)
    # NB reference call order in output_fc: lang_inter, visn_inter, lang_output, visn_output
    # (vilmodel.py:391-399).  Only matters for the order dropout masks are consumed in; the drop
    # plan used by the tests is keyed per site, so we keep the natural per-stream order here.
    if no_lang_ca:
        lang_o = lang_s
    else:
        lang_o = ffn(sd, prefix + ".lang_inter", prefix + ".lang_output", lang_s, rg, dp)
    visn_o = ffn(sd, prefix + ".visn_inter", prefix + ".visn_output", visn_s, rg, dp)
    return lang_o, visn_o


# --------------------------------------------------------------------------------------
# embedders
# --------------------------------------------------------------------------------------

def text_embeddings(sd: State, prefix: str, txt_ids: Tensor, rg: Regime, dp: Optional[DropPlan]) -> Tensor:
    """BertEmbeddings.forward, token_type 0, positions arange(L); vilmodel.py:54-69."""
    L = txt_ids.shape[1]
    e = sd[prefix + ".word_embeddings.weight"][txt_ids] \
        + sd[prefix + ".position_embeddings.weight"][:L][None] \
        + sd[prefix + ".token_type_embeddings.weight"][0][None, None]
    return rg.q(_hidden_drop(layer_norm(sd, prefix + ".LayerNorm", e), dp))


def _feat_embed(sd: State, prefix: str, img: Tensor, ang: Tensor, rg: Regime) -> Tensor:
    """LN(img_linear(x)) + LN(ang_linear(a)); vilmodel.py:497-498.  The angle projection (K=4) is
    done in fp32 on CUDA cores in the product, so it is not quantised in the bf16 regime."""
    ti = layer_norm(sd, prefix + "img_layer_norm", linear(sd, prefix + "img_linear", img, rg))   # GEMM output is stored bf16
    ta = layer_norm(sd, prefix + "ang_layer_norm", linear(sd, prefix + "ang_linear", ang, FP32))
    return ti + ta


def image_embeddings(sd: State, prefix: str, img: Tensor, ang: Tensor, type_emb: Tensor, nav_types: Optional[Tensor],
                     rg: Regime, dp: Optional[DropPlan]) -> Tensor:
    """ImageEmbeddings.forward; vilmodel.py:496-505."""
    e = _feat_embed(sd, prefix + ".", img, ang, rg) + type_emb
    if nav_types is not None:
        e = e + sd[prefix + ".nav_type_embedding.weight"][nav_types]
    return rg.q(_hidden_drop(layer_norm(sd, prefix + ".layer_norm", e), dp))


def pano_encode(sd: State, prefix: str, pano_img: Tensor, pano_ang: Tensor, n_layers: int, n_heads: int,
                rg: Regime, dp: Optional[DropPlan], drop_pano_emb: bool) -> Tensor:
    """The hierarchical part of HistoryEmbeddings: [N,P,F] views -> one token per panorama.
    vilmodel.py:553-564 (pretrain: no dropout on the pano token embeddings) /
    vilmodel_cmt.py:581-590 (finetune: hidden dropout on them, :583).  Zero mask, plain mean."""
    N, P, _ = pano_img.shape
    e = _feat_embed(sd, prefix + ".pano_", pano_img, pano_ang, rg)
    if drop_pano_emb:
        e = _hidden_drop(e, dp)
    e = rg.q(e)
    zero_mask = torch.zeros(N, 1, 1, P, device=e.device)
    for l in range(n_layers):
        e = bert_layer(sd, f"{prefix}.pano_encoder.layer.{l}", e, zero_mask, n_heads, rg, dp)
    return rg.q(e).mean(dim=1)          # the mean-pool kernel reads the bf16 copy


def history_embeddings(sd: State, cfg, prefix: str, img: Optional[Tensor], ang: Optional[Tensor],
                       pano_img: Optional[Tensor], pano_ang: Optional[Tensor], pos_ids: Optional[Tensor],
                       batch_size: int, rg: Regime, dp: Optional[DropPlan]) -> Tuple[Tensor, Optional[Tensor]]:
    """HistoryEmbeddings.forward (pretrain); vilmodel.py:540-575.
    Returns (cls [B,1,H], emb [B,T,H] or None).  With pos_ids None the step embeddings are returned
    *before* position/LN/dropout (used by forward_itm, vilmodel.py:664-666)."""
    type_emb = sd[prefix + ".type_embedding.weight"][0][None, None]                  # [1,1,H]
    cls = sd[prefix + ".cls_token"].expand(batch_size, -1, -1) + type_emb
    cls = rg.q(_hidden_drop(layer_norm(sd, prefix + ".layer_norm", cls), dp))
    if img is None:
        return cls, None
    e = _feat_embed(sd, prefix + ".", img, ang, rg) + type_emb
    if cfg.num_h_pano_layers > 0:
        B, T, P, Fd = pano_img.shape
        pe = pano_encode(sd, prefix, pano_img.reshape(B * T, P, Fd), pano_ang.reshape(B * T, P, -1),
                         cfg.num_h_pano_layers, cfg.num_attention_heads, rg, dp, drop_pano_emb=False)
        e = e + pe.view(B, T, -1)
    if pos_ids is not None:
        e = e + sd[prefix + ".position_embeddings.weight"][pos_ids]
        e = rg.q(_hidden_drop(layer_norm(sd, prefix + ".layer_norm", e), dp))
    else:
        e = rg.q(e)            # ITM: the pre-position sum is materialised (bf16 on the CUDA path)
    return cls, e


# --------------------------------------------------------------------------------------
# backbone
# --------------------------------------------------------------------------------------

def encoder(sd: State, cfg, prefix: str, txt: Tensor, txt_mask: Tensor, hist: Tensor, hist_mask: Tensor,
            ob: Optional[Tensor], ob_mask: Optional[Tensor], rg: Regime, dp: Optional[DropPlan]):
    """LxmertEncoder.forward; vilmodel.py:438-478 (num_r_layers / num_h_layers honoured)."""
    nh = cfg.num_attention_heads
    for l in range(cfg.num_l_layers):
        txt = bert_layer(sd, f"{prefix}.layer.{l}", txt, txt_mask, nh, rg, dp)
    if not getattr(cfg, "update_lang_bert", True):
        txt = txt.detach()
    if ob is not None:
        for l in range(cfg.num_r_layers):
            ob = bert_layer(sd, f"{prefix}.r_layers.{l}", ob, ob_mask, nh, rg, dp)
    for l in range(cfg.num_h_layers):
        hist = bert_layer(sd, f"{prefix}.h_layers.{l}", hist, hist_mask, nh, rg, dp)
    T1 = hist.shape[1]
    if ob is None:
        visn, visn_mask = hist, hist_mask
    else:
        visn, visn_mask = torch.cat([hist, ob], 1), torch.cat([hist_mask, ob_mask], -1)
    for l in range(cfg.num_x_layers):
        txt, visn = lxrt_x_layer(sd, f"{prefix}.x_layers.{l}", txt, txt_mask, visn, visn_mask, nh, rg, dp)
    txt, visn = rg.q(txt), rg.q(visn)   # what the backbone hands to the heads is the bf16 copy
    hist = visn[:, :T1]
    ob_out = visn[:, T1:] if ob is not None else None
    return txt, hist, ob_out


def backbone(sd: State, cfg, txt_ids, txt_masks, hist_img, hist_ang, hist_pano_img, hist_pano_ang, hist_masks,
             ob_img, ob_ang, ob_nav_types, ob_masks, rg: Regime = FP32, dp: Optional[DropPlan] = None,
             prefix: str = "bert"):
    """NavPreTrainedModel.forward; vilmodel.py:591-638."""
    B = txt_ids.shape[0]
    txt_mask = ext_mask(txt_masks)
    txt = text_embeddings(sd, prefix + ".embeddings", txt_ids, rg, dp)
    hist_mask = ext_mask(hist_masks)
    pos_ids = torch.arange(hist_img.shape[1], device=hist_img.device)[None] if hist_img is not None else None
    cls, vp = history_embeddings(sd, cfg, prefix + ".hist_embeddings", hist_img, hist_ang, hist_pano_img,
                                 hist_pano_ang, pos_ids, B, rg, dp)
    hist = cls if vp is None else torch.cat([cls, vp], 1)
    if ob_img is not None:
        type_emb = sd[prefix + ".embeddings.token_type_embeddings.weight"][1][None, None]   # vilmodel.py:622-625
        ob = image_embeddings(sd, prefix + ".img_embeddings", ob_img, ob_ang, type_emb, ob_nav_types, rg, dp)
        ob_mask = ext_mask(ob_masks)
    else:
        ob, ob_mask = None, None
    return encoder(sd, cfg, prefix + ".encoder", txt, txt_mask, hist, hist_mask, ob, ob_mask, rg, dp)


def itm_negative_plan(batch_size: int, hist_masks: Tensor, hist_max_len: int, num_neg_trajs: int = 4):
    """Host-side RNG draws of forward_itm in the reference's exact call order
    (vilmodel.py:676-704): np.random.choice per sample, then torch.randperm per sample per K."""
    K = num_neg_trajs // 2
    neg_idxs = None
    if batch_size > 1:
        rows = []
        for i in range(batch_size):
            rows.append(np.random.choice(np.arange(0, i).tolist() + np.arange(i + 1, batch_size).tolist(), K))
        neg_idxs = torch.from_numpy(np.stack(rows, 0))
    else:
        K = num_neg_trajs
    hist_lens = torch.sum(hist_masks, 1) - 1
    shuffled = []
    for _ in range(K):
        per = []
        for i in range(batch_size):
            idx = torch.randperm(int(hist_lens[i]))
            idx = torch.cat([idx, torch.arange(int(hist_lens[i]), hist_max_len, dtype=torch.long)], 0)
            per.append(idx)
        shuffled.append(torch.stack(per, 0))
    return neg_idxs, shuffled


def backbone_itm(sd: State, cfg, txt_ids, txt_masks, hist_img, hist_ang, hist_pano_img, hist_pano_ang, hist_masks,
                 num_neg_trajs: int = 4, rg: Regime = FP32, dp: Optional[DropPlan] = None, prefix: str = "bert",
                 plan=None):
    """NavPreTrainedModel.forward_itm; vilmodel.py:640-724.  ``plan`` = itm_negative_plan(...) output;
    when None it is drawn here from the global numpy/torch RNGs exactly as the reference does."""
    nh = cfg.num_attention_heads
    B, T, _ = hist_img.shape
    txt_mask = ext_mask(txt_masks)
    txt = text_embeddings(sd, prefix + ".embeddings", txt_ids, rg, dp)
    for l in range(cfg.num_l_layers):
        txt = bert_layer(sd, f"{prefix}.encoder.layer.{l}", txt, txt_mask, nh, rg, dp)
    txt = txt.repeat(1 + num_neg_trajs, 1, 1)
    txt_mask_r = txt_mask.repeat(1 + num_neg_trajs, 1, 1, 1)

    hist_mask = ext_mask(hist_masks)
    hp = prefix + ".hist_embeddings"
    cls, vp_nopos = history_embeddings(sd, cfg, hp, hist_img, hist_ang, hist_pano_img, hist_pano_ang, None, B, rg, dp)
    pos_w = sd[hp + ".position_embeddings.weight"]

    def with_pos(pos_ids):
        return rg.q(_hidden_drop(layer_norm(sd, hp + ".layer_norm", vp_nopos + pos_w[pos_ids.to(pos_w.device)]), dp))

    def h_layers(x):
        for l in range(cfg.num_h_layers):
            x = bert_layer(sd, f"{prefix}.encoder.h_layers.{l}", x, hist_mask, nh, rg, dp)
        return x

    hist = h_layers(torch.cat([cls, with_pos(torch.arange(T)[None])], 1))
    if plan is None:
        plan = itm_negative_plan(B, hist_masks.cpu(), T, num_neg_trajs)
    neg_idxs, shuffled = plan
    if neg_idxs is not None:
        neg_idxs = neg_idxs.to(hist.device)
    neg_embeds, neg_masks = [], []
    if neg_idxs is not None:
        for k in range(neg_idxs.shape[1]):
            neg_embeds.append(hist[neg_idxs[:, k]])
            neg_masks.append(hist_mask[neg_idxs[:, k]])
    for pos_ids in shuffled:
        neg_embeds.append(h_layers(torch.cat([cls, with_pos(pos_ids)], 1)))
        neg_masks.append(hist_mask)
    visn = torch.cat([hist] + neg_embeds, 0)
    visn_mask = torch.cat([hist_mask] + neg_masks, 0)
    for l in range(cfg.num_x_layers):
        txt, visn = lxrt_x_layer(sd, f"{prefix}.encoder.x_layers.{l}", txt, txt_mask_r, visn, visn_mask, nh, rg, dp)
    fused = rg.q(txt[:, 0]) * rg.q(visn[:, 0])
    return torch.stack(torch.split(fused, B), 1)


# --------------------------------------------------------------------------------------
# heads + losses (pretrain_cmt.py)
# --------------------------------------------------------------------------------------

def mlp_head(sd: State, prefix: str, x: Tensor, rg: Regime, dp: Optional[DropPlan], has_dropout: bool) -> Tensor:
    """Linear -> ReLU -> LN(1e-12) -> [Dropout] -> Linear; pretrain_cmt.py:13-71.  The head output
    (logits) is kept fp32."""
    h = rg.q(torch.relu(linear(sd, prefix + ".net.0", x, rg, quant_out=False)))
    h = rg.q(layer_norm(sd, prefix + ".net.2", h))
    last = ".net.4" if has_dropout else ".net.3"
    if has_dropout:
        h = rg.q(_hidden_drop(h, dp))
    w = sd[prefix + last + ".weight"]
    if w.shape[0] <= 4:      # narrow output: the CUDA path does warp dot products against the fp32 weights
        return F.linear(h, w, sd[prefix + last + ".bias"])
    return linear(sd, prefix + last, h, rg, quant_out=False)


def mlm_head(sd: State, prefix: str, x: Tensor, rg: Regime) -> Tensor:
    """BertOnlyMLMHead; vilmodel.py:252-295; decoder tied to word embeddings (pretrain_cmt.py:96-99)."""
    p = prefix + ".predictions"
    h = rg.q(gelu_erf(linear(sd, p + ".transform.dense", x, rg, quant_out=False)))
    h = rg.q(layer_norm(sd, p + ".transform.LayerNorm", h))
    w = sd[p + ".decoder.weight"]
    return F.linear(h, rg.q(w)) + sd[p + ".bias"]


def masked_rows(hidden: Tensor, mask: Tensor) -> Tensor:
    """_compute_masked_hidden; pretrain_cmt.py:161-165."""
    return hidden[mask]


def pretrain_forward(sd: State, cfg, batch: dict, task: str, compute_loss: bool = True, rg: Regime = FP32,
                     dp: Optional[DropPlan] = None, itm_plan=None):
    """MultiStepNavCMTPreTraining.forward; pretrain_cmt.py:101-262.  Returns exactly what the
    reference returns: un-reduced loss vectors, or logits when compute_loss is False."""
    g = lambda k: batch.get(k)
    hist_args = (g("hist_img_fts"), g("hist_ang_fts"), g("hist_pano_img_fts"), g("hist_pano_ang_fts"), g("hist_masks"))
    ob_args = (g("ob_img_fts"), g("ob_ang_fts"), g("ob_nav_types"), g("ob_masks"))
    none4 = (None, None, None, None)
    if task.startswith("mlm"):
        txt, _, _ = backbone(sd, cfg, g("txt_ids"), g("txt_masks"), *hist_args, *none4, rg=rg, dp=dp)
        labels = g("txt_labels")
        scores = mlm_head(sd, "mlm_head", masked_rows(txt, labels != -1), rg)
        if compute_loss:
            return F.cross_entropy(scores, labels[labels != -1], reduction="none")
        return scores
    if task.startswith("sap"):
        txt, hist, ob = backbone(sd, cfg, g("txt_ids"), g("txt_masks"), *hist_args, *ob_args, rg=rg, dp=dp)
        scores = mlp_head(sd, "next_action", ob * txt[:, :1], rg, dp, True).squeeze(-1)
        scores = scores.masked_fill(g("ob_nav_types") == 0, -float("inf"))
        if compute_loss:
            return F.cross_entropy(scores, g("ob_action_viewindex"), reduction="none")
        return scores
    if task.startswith("sar"):
        txt, hist, ob = backbone(sd, cfg, g("txt_ids"), g("txt_masks"), *hist_args, *ob_args, rg=rg, dp=dp)
        scores = mlp_head(sd, "regress_action", txt[:, 0], rg, dp, True)
        if compute_loss:
            tgt = torch.cat([g("ob_action_angles"), g("ob_progress").unsqueeze(1)], dim=1)
            return F.mse_loss(scores, tgt, reduction="none")
        return scores
    if task.startswith("sprel"):
        txt, hist, ob = backbone(sd, cfg, g("txt_ids"), g("txt_masks"), *hist_args, *ob_args, rg=rg, dp=dp)
        idx = g("sp_anchor_idxs")
        anchor = torch.gather(ob, 1, idx[:, None, None].repeat(1, 36, ob.shape[-1]))
        scores = mlp_head(sd, "sprel_head", torch.cat([anchor, ob[:, :-1]], -1), rg, dp, True)
        if compute_loss:
            return F.mse_loss(scores, g("sp_targets"), reduction="none")
        return scores
    if task.startswith("mrc"):
        txt, hist, _ = backbone(sd, cfg, g("txt_ids"), g("txt_masks"), *hist_args, *none4, rg=rg, dp=dp)
        m = g("hist_mrc_masks")
        pred = mlp_head(sd, "image_classifier", masked_rows(hist[:, 1:], m), rg, dp, False)
        tgt = masked_rows(g("hist_img_probs"), m)
        if compute_loss:
            return F.kl_div(F.log_softmax(pred, dim=-1), tgt, reduction="none").sum(dim=1)
        return pred, tgt
    if task.startswith("itm"):
        fused = backbone_itm(sd, cfg, g("txt_ids"), g("txt_masks"), *hist_args, 4, rg=rg, dp=dp, plan=itm_plan)
        scores = mlp_head(sd, "itm_head", fused, rg, dp, False).squeeze(2)
        tgt = torch.zeros(fused.shape[0], dtype=torch.long, device=fused.device)
        if compute_loss:
            return F.cross_entropy(scores, tgt, reduction="none")
        return scores, tgt
    raise ValueError("invalid task")      # pretrain_cmt.py:140


# --------------------------------------------------------------------------------------
# finetune facade (vilmodel_cmt.py NavCMT.forward)
# --------------------------------------------------------------------------------------

def navcmt_language(sd: State, cfg, txt_ids, txt_masks, rg: Regime = FP32, dp: Optional[DropPlan] = None):
    """mode == 'language'; vilmodel_cmt.py:632-653."""
    nh = cfg.num_attention_heads
    m = ext_mask(txt_masks)
    txt = text_embeddings(sd, "embeddings", txt_ids, rg, dp)
    for l in range(cfg.num_l_layers):
        txt = bert_layer(sd, f"encoder.layer.{l}", txt, m, nh, rg, dp)
    if cfg.fix_lang_embedding:
        txt = txt.detach()
    if cfg.no_lang_ca:
        outs = [rg.q(txt)]
        for l in range(cfg.num_x_layers):
            p = f"encoder.x_layers.{l}"
            a = bert_self_attention(sd, p + ".lang_self_att", txt, m, nh, rg, dp)
            outs.append(rg.q(ffn(sd, p + ".lang_inter", p + ".lang_output", a, rg, dp)))
        return outs
    return rg.q(txt)                    # the mode returns the bf16 copy


def navcmt_history(sd: State, cfg, hist_img, hist_ang, ob_step_ids, pano_img=None, pano_ang=None,
                   rg: Regime = FP32, dp: Optional[DropPlan] = None):
    """mode == 'history' (one step of the hierarchical encoder); vilmodel_cmt.py:553-594,656-661."""
    hp = "hist_embeddings"
    type_emb = sd[hp + ".type_embedding.weight"][0][None]                           # [1,H]
    if hist_img is None:
        cls = sd[hp + ".cls_token"].expand(1, -1, -1)[:, 0] + type_emb
        out = rg.q(_hidden_drop(layer_norm(sd, hp + ".layer_norm", cls), dp))
    else:
        e = _feat_embed(sd, hp + ".", hist_img, hist_ang, rg) \
            + sd[hp + ".position_embeddings.weight"][ob_step_ids] + type_emb
        if cfg.hist_enc_pano:
            e = e + pano_encode(sd, hp, pano_img, pano_ang, cfg.num_h_pano_layers, cfg.num_attention_heads,
                                rg, dp, drop_pano_emb=True)
        out = rg.q(_hidden_drop(layer_norm(sd, hp + ".layer_norm", e), dp))
    if cfg.fix_hist_embedding:
        out = out.detach()
    return out


def navcmt_visual(sd: State, cfg, txt_embeds, txt_masks, hist_embeds, hist_masks, ob_img, ob_ang, ob_nav_types,
                  ob_masks, rg: Regime = FP32, dp: Optional[DropPlan] = None):
    """mode == 'visual'; vilmodel_cmt.py:664-728."""
    nh = cfg.num_attention_heads
    hist_mask = ext_mask(hist_masks)
    hist = hist_embeds
    for l in range(cfg.num_h_layers):
        hist = bert_layer(sd, f"encoder.h_layers.{l}", hist, hist_mask, nh, rg, dp)
    ob_mask = ext_mask(ob_masks)
    type_emb = sd["embeddings.token_type_embeddings.weight"][1][None, None]
    ob = image_embeddings(sd, "img_embeddings", ob_img, ob_ang, type_emb, ob_nav_types, rg, dp)
    for l in range(cfg.num_r_layers):
        ob = bert_layer(sd, f"encoder.r_layers.{l}", ob, ob_mask, nh, rg, dp)
    if cfg.fix_obs_embedding:
        ob = ob.detach()
    T1 = hist.shape[1]
    visn = torch.cat([hist, ob], 1)
    visn_mask = torch.cat([hist_mask, ob_mask], -1)
    txt_mask = ext_mask(txt_masks)
    all_txt = txt_embeds if cfg.no_lang_ca else None
    txt = None if cfg.no_lang_ca else txt_embeds
    for l in range(cfg.num_x_layers):
        if cfg.no_lang_ca:
            txt = all_txt[l]
        txt, visn = lxrt_x_layer(sd, f"encoder.x_layers.{l}", txt, txt_mask, visn, visn_mask, nh, rg, dp,
                                 no_lang_ca=cfg.no_lang_ca)
    txt, visn = rg.q(txt), rg.q(visn)
    hist, ob = visn[:, :T1], visn[:, T1:]
    if cfg.no_lang_ca or cfg.act_pred_token == "ob":
        fused = ob
    elif cfg.act_pred_token == "ob_txt":
        fused = ob * txt[:, :1]
    elif cfg.act_pred_token == "ob_hist":
        fused = ob * hist[:, :1]
    elif cfg.act_pred_token == "ob_txt_hist":
        fused = ob * (txt[:, :1] + hist[:, :1])
    else:
        raise ValueError(cfg.act_pred_token)
    logits = mlp_head(sd, "next_action", fused, rg, dp, True).squeeze(-1)
    logits = logits.masked_fill(ob_nav_types == 0, -float("inf"))
    return logits, txt, hist, ob
