"""B200-native mirror of the reference backbone ``pretrain_src/model/vilmodel.py``.

Same class names, constructor ``(config)``, attribute names (=> identical ``state_dict`` keys) and the
same ``NavPreTrainedModel.forward`` / ``forward_itm`` signatures (vilmodel.py:591-593, :640-642), so a
reference checkpoint loads unchanged and ``pretrain_src/main_r2r.py:237`` can call it as is.  The
modules are parameter containers; the arithmetic runs in the hand-written sm_100a kernels through the
fused blocks of functional.py (no eager / CPU fallback: calling forward without the native library or
on a CPU tensor raises).
"""
from __future__ import annotations

import copy
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import functional as Fn
from . import ops
from .arena import ParamArena

BertLayerNorm = nn.LayerNorm
BF16 = torch.bfloat16


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(f"{type(self).__name__} is fused into its parent block on the sm_100a path; call the parent model")


class BertEmbeddings(_Container):
    """vilmodel.py:40-69"""

    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class BertSelfAttention(_Container):
    """vilmodel.py:72-129"""

    def __init__(self, config):
        super().__init__()
        if config.hidden_size % config.num_attention_heads != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)"
                             % (config.hidden_size, config.num_attention_heads))
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = int(config.hidden_size / config.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(config.hidden_size, self.all_head_size)
        self.value = nn.Linear(config.hidden_size, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)


class BertSelfOutput(_Container):
    """vilmodel.py:132-143"""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class BertAttention(_Container):
    """vilmodel.py:146-157"""

    def __init__(self, config):
        super().__init__()
        self.self = BertSelfAttention(config)
        self.output = BertSelfOutput(config)


class BertIntermediate(_Container):
    """vilmodel.py:159-171 (hidden_act must be the exact-erf 'gelu')."""

    def __init__(self, config):
        super().__init__()
        if config.hidden_act != "gelu":
            raise ValueError("hamt_b200: only hidden_act='gelu' (exact erf, vilmodel.py:23-29) is implemented")
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)


class BertOutput(_Container):
    """vilmodel.py:174-185"""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class BertLayer(_Container):
    """vilmodel.py:188-201"""

    def __init__(self, config):
        super().__init__()
        self.attention = BertAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)


class BertEncoder(_Container):
    """vilmodel.py:204-236"""

    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(config) for _ in range(config.num_hidden_layers)])


class BertPredictionHeadTransform(_Container):
    """vilmodel.py:252-266"""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class BertLMPredictionHead(_Container):
    """vilmodel.py:269-285"""

    def __init__(self, config):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))


class BertOnlyMLMHead(_Container):
    """vilmodel.py:288-295"""

    def __init__(self, config):
        super().__init__()
        self.predictions = BertLMPredictionHead(config)


class BertOutAttention(_Container):
    """vilmodel.py:298-349"""

    def __init__(self, config, ctx_dim=None):
        super().__init__()
        if config.hidden_size % config.num_attention_heads != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)"
                             % (config.hidden_size, config.num_attention_heads))
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = int(config.hidden_size / config.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        if ctx_dim is None:
            ctx_dim = config.hidden_size
        if ctx_dim != config.hidden_size:
            raise ValueError("hamt_b200: cross-attention context width must equal hidden_size (shared fused QKV)")
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(ctx_dim, self.all_head_size)
        self.value = nn.Linear(ctx_dim, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)


class BertXAttention(_Container):
    """vilmodel.py:351-360"""

    def __init__(self, config, ctx_dim=None):
        super().__init__()
        self.att = BertOutAttention(config, ctx_dim=ctx_dim)
        self.output = BertSelfOutput(config)


class LXRTXLayer(_Container):
    """vilmodel.py:362-412"""

    def __init__(self, config):
        super().__init__()
        self.lang_self_att = BertAttention(config)
        self.lang_inter = BertIntermediate(config)
        self.lang_output = BertOutput(config)
        self.visn_self_att = BertAttention(config)
        self.visn_inter = BertIntermediate(config)
        self.visn_output = BertOutput(config)
        self.visual_attention = BertXAttention(config)


class LxmertEncoder(_Container):
    """vilmodel.py:414-478"""

    def __init__(self, config):
        super().__init__()
        self.num_l_layers = config.num_l_layers
        self.num_r_layers = config.num_r_layers
        self.num_h_layers = config.num_h_layers
        self.num_x_layers = config.num_x_layers
        self.update_lang_bert = config.update_lang_bert
        self.layer = nn.ModuleList([BertLayer(config) for _ in range(self.num_l_layers)])
        self.h_layers = nn.ModuleList([BertLayer(config) for _ in range(self.num_h_layers)]) if self.num_h_layers > 0 else None
        self.r_layers = nn.ModuleList([BertLayer(config) for _ in range(self.num_r_layers)]) if self.num_r_layers > 0 else None
        self.x_layers = nn.ModuleList([LXRTXLayer(config) for _ in range(self.num_x_layers)])


class ImageEmbeddings(_Container):
    """vilmodel.py:482-505"""

    def __init__(self, config):
        super().__init__()
        self.img_linear = nn.Linear(config.image_feat_size, config.hidden_size)
        self.img_layer_norm = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.ang_linear = nn.Linear(config.angle_feat_size, config.hidden_size)
        self.ang_layer_norm = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.nav_type_embedding = nn.Embedding(3, config.hidden_size)
        self.layer_norm = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class HistoryEmbeddings(_Container):
    """vilmodel.py:507-575"""

    def __init__(self, config):
        super().__init__()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, config.hidden_size))
        self.img_linear = nn.Linear(config.image_feat_size, config.hidden_size)
        self.img_layer_norm = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.ang_linear = nn.Linear(config.angle_feat_size, config.hidden_size)
        self.ang_layer_norm = BertLayerNorm(config.hidden_size, eps=1e-12)
        if config.num_h_pano_layers > 0:
            self.pano_img_linear = nn.Linear(config.image_feat_size, config.hidden_size)
            self.pano_img_layer_norm = BertLayerNorm(config.hidden_size, eps=1e-12)
            self.pano_ang_linear = nn.Linear(config.angle_feat_size, config.hidden_size)
            self.pano_ang_layer_norm = BertLayerNorm(config.hidden_size, eps=1e-12)
            pano_encoder_config = copy.copy(config)
            pano_encoder_config.num_hidden_layers = config.num_h_pano_layers
            self.pano_encoder = BertEncoder(pano_encoder_config)
        else:
            self.pano_encoder = None
        self.position_embeddings = nn.Embedding(config.max_action_steps, config.hidden_size)
        self.type_embedding = nn.Embedding(1, config.hidden_size)
        self.layer_norm = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


# ------------------------------------------------------------------------------------------------
# base class (stands in for transformers.BertPreTrainedModel: init / tying / (de)serialisation only)
# ------------------------------------------------------------------------------------------------
class HamtPreTrainedModel(nn.Module):
    base_model_prefix = "bert"

    def __init__(self, config, *inputs, **kwargs):
        super().__init__()
        self.config = config
        self._arena: Optional[ParamArena] = None
        heads, hidden = getattr(config, "num_attention_heads", None), getattr(config, "hidden_size", None)
        if heads and hidden and (hidden % heads != 0 or hidden // heads != 64):
            raise ValueError(f"hamt_b200: the attention kernels are built for head_dim 64 (hidden_size {hidden} / num_attention_heads {heads} "
                             f"= {hidden / heads:g})")

    def _init_weights(self, module):
        """BERT init (transformers 4.12.3 BertPreTrainedModel._init_weights): N(0, initializer_range), zero bias, LN = (1, 0)."""
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def init_weights(self):
        self.apply(self._init_weights)

    def _tie_or_clone_weights(self, output_embeddings, input_embeddings):
        output_embeddings.weight = input_embeddings.weight

    def tie_weights(self):
        pass

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, *args, config=None, state_dict=None, **kwargs):
        """Reference call shape: ``Model.from_pretrained(None, config=cfg, state_dict=sd)`` (main_r2r.py:146-148,
        vlnbert_init.py:65-68).  Missing / unexpected keys are tolerated like HF's non-strict load."""
        if pretrained_model_name_or_path is not None and state_dict is None:
            state_dict = torch.load(pretrained_model_name_or_path, map_location="cpu")
        model = cls(config)
        if state_dict:
            model.load_state_dict(state_dict, strict=False)
            model.tie_weights()
        return model

    # ---- arena / run plumbing -------------------------------------------------------------
    def arena(self) -> ParamArena:
        root = self._arena_root()
        if root._arena is None:
            root._arena = ParamArena(root)
        return root._arena

    def _arena_root(self):
        return getattr(self, "_arena_owner", None) or self

    def begin(self) -> Fn.Run:
        """Start a forward: refresh the bf16 shadow, attach/zero gradients, advance the dropout seed."""
        pend = getattr(self, "_pending_run", None)
        if pend is not None:            # a caller (image_pretrain.py) already opened this step's Run for the vision backbone
            self._pending_run = None
            return pend
        arena = self.arena()
        arena.step_begin(self.training and torch.is_grad_enabled())
        seed = None
        if self.training:
            arena.next_seed()
            seed = arena.run_seed()
        return Fn.Run(arena, self.training, self.config.num_attention_heads, float(self.config.layer_norm_eps), seed=seed)


def itm_negative_plan(batch_size: int, hist_masks: torch.Tensor, hist_max_len: int, num_neg_trajs: int = 4):
    """Negative-trajectory indices of forward_itm, drawn from the global numpy / torch RNGs in the reference's exact call
    order (vilmodel.py:676-704): K in-batch negatives per sample via np.random.choice, then K position shuffles via
    torch.randperm(hist_len) per sample.  Returns (neg_idxs [B,K] or None, [K tensors [B,T]]) on the CPU."""
    K = num_neg_trajs // 2
    neg_idxs = None
    if batch_size > 1:
        rows = []
        for i in range(batch_size):
            rows.append(np.random.choice(np.arange(0, i).tolist() + np.arange(i + 1, batch_size).tolist(), K))
        neg_idxs = torch.from_numpy(np.stack(rows, 0))
    else:
        K = num_neg_trajs
    hist_lens = (torch.sum(hist_masks, 1) - 1).tolist()
    shuffled = []
    for _ in range(K):
        rows = []
        for i in range(batch_size):
            idx = torch.randperm(int(hist_lens[i]))
            rows.append(torch.cat([idx, torch.arange(int(hist_lens[i]), hist_max_len, dtype=torch.long)], 0))
        shuffled.append(torch.stack(rows, 0))
    return neg_idxs, shuffled


def itm_negative_plan_device(batch_size: int, hist_masks: torch.Tensor, hist_max_len: int, num_neg_trajs: int = 4):
    """Device-side negative sampling for forward_itm: the same DISTRIBUTION as itm_negative_plan -- K in-batch negatives
    per sample, uniform over the other samples with replacement (np.random.choice's default), and K position shuffles, each a uniform
    permutation of the sample's valid history steps followed by the padding positions in order -- drawn with torch's CUDA generator,
    no host loop, no host sync, capturable in the step graph (fresh draws on every replay).  Not the reference's RNG STREAM: opt-in
    through `config.itm_device_negatives`; the default keeps the host plan so that the draws match the reference call for call."""
    dev = hist_masks.device
    K = num_neg_trajs // 2
    neg_idxs = None
    if batch_size > 1:
        r = torch.randint(0, batch_size - 1, (batch_size, K), device=dev)
        neg_idxs = r + (r >= torch.arange(batch_size, device=dev).unsqueeze(1)).long()          # skip the sample itself
    else:
        K = num_neg_trajs
    lens = hist_masks.long().sum(1, keepdim=True) - 1                                            # valid steps (CLS slot excluded)
    pos = torch.arange(hist_max_len, device=dev).unsqueeze(0).expand(batch_size, -1)
    shuffled = []
    for _ in range(K):
        keys = torch.rand(batch_size, hist_max_len, device=dev)
        keys = torch.where(pos < lens, keys, 2.0 + pos.float())                                  # padding keeps its place behind the valid steps
        shuffled.append(torch.argsort(keys, dim=1))
    return neg_idxs, shuffled


def _additive_mask(mask: torch.Tensor) -> torch.Tensor:
    """(1 - m) * -10000 as an fp32 row per sample (vilmodel.py:597-599); the kernels add it after the 1/sqrt(d) scale."""
    return ((1.0 - mask.to(torch.float32)) * -10000.0).contiguous()


def _feat16(x: torch.Tensor) -> torch.Tensor:
    """Precomputed view features arrive as fp32 (collate) or bf16 (device-resident feature store): one cast kernel."""
    x = x.reshape(-1, x.shape[-1])
    if x.dtype != BF16 and x.requires_grad:        # end-to-end stage: features computed by the ViT backbone inside this step
        return Fn.CastFn.apply(x.float().contiguous())
    return ops.cast_bf16(x) if x.dtype != BF16 else x.contiguous()


class NavPreTrainedModel(HamtPreTrainedModel):
    """Modification of LXMERT (vilmodel.py:578-724)."""

    def __init__(self, config):
        super().__init__(config)
        self.embeddings = BertEmbeddings(config)
        self.img_embeddings = ImageEmbeddings(config)
        self.hist_embeddings = HistoryEmbeddings(config)
        self.encoder = LxmertEncoder(config)
        self.init_weights()

    # ---- embedders ---------------------------------------------------------------------------
    def _text(self, run, txt_ids):
        return Fn.TextEmbedFn.apply(run.arena.anchor, run, self.embeddings, txt_ids.contiguous())

    def _pano_tokens(self, run, pano_img, pano_ang, drop_mod=None):
        """[N,P,F] views -> [N,H] fp32 mean of the pano-encoder outputs (vilmodel.py:553-564)."""
        he = self.hist_embeddings
        N, P = pano_img.shape[0], pano_img.shape[1]
        Pd = dict(img_linear=he.pano_img_linear, ang_linear=he.pano_ang_linear, ln_img=he.pano_img_layer_norm, ln_ang=he.pano_ang_layer_norm)
        e = Fn.FeatEmbedFn.apply(run.arena.anchor, None, run, Pd, _feat16(pano_img), pano_ang.reshape(N * P, -1).float().contiguous(), None, None, 1,
                                 drop_mod)
        e32 = None
        for layer in he.pano_encoder.layer:
            e, e32 = Fn.BertLayerFn.apply(run.arena.anchor, e, e32, run, layer, N, P, None)     # all-zero mask (vilmodel.py:560)
        return Fn.MeanPoolFn.apply(e, N, P)

    def _hist_cls(self, run, batch_size):
        he = self.hist_embeddings
        if run.training and run.save:
            # these two tiny parameters get their gradient through torch autograd: pre-attach the arena views so
            # AccumulateGrad adds in place into the flat gradient buffer
            run.arena.grad(he.cls_token), run.arena.grad(he.type_embedding.weight)
        x = (he.cls_token[0] + he.type_embedding.weight[:1]).expand(batch_size, -1)       # [B,H] fp32, autograd tracks both params
        return Fn.RowLNFn.apply(run.arena.anchor, x, run, he.layer_norm, he.dropout)

    def _hist_steps(self, run, hist_img, hist_ang, pano_img, pano_ang, with_pos: bool):
        """Per-step history embeddings [B*T,H] (vilmodel.py:548-571); with_pos False = the pre-position sum used by ITM."""
        he = self.hist_embeddings
        B, T = hist_img.shape[0], hist_img.shape[1]
        extra = None
        if he.pano_encoder is not None:
            extra = self._pano_tokens(run, pano_img.reshape(B * T, pano_img.shape[2], -1), pano_ang.reshape(B * T, pano_ang.shape[2], -1))
        Pd = dict(img_linear=he.img_linear, ang_linear=he.ang_linear, ln_img=he.img_layer_norm, ln_ang=he.ang_layer_norm,
                  add_vec=he.type_embedding.weight[0], add_vec_grad=lambda A: A.grad(he.type_embedding.weight)[0])
        if with_pos:
            Pd.update(pos_table=he.position_embeddings, ln_f=he.layer_norm)
        return Fn.FeatEmbedFn.apply(run.arena.anchor, extra, run, Pd, _feat16(hist_img), hist_ang.reshape(B * T, -1).float().contiguous(), None, None,
                                    T, he.dropout if with_pos else None)

    def _obs(self, run, ob_img, ob_ang, ob_nav_types):
        ie = self.img_embeddings
        B, O = ob_img.shape[0], ob_img.shape[1]
        tt = self.embeddings.token_type_embeddings
        Pd = dict(img_linear=ie.img_linear, ang_linear=ie.ang_linear, ln_img=ie.img_layer_norm, ln_ang=ie.ang_layer_norm,
                  add_vec=tt.weight[1], add_vec_grad=lambda A: A.grad(tt.weight)[1], nav_table=ie.nav_type_embedding, ln_f=ie.layer_norm)
        nav = ob_nav_types.reshape(-1).contiguous() if ob_nav_types is not None else None
        return Fn.FeatEmbedFn.apply(run.arena.anchor, None, run, Pd, _feat16(ob_img), ob_ang.reshape(B * O, -1).float().contiguous(), nav, None, 1,
                                    ie.dropout)

    # ---- encoder -----------------------------------------------------------------------------
    def _text_branch(self, run, txt_ids, B, L, txt_mask):
        """Text embedder + text layers."""
        return self._text_layers(run, self._text(run, txt_ids), B, L, txt_mask)

    def _text_layers(self, run, txt, B, L, txt_mask):
        """-> (txt bf16, txt32: its fp32 twin from the last LayerNorm, or None without layers)."""
        txt32 = None
        for layer in self.encoder.layer:
            txt, txt32 = Fn.BertLayerFn.apply(run.arena.anchor, txt, txt32, run, layer, B, L, txt_mask)
        if not self.encoder.update_lang_bert:
            txt = txt.detach()
        return txt, txt32

    def _x_layers(self, run, txt, txt32, visn, B, L, V, txt_mask, visn_mask):
        xcat = torch.cat([txt, visn], 0)
        xcat32 = torch.cat([Fn._as32(txt, txt32), visn.detach().float()], 0)
        for layer in self.encoder.x_layers:
            xcat, xcat32 = Fn.XLayerFn.apply(run.arena.anchor, xcat, xcat32, run, layer, B, L, V, txt_mask, visn_mask, True)
        return xcat[:B * L], xcat[B * L:]

    def forward(self, txt_ids, txt_masks, hist_img_feats, hist_ang_feats, hist_pano_img_feats, hist_pano_ang_feats, hist_masks,
                ob_img_feats, ob_ang_feats, ob_nav_types, ob_masks, _run=None):
        """vilmodel.py:591-638.  Returns (txt_embeds [B,L,H], hist_embeds [B,T+1,H], ob_embeds [B,O,H] or None) in bf16."""
        run = _run or self.begin()
        B, L = txt_ids.shape
        H = self.config.hidden_size
        txt_mask = _additive_mask(txt_masks)
        hist_mask = _additive_mask(hist_masks)
        cls = self._hist_cls(run, B)
        if hist_img_feats is not None:
            T = hist_img_feats.shape[1]
            vp = self._hist_steps(run, hist_img_feats, hist_ang_feats, hist_pano_img_feats, hist_pano_ang_feats, with_pos=True)
            hist = torch.cat([cls.view(B, 1, H), vp.view(B, T, H)], 1)
        else:
            T = 0
            hist = cls.view(B, 1, H)
        if ob_img_feats is not None:
            O = ob_img_feats.shape[1]
            ob = self._obs(run, ob_img_feats, ob_ang_feats, ob_nav_types).view(B, O, H)
            ob_mask = _additive_mask(ob_masks)
        else:
            O, ob, ob_mask = 0, None, None

        # The text embedder is created AFTER the history / observation embedders so that autograd runs its backward right after
        # the first text layer's -- before the panorama encoder's backward, not as the very last node of the step: its 94 MB
        # word-embedding gradient is the largest slice of the data-parallel exchange and now overlaps ~2.5 ms of remaining backward
        # work instead of being exposed at the end.
        txt, txt32 = self._text_branch(run, txt_ids, B, L, txt_mask)
        if ob is not None and self.encoder.r_layers is not None:
            o2, o32 = ob.reshape(B * O, H), None
            for layer in self.encoder.r_layers:
                o2, o32 = Fn.BertLayerFn.apply(run.arena.anchor, o2, o32, run, layer, B, O, ob_mask)
            ob = o2.view(B, O, H)
        if self.encoder.h_layers is not None:
            h2, h32 = hist.reshape(B * (T + 1), H).contiguous(), None
            for layer in self.encoder.h_layers:
                h2, h32 = Fn.BertLayerFn.apply(run.arena.anchor, h2, h32, run, layer, B, T + 1, hist_mask)
            hist = h2.view(B, T + 1, H)
        if ob is None:
            visn, visn_mask, V = hist, hist_mask, T + 1
        else:
            visn, visn_mask, V = torch.cat([hist, ob], 1), torch.cat([hist_mask, ob_mask], -1).contiguous(), T + 1 + O
        txt, visn = self._x_layers(run, txt, txt32, visn.reshape(B * V, H), B, L, V, txt_mask, visn_mask)
        txt = txt.view(B, L, H)
        visn = visn.view(B, V, H)
        hist_out = visn[:, :T + 1]
        ob_out = visn[:, T + 1:] if ob is not None else None
        return txt, hist_out, ob_out

    def forward_itm(self, txt_ids, txt_masks, hist_img_feats, hist_ang_feats, hist_pano_img_feats, hist_pano_ang_feats, hist_masks,
                    num_neg_trajs=4, _run=None, _plan=None):
        """vilmodel.py:640-724.  The negative-trajectory indices are drawn on the host from the global numpy / torch
        RNGs in the reference's exact call order (np.random.choice per sample, then torch.randperm per sample per K)."""
        run = _run or self.begin()
        he = self.hist_embeddings
        B, T = hist_img_feats.shape[0], hist_img_feats.shape[1]
        L, H = txt_ids.shape[1], self.config.hidden_size
        R = 1 + num_neg_trajs
        txt_mask = _additive_mask(txt_masks)
        hist_mask = _additive_mask(hist_masks)
        cls = self._hist_cls(run, B).view(B, 1, H)
        nopos = self._hist_steps(run, hist_img_feats, hist_ang_feats, hist_pano_img_feats, hist_pano_ang_feats, with_pos=False).view(B, T, H)
        # text side after the history embedder (see forward(): backward order / gradient-exchange overlap)
        txt, txt32 = self._text_branch(run, txt_ids, B, L, txt_mask)
        txt = txt.view(B, L, H).repeat(R, 1, 1).reshape(R * B * L, H)
        if txt32 is not None:
            txt32 = txt32.view(B, L, H).repeat(R, 1, 1).reshape(R * B * L, H)
        txt_mask_r = txt_mask.repeat(R, 1).contiguous()

        pos_w = he.position_embeddings.weight
        if run.training and run.save:
            run.arena.grad(pos_w)          # gathered through torch autograd below: accumulate in place into the arena

        def with_pos(pos_ids):
            x = nopos.float() + pos_w[pos_ids]                                        # [B,T,H] fp32 (autograd: gather on the table)
            return Fn.RowLNFn.apply(run.arena.anchor, x.reshape(B * T, H), run, he.layer_norm, he.dropout).view(B, T, H)

        def h_layers(x):
            if self.encoder.h_layers is None:
                return x
            x2, x32 = x.reshape(B * (T + 1), H).contiguous(), None
            for layer in self.encoder.h_layers:
                x2, x32 = Fn.BertLayerFn.apply(run.arena.anchor, x2, x32, run, layer, B, T + 1, hist_mask)
            return x2.view(B, T + 1, H)

        dev = txt_ids.device
        hist = h_layers(torch.cat([cls, with_pos(torch.arange(T, device=dev).expand(B, -1))], 1))
        neg_embeds, neg_masks = [], []
        if _plan is None and getattr(self.config, "itm_device_negatives", False):
            _plan = itm_negative_plan_device(B, hist_masks, T, num_neg_trajs)
        elif _plan is None:
            _plan = itm_negative_plan(B, hist_masks, T, num_neg_trajs)      # host RNG draws in the reference's order (+1 host sync)
            _plan = (None if _plan[0] is None else _plan[0].to(dev), [t.to(dev) for t in _plan[1]])
        neg_idxs, shuffled = _plan
        if neg_idxs is not None:
            for k in range(neg_idxs.shape[1]):
                neg_embeds.append(hist[neg_idxs[:, k]])
                neg_masks.append(hist_mask[neg_idxs[:, k]])
        for pos_ids in shuffled:
            neg_embeds.append(h_layers(torch.cat([cls, with_pos(pos_ids)], 1)))
            neg_masks.append(hist_mask)
        visn = torch.cat([hist] + neg_embeds, 0)                                       # [R*B, T+1, H]
        visn_mask = torch.cat([hist_mask] + neg_masks, 0).contiguous()
        txt, visn = self._x_layers(run, txt, txt32, visn.reshape(R * B * (T + 1), H), R * B, L, T + 1, txt_mask_r, visn_mask)
        fused = Fn.MulRowsFn.apply(txt.view(R * B, L, H)[:, 0].contiguous(), visn.view(R * B, T + 1, H)[:, 0].contiguous(), R * B, 1)
        return torch.stack(torch.split(fused, B), 1)                                   # [B, R, H]
