"""B200-native mirror of the finetune backbone ``finetune_src/models/vilmodel_cmt.py``.

``NavCMT.forward(mode, ...)`` keeps the reference signature (vilmodel_cmt.py:624-629) and its three
modes: 'language' (once per episode), 'history' (one step of the hierarchical encoder), 'visual'
(cross-modal layers + action logits).  Block classes are shared with vilmodel.py (identical state_dict
keys); only the pieces whose constructor / registration order differs in the finetune file are
redefined here.
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from . import functional as Fn
from . import ops
from .vilmodel import (BertAttention, BertEmbeddings, BertEncoder, BertIntermediate, BertLayer, BertLayerNorm, BertOutput,  # noqa: F401
                       BertXAttention, HamtPreTrainedModel, ImageEmbeddings, _Container, _additive_mask, _feat16)
from .pretrain_cmt import NextActionPrediction

BF16 = torch.bfloat16


class LXRTXLayer(_Container):
    """vilmodel_cmt.py:361-424 (adds no_lang_ca)."""

    def __init__(self, config):
        super().__init__()
        self.no_lang_ca = config.no_lang_ca
        self.lang_self_att = BertAttention(config)
        self.lang_inter = BertIntermediate(config)
        self.lang_output = BertOutput(config)
        self.visn_self_att = BertAttention(config)
        self.visn_inter = BertIntermediate(config)
        self.visn_output = BertOutput(config)
        self.visual_attention = BertXAttention(config)


class LxmertEncoder(_Container):
    """vilmodel_cmt.py:426-491 (text layers frozen when update_lang_bert is False, :440-442)."""

    def __init__(self, config):
        super().__init__()
        self.num_l_layers = config.num_l_layers
        self.num_r_layers = config.num_r_layers
        self.num_h_layers = config.num_h_layers
        self.num_x_layers = config.num_x_layers
        self.update_lang_bert = config.update_lang_bert
        self.layer = nn.ModuleList([BertLayer(config) for _ in range(self.num_l_layers)])
        if not self.update_lang_bert:
            for _, param in self.layer.named_parameters():
                param.requires_grad = False
        self.h_layers = nn.ModuleList([BertLayer(config) for _ in range(self.num_h_layers)]) if self.num_h_layers > 0 else None
        self.r_layers = nn.ModuleList([BertLayer(config) for _ in range(self.num_r_layers)]) if self.num_r_layers > 0 else None
        self.x_layers = nn.ModuleList([LXRTXLayer(config) for _ in range(self.num_x_layers)])


class HistoryEmbeddings(_Container):
    """vilmodel_cmt.py:523-594 (single-step variant; registration order differs from the pretrain class)."""

    def __init__(self, config):
        super().__init__()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, config.hidden_size))
        self.img_linear = nn.Linear(config.image_feat_size, config.hidden_size)
        self.img_layer_norm = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.ang_linear = nn.Linear(config.angle_feat_size, config.hidden_size)
        self.ang_layer_norm = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.position_embeddings = nn.Embedding(config.max_action_steps, config.hidden_size)
        self.type_embedding = nn.Embedding(1, config.hidden_size)
        self.layer_norm = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.hist_enc_pano = config.hist_enc_pano
        if config.hist_enc_pano:
            self.pano_img_linear = nn.Linear(config.image_feat_size, config.hidden_size)
            self.pano_img_layer_norm = BertLayerNorm(config.hidden_size, eps=1e-12)
            self.pano_ang_linear = nn.Linear(config.angle_feat_size, config.hidden_size)
            self.pano_ang_layer_norm = BertLayerNorm(config.hidden_size, eps=1e-12)
            pano_enc_config = copy.copy(config)
            pano_enc_config.num_hidden_layers = config.num_h_pano_layers
            self.pano_encoder = BertEncoder(pano_enc_config)
        else:
            self.pano_encoder = None


class NavCMT(HamtPreTrainedModel):
    """vilmodel_cmt.py:610-728"""

    def __init__(self, config):
        super().__init__(config)
        self.embeddings = BertEmbeddings(config)
        self.img_embeddings = ImageEmbeddings(config)
        self.hist_embeddings = HistoryEmbeddings(config)
        self.encoder = LxmertEncoder(config)
        self.next_action = NextActionPrediction(config.hidden_size, config.pred_head_dropout_prob)
        self.init_weights()

    # ------------------------------------------------------------------------------------------
    def _language(self, run, txt_ids, txt_masks):
        B, L = txt_ids.shape
        H = self.config.hidden_size
        anchor = run.arena.anchor
        m = _additive_mask(txt_masks)
        txt = Fn.TextEmbedFn.apply(anchor, run, self.embeddings, txt_ids.contiguous())
        txt32 = None
        for layer in self.encoder.layer:
            txt, txt32 = Fn.BertLayerFn.apply(anchor, txt, txt32, run, layer, B, L, m)
        if self.config.fix_lang_embedding:
            txt = txt.detach()
        if self.config.no_lang_ca:                      # run the language self-attention stacks of the x-layers
            outs = [txt.view(B, L, H)]
            for layer in self.encoder.x_layers:
                outs.append(Fn.LangSelfFn.apply(anchor, txt, txt32, run, layer, B, L, m).view(B, L, H))
            return outs
        return txt.view(B, L, H)

    def _history(self, run, hist_img_feats, hist_ang_feats, ob_step_ids, pano_img_feats, pano_ang_feats):
        he = self.hist_embeddings
        anchor = run.arena.anchor
        if hist_img_feats is None:
            if self.training and torch.is_grad_enabled():
                run.arena.grad(he.cls_token), run.arena.grad(he.type_embedding.weight)
            x = he.cls_token[0] + he.type_embedding.weight[:1]                      # batch_size forced to 1 (vilmodel_cmt.py:561-572)
            out = Fn.RowLNFn.apply(anchor, x, run, he.layer_norm, he.dropout)
        else:
            B = hist_img_feats.shape[0]
            extra = None
            if he.pano_encoder is not None:
                P = pano_img_feats.shape[1]
                Pd = dict(img_linear=he.pano_img_linear, ang_linear=he.pano_ang_linear, ln_img=he.pano_img_layer_norm, ln_ang=he.pano_ang_layer_norm)
                e = Fn.FeatEmbedFn.apply(anchor, None, run, Pd, _feat16(pano_img_feats), pano_ang_feats.reshape(B * P, -1).float().contiguous(), None,
                                         None, 1, he.dropout)                       # finetune drops the pano token embeddings (:583)
                e32 = None
                for layer in he.pano_encoder.layer:
                    e, e32 = Fn.BertLayerFn.apply(anchor, e, e32, run, layer, B, P, None)
                extra = Fn.MeanPoolFn.apply(e, B, P)
            pos_ids = ob_step_ids.reshape(-1).to(hist_img_feats.device)
            if pos_ids.numel() == 1:
                pos_ids = pos_ids.expand(B)
            Pd = dict(img_linear=he.img_linear, ang_linear=he.ang_linear, ln_img=he.img_layer_norm, ln_ang=he.ang_layer_norm,
                      add_vec=he.type_embedding.weight[0], add_vec_grad=lambda A: A.grad(he.type_embedding.weight)[0],
                      pos_table=he.position_embeddings, ln_f=he.layer_norm)
            out = Fn.FeatEmbedFn.apply(anchor, extra, run, Pd, _feat16(hist_img_feats), hist_ang_feats.float().contiguous(), None,
                                       pos_ids.contiguous(), 1, he.dropout)
        if self.config.fix_hist_embedding:
            out = out.detach()
        return out

    def _visual(self, run, txt_embeds, txt_masks, hist_embeds, hist_masks, ob_img_feats, ob_ang_feats, ob_nav_types, ob_masks):
        cfg = self.config
        anchor = run.arena.anchor
        H = cfg.hidden_size
        B, T1 = hist_embeds.shape[0], hist_embeds.shape[1]
        hist_mask = _additive_mask(hist_masks)
        hist = hist_embeds.to(BF16)
        if self.encoder.h_layers is not None:
            h2, h32 = hist.reshape(B * T1, H).contiguous(), None
            for layer in self.encoder.h_layers:
                h2, h32 = Fn.BertLayerFn.apply(anchor, h2, h32, run, layer, B, T1, hist_mask)
            hist = h2.view(B, T1, H)
        O = ob_img_feats.shape[1]
        ob_mask = _additive_mask(ob_masks)
        ie, tt = self.img_embeddings, self.embeddings.token_type_embeddings
        Pd = dict(img_linear=ie.img_linear, ang_linear=ie.ang_linear, ln_img=ie.img_layer_norm, ln_ang=ie.ang_layer_norm,
                  add_vec=tt.weight[1], add_vec_grad=lambda A: A.grad(tt.weight)[1], nav_table=ie.nav_type_embedding, ln_f=ie.layer_norm)
        ob = Fn.FeatEmbedFn.apply(anchor, None, run, Pd, _feat16(ob_img_feats), ob_ang_feats.reshape(B * O, -1).float().contiguous(),
                                  ob_nav_types.reshape(-1).contiguous(), None, 1, ie.dropout)
        if self.encoder.r_layers is not None:
            ob32 = None
            for layer in self.encoder.r_layers:
                ob, ob32 = Fn.BertLayerFn.apply(anchor, ob, ob32, run, layer, B, O, ob_mask)
        if cfg.fix_obs_embedding:
            ob = ob.detach()
        V = T1 + O
        visn = torch.cat([hist, ob.view(B, O, H)], 1).reshape(B * V, H)
        visn_mask = torch.cat([hist_mask, ob_mask], -1).contiguous()
        txt_mask = _additive_mask(txt_masks)
        all_txt = txt_embeds if cfg.no_lang_ca else None
        txt = None if cfg.no_lang_ca else txt_embeds.to(BF16)
        L = (all_txt[0] if cfg.no_lang_ca else txt).shape[1]
        txt32 = visn32 = None                              # fp32 twins of the two streams between x-layers
        for l, layer in enumerate(self.encoder.x_layers):
            if cfg.no_lang_ca:
                txt, txt32 = all_txt[l].to(BF16), None
            xcat = torch.cat([txt.reshape(B * L, H), visn], 0)
            xcat32 = torch.cat([Fn._as32(txt.reshape(B * L, H), txt32), Fn._as32(visn, visn32)], 0)
            xcat, xcat32 = Fn.XLayerFn.apply(anchor, xcat, xcat32, run, layer, B, L, V, txt_mask, visn_mask, not cfg.no_lang_ca)
            txt, visn = xcat[:B * L].view(B, L, H), xcat[B * L:]
            txt32, visn32 = xcat32[:B * L], xcat32[B * L:]
        visn = visn.view(B, V, H)
        hist_out, ob_out = visn[:, :T1], visn[:, T1:]
        if cfg.no_lang_ca or cfg.act_pred_token == 'ob':
            fused = ob_out.reshape(B * O, H).contiguous()
        elif cfg.act_pred_token == 'ob_txt':
            fused = Fn.MulRowsFn.apply(ob_out.reshape(B * O, H), txt[:, 0], B, O)
        elif cfg.act_pred_token == 'ob_hist':
            fused = Fn.MulRowsFn.apply(ob_out.reshape(B * O, H), hist_out[:, 0], B, O)
        elif cfg.act_pred_token == 'ob_txt_hist':
            fused = Fn.MulRowsFn.apply(ob_out.reshape(B * O, H), (txt[:, 0].float() + hist_out[:, 0].float()).to(BF16), B, O)
        else:
            raise ValueError(cfg.act_pred_token)
        act_logits = self.next_action._run(run, fused).view(B, O)
        act_logits = act_logits.masked_fill(ob_nav_types == 0, -float('inf'))
        return act_logits, txt, hist_out, ob_out

    def forward(self, mode, txt_ids=None, txt_embeds=None, txt_masks=None,
                hist_img_feats=None, hist_ang_feats=None,
                hist_pano_img_feats=None, hist_pano_ang_feats=None,
                hist_embeds=None, ob_step_ids=None, hist_masks=None,
                ob_img_feats=None, ob_ang_feats=None, ob_nav_types=None,
                ob_masks=None):
        run = self.begin()
        if mode == 'language':
            return self._language(run, txt_ids, txt_masks)
        if mode == 'history':
            return self._history(run, hist_img_feats, hist_ang_feats, ob_step_ids, hist_pano_img_feats, hist_pano_ang_feats)
        elif mode == 'visual':
            return self._visual(run, txt_embeds, txt_masks, hist_embeds, hist_masks, ob_img_feats, ob_ang_feats, ob_nav_types, ob_masks)
        raise ValueError(f'invalid mode {mode!r}')
