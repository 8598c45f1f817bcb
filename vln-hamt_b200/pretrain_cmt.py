"""B200-native mirror of ``pretrain_src/model/pretrain_cmt.py``: the six proxy-task heads and
``MultiStepNavCMTPreTraining.forward(batch, task, compute_loss=True)`` (pretrain_cmt.py:101-140).

Return conventions are the reference's: un-reduced loss vectors (``reduction='none'``) or the logits
when ``compute_loss`` is False; SAP logits carry ``-inf`` where ``ob_nav_types == 0``.  Logits /
losses are fp32; the hidden states feeding the heads are bf16.
"""
from __future__ import annotations

from collections import defaultdict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn
from . import ops
from .vilmodel import BertLayerNorm, BertOnlyMLMHead, HamtPreTrainedModel, NavPreTrainedModel, _Container

BF16 = torch.bfloat16


class _HeadMLP(_Container):
    """Linear -> ReLU -> LayerNorm(1e-12) -> [Dropout] -> Linear  (pretrain_cmt.py:13-71).  ``self.net`` keeps the
    reference's nn.Sequential indices so the state_dict keys match (net.0 / net.2 / net.4, or net.3 without dropout)."""

    def _run(self, run: Fn.Run, x: torch.Tensor) -> torch.Tensor:
        """x: bf16 [M, in]; returns fp32 logits [M, out]."""
        anchor = run.arena.anchor
        mods = list(self.net)
        lin0, ln = mods[0], mods[2]
        has_drop = isinstance(mods[3], nn.Dropout)
        last = mods[4] if has_drop else mods[3]
        h = Fn.LinearFn.apply(anchor, x.contiguous(), run, lin0, ops.ACT_RELU, False, True)
        h = Fn.LayerNormFn.apply(anchor, h, run, ln)
        if has_drop:
            h = Fn.dropout(run, h, mods[3])
        if last.out_features <= 4:
            return Fn.RowdotFn.apply(anchor, h, run, last)
        return Fn.LinearFn.apply(anchor, h, run, last, ops.ACT_NONE, True, True)


class NextActionPrediction(_HeadMLP):
    def __init__(self, hidden_size, dropout_rate):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.ReLU(), BertLayerNorm(hidden_size, eps=1e-12),
                                 nn.Dropout(dropout_rate), nn.Linear(hidden_size, 1))


class NextActionRegression(_HeadMLP):
    def __init__(self, hidden_size, dropout_rate):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.ReLU(), BertLayerNorm(hidden_size, eps=1e-12),
                                 nn.Dropout(dropout_rate), nn.Linear(hidden_size, 3))


class SpatialRelRegression(_HeadMLP):
    def __init__(self, hidden_size, dropout_rate):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(hidden_size * 2, hidden_size), nn.ReLU(), BertLayerNorm(hidden_size, eps=1e-12),
                                 nn.Dropout(dropout_rate), nn.Linear(hidden_size, 2))


class RegionClassification(_HeadMLP):
    " for MRC(-kl)"

    def __init__(self, hidden_size, label_dim):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.ReLU(), BertLayerNorm(hidden_size, eps=1e-12),
                                 nn.Linear(hidden_size, label_dim))


class ItmPrediction(_HeadMLP):
    def __init__(self, hidden_size):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.ReLU(), BertLayerNorm(hidden_size, eps=1e-12),
                                 nn.Linear(hidden_size, 1))


class _TiedDecoder:
    """weight = decoder.weight (tied to word_embeddings, pretrain_cmt.py:96-99), bias = predictions.bias (vilmodel.py:280-284)."""

    def __init__(self, pred):
        self.weight, self.bias = pred.decoder.weight, pred.bias


class MultiStepNavCMTPreTraining(HamtPreTrainedModel):
    def __init__(self, config):
        super().__init__(config)
        self.config = config
        self.bert = self._make_backbone(config)
        self.bert._arena_owner = None
        object.__setattr__(self.bert, "_arena_owner", self)       # one arena for backbone + heads (not a submodule cycle)
        vb = getattr(self.bert, "vision_backbone", None)          # end-to-end stage (image_pretrain.py): the ViT shares it too
        if vb is not None:
            object.__setattr__(vb, "_arena_owner", self)
        if 'mlm' in config.pretrain_tasks:
            self.mlm_head = BertOnlyMLMHead(self.config)
        if 'sap' in config.pretrain_tasks:
            self.next_action = NextActionPrediction(self.config.hidden_size, self.config.pred_head_dropout_prob)
        if 'sar' in config.pretrain_tasks:
            self.regress_action = NextActionRegression(self.config.hidden_size, self.config.pred_head_dropout_prob)
        if 'sprel' in config.pretrain_tasks:
            self.sprel_head = SpatialRelRegression(self.config.hidden_size, self.config.pred_head_dropout_prob)
        if 'mrc' in config.pretrain_tasks:
            self.image_classifier = RegionClassification(self.config.hidden_size, self.config.image_prob_size)
        if 'itm' in config.pretrain_tasks:
            self.itm_head = ItmPrediction(self.config.hidden_size)
        self.init_weights()
        self.tie_weights()

    def _make_backbone(self, config):
        return NavPreTrainedModel(config)

    def tie_weights(self):
        if 'mlm' in self.config.pretrain_tasks:
            self._tie_or_clone_weights(self.mlm_head.predictions.decoder, self.bert.embeddings.word_embeddings)

    def forward(self, batch, task, compute_loss=True):
        batch = defaultdict(lambda: None, batch)
        hist = (batch['hist_img_fts'], batch['hist_ang_fts'], batch['hist_pano_img_fts'], batch['hist_pano_ang_fts'], batch['hist_masks'])
        ob = (batch['ob_img_fts'], batch['ob_ang_fts'], batch['ob_nav_types'], batch['ob_masks'])
        # optional sync-free extras (graph.py): precomputed row indices of the masked tokens / regions and the ITM negative plan
        self._rows = batch['txt_label_rows'] if task.startswith('mlm') else batch['hist_mrc_rows']
        self._itm_plan = batch['itm_plan']
        if task.startswith('mlm'):
            return self.forward_mlm(batch['txt_ids'], batch['txt_masks'], *hist, batch['txt_labels'], compute_loss)
        elif task.startswith('sap'):
            return self.forward_sap(batch['txt_ids'], batch['txt_masks'], *hist, *ob, batch['ob_action_viewindex'], compute_loss)
        elif task.startswith('sar'):
            return self.forward_sar(batch['txt_ids'], batch['txt_masks'], *hist, *ob, batch['ob_action_angles'], batch['ob_progress'], compute_loss)
        elif task.startswith('sprel'):
            return self.forward_sprel(batch['txt_ids'], batch['txt_masks'], *hist, *ob, batch['sp_anchor_idxs'], batch['sp_targets'], compute_loss)
        elif task.startswith('mrc'):
            return self.forward_mrc(batch['txt_ids'], batch['txt_masks'], *hist, batch['hist_mrc_masks'], batch['hist_img_probs'], compute_loss)
        elif task.startswith('itm'):
            return self.forward_itm(batch['txt_ids'], batch['txt_masks'], *hist, 4, compute_loss)
        else:
            raise ValueError('invalid task')

    # ------------------------------------------------------------------------------------------
    def _row_index(self, mask: torch.Tensor) -> torch.Tensor:
        """Row indices of `mask` (pretrain_cmt.py:161-165).  One nonzero() = one host sync, as in the reference's `hidden[mask]`,
        unless the caller supplied them with the batch (`txt_label_rows` / `hist_mrc_rows`) -- the CUDA-graph path does."""
        rows = getattr(self, "_rows", None)
        if rows is not None:
            return rows
        return torch.nonzero(mask.reshape(-1), as_tuple=False).squeeze(1)

    def _masked_rows(self, hidden: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        return Fn.GatherRowsFn.apply(hidden.reshape(-1, hidden.shape[-1]), idx)

    def forward_mlm(self, txt_ids, txt_masks, hist_img_fts, hist_ang_fts, hist_pano_img_fts, hist_pano_ang_fts, hist_masks, txt_labels, compute_loss):
        run = self.begin()
        txt_embeds, _, _ = self.bert(txt_ids, txt_masks, hist_img_fts, hist_ang_fts, hist_pano_img_fts, hist_pano_ang_fts, hist_masks,
                                     None, None, None, None, _run=run)
        idx = self._row_index(txt_labels != -1)
        masked_output = self._masked_rows(txt_embeds, idx)
        pred = self.mlm_head.predictions
        anchor = run.arena.anchor
        h = Fn.LinearFn.apply(anchor, masked_output, run, pred.transform.dense, ops.ACT_GELU, False, True)
        h = Fn.LayerNormFn.apply(anchor, h, run, pred.transform.LayerNorm)
        prediction_scores = Fn.LinearFn.apply(anchor, h, run, _TiedDecoder(pred), ops.ACT_NONE, True, True)
        if compute_loss:
            return Fn.CrossEntropyFn.apply(prediction_scores, txt_labels.reshape(-1)[idx])
        return prediction_scores

    def forward_sap(self, txt_ids, txt_masks, hist_img_fts, hist_ang_fts, hist_pano_img_fts, hist_pano_ang_fts, hist_masks,
                    ob_img_fts, ob_ang_fts, ob_nav_types, ob_masks, act_labels, compute_loss):
        run = self.begin()
        txt_embeds, hist_embeds, ob_embeds = self.bert(txt_ids, txt_masks, hist_img_fts, hist_ang_fts, hist_pano_img_fts, hist_pano_ang_fts,
                                                       hist_masks, ob_img_fts, ob_ang_fts, ob_nav_types, ob_masks, _run=run)
        B, O, H = ob_embeds.shape
        fused = Fn.MulRowsFn.apply(ob_embeds.reshape(B * O, H), txt_embeds[:, 0], B, O)          # ob * txt[:, :1]
        prediction_scores = self.next_action._run(run, fused).view(B, O)
        prediction_scores = prediction_scores.masked_fill(ob_nav_types == 0, -float('inf'))
        if compute_loss:
            return Fn.CrossEntropyFn.apply(prediction_scores, act_labels)
        return prediction_scores

    def forward_sar(self, txt_ids, txt_masks, hist_img_fts, hist_ang_fts, hist_pano_img_fts, hist_pano_ang_fts, hist_masks,
                    ob_img_fts, ob_ang_fts, ob_nav_types, ob_masks, ob_act_angles, ob_progress, compute_loss):
        run = self.begin()
        txt_embeds, hist_embeds, ob_embeds = self.bert(txt_ids, txt_masks, hist_img_fts, hist_ang_fts, hist_pano_img_fts, hist_pano_ang_fts,
                                                       hist_masks, ob_img_fts, ob_ang_fts, ob_nav_types, ob_masks, _run=run)
        prediction_scores = self.regress_action._run(run, txt_embeds[:, 0].contiguous())       # [CLS] token
        if compute_loss:
            act_targets = torch.cat([ob_act_angles, ob_progress.unsqueeze(1)], dim=1)
            return F.mse_loss(prediction_scores, act_targets.float(), reduction='none')
        return prediction_scores

    def forward_sprel(self, txt_ids, txt_masks, hist_img_fts, hist_ang_fts, hist_pano_img_fts, hist_pano_ang_fts, hist_masks,
                      ob_img_fts, ob_ang_fts, ob_nav_types, ob_masks, sp_anchor_idxs, sp_targets, compute_loss):
        run = self.begin()
        txt_embeds, hist_embeds, ob_embeds = self.bert(txt_ids, txt_masks, hist_img_fts, hist_ang_fts, hist_pano_img_fts, hist_pano_ang_fts,
                                                       hist_masks, ob_img_fts, ob_ang_fts, ob_nav_types, ob_masks, _run=run)
        B, O, H = ob_embeds.shape
        # reference: torch.gather(ob_embeds, 1, anchor.unsqueeze(1).unsqueeze(2).repeat(1, 36, H)) (pretrain_cmt.py:202) -- the same values
        # as one row per sample broadcast over the 36 views; written this way the backward is a 36-way sum + a collision-free row
        # scatter instead of gather's bf16 atomic scatter_add (order-dependent rounding: the only run-to-run noise above fp32 level)
        anchor_ob_embeds = ob_embeds[torch.arange(B, device=ob_embeds.device), sp_anchor_idxs].unsqueeze(1).expand(B, 36, H)
        cat_ob_embeds = torch.cat([anchor_ob_embeds, ob_embeds[:, :-1]], -1)                   # (batch, 36, 2H)
        prediction_scores = self.sprel_head._run(run, cat_ob_embeds.reshape(B * 36, 2 * H)).view(B, 36, 2)
        if compute_loss:
            return F.mse_loss(prediction_scores, sp_targets.float(), reduction='none')
        return prediction_scores

    def forward_mrc(self, txt_ids, txt_masks, hist_img_fts, hist_ang_fts, hist_pano_img_fts, hist_pano_ang_fts, hist_masks,
                    hist_mrc_masks, hist_img_probs, compute_loss=True):
        run = self.begin()
        txt_embeds, hist_embeds, _ = self.bert(txt_ids, txt_masks, hist_img_fts, hist_ang_fts, hist_pano_img_fts, hist_pano_ang_fts, hist_masks,
                                               None, None, None, None, _run=run)
        hist_embeds = hist_embeds[:, 1:]                                                        # remove global embedding
        idx = self._row_index(hist_mrc_masks)
        masked_output = self._masked_rows(hist_embeds.contiguous(), idx)
        prediction_soft_labels = self.image_classifier._run(run, masked_output)
        hist_mrc_targets = hist_img_probs.reshape(-1, hist_img_probs.shape[-1])[idx].float()
        if compute_loss:
            prediction_soft_labels = F.log_softmax(prediction_soft_labels, dim=-1)
            return F.kl_div(prediction_soft_labels, hist_mrc_targets, reduction='none').sum(dim=1)
        return prediction_soft_labels, hist_mrc_targets

    def forward_itm(self, txt_ids, txt_masks, hist_img_fts, hist_ang_fts, hist_pano_img_fts, hist_pano_ang_fts, hist_masks,
                    num_neg_trajs, compute_loss):
        run = self.begin()
        fused_embeds = self.bert.forward_itm(txt_ids, txt_masks, hist_img_fts, hist_ang_fts, hist_pano_img_fts, hist_pano_ang_fts, hist_masks,
                                             num_neg_trajs=num_neg_trajs, _run=run, _plan=getattr(self, "_itm_plan", None))
        B, R, H = fused_embeds.shape
        prediction_scores = self.itm_head._run(run, fused_embeds.reshape(B * R, H)).view(B, R)
        itm_targets = torch.zeros(B, dtype=torch.long, device=fused_embeds.device)
        if compute_loss:
            return Fn.CrossEntropyFn.apply(prediction_scores, itm_targets)
        return prediction_scores, itm_targets
