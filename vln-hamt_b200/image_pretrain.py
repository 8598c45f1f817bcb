"""End-to-end pretraining model: the six proxy-task heads over the image model.

Mirrors `MultiStepNavImagePreTraining` (pretrain_src/model/image_pretrain.py:18-171): constructor `(config)`, `forward(batch, task,
compute_loss)` with the image batch keys (`hist_images` [B,T,3,224,224], `hist_pano_images` [B,T,36,3,224,224], `ob_images`
[B,O-1,3,224,224], `ob_v_exists`), same state_dict keys (`bert.vision_backbone.*`, `bert.embeddings.*`, ..., head modules).  (The
reference file itself does not import as shipped -- `from .pretrain import ...` names a module that does not exist, SURVEY row 10 --
so the behaviour restated here is: backbone features (image_vilmodel.py:40-59, :75-82, :98-106) -> the heads / losses of
pretrain_cmt.py, which image_pretrain.py repeats verbatim.)

One Run spans the step: the view features are computed by the ViT backbone first (panorama views under no_grad), then handed to the
feature path of the parent class; the gradient reaches the backbone through the image-feature GEMMs' dgrad.
"""
from __future__ import annotations

from collections import defaultdict

import torch

from .image_vilmodel import NavImagePreTrainedModel
from .pretrain_cmt import MultiStepNavCMTPreTraining
from .vilmodel import NavPreTrainedModel


class MultiStepNavImagePreTraining(MultiStepNavCMTPreTraining):
    def __init__(self, config, vit_depth: int = 12):
        self._vit_depth = vit_depth
        super().__init__(config)

    def _make_backbone(self, config):
        return NavImagePreTrainedModel(config, vit_depth=self._vit_depth)

    def forward(self, batch, task, compute_loss=True):
        b = defaultdict(lambda: None, batch)
        if b['hist_images'] is None and b['ob_images'] is None:
            return super().forward(batch, task, compute_loss)            # feature batch: the parent's path unchanged
        run = self.bert.begin()
        fb = {k: v for k, v in batch.items() if k not in ('hist_images', 'hist_pano_images', 'ob_images', 'ob_v_exists')}
        if b['hist_images'] is not None:
            hf = self.bert.forward_vision_backbone(b['hist_images'], _run=run)
            pf = self.bert.forward_vision_backbone(b['hist_pano_images'], detach=True, _run=run)
            if task.startswith('mrc') and b['hist_mrc_masks'] is not None:      # image_vilmodel.py:80-82
                hf = hf.masked_fill(b['hist_mrc_masks'].unsqueeze(-1), 0)
                pf = pf.masked_fill(b['hist_mrc_masks'].unsqueeze(-1).unsqueeze(-1), 0)
            fb['hist_img_fts'], fb['hist_pano_img_fts'] = hf, pf
        if b['ob_images'] is not None:
            of = self.bert.forward_vision_backbone(b['ob_images'], _run=run)
            if b['ob_v_exists'] is not None:                                     # image_vilmodel.py:100-101
                ex = b['ob_v_exists']
                of = of.masked_fill(ex.logical_not().unsqueeze(-1) if ex.dim() == 2 else ex.logical_not().view(-1, 1, 1), 0)
            B = of.shape[0]
            fb['ob_img_fts'] = torch.cat([of, torch.zeros(B, 1, of.size(2), dtype=of.dtype, device=of.device)], 1)      # STOP token, :103-106
        self.bert._pending_run = run          # the feature path continues inside the same Run (same arena state, dropout sites keep counting)
        try:
            return super().forward(fb, task, compute_loss)
        finally:
            self.bert._pending_run = None
