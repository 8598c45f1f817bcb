"""End-to-end stage (BASELINE config 3): the cross-modal transformer fed from raw 224 x 224 views through the
ViT-B/16 backbone instead of precomputed features.

Mirrors `NavTHORImagePreTrainedModel` (pretrain_src/model/image_vilmodel.py:22-123): same constructor `(config)`, attribute names
(`vision_backbone`, `embeddings`, `img_embeddings`, `hist_embeddings`, `encoder` -> same state_dict keys), `forward_vision_backbone`
and `forward` signatures.  The backbone runs on the same arena / kernels as the rest of the model; panorama views go through it under
no_grad exactly like the reference ("due to memory issue, we cannot propagate to pano images in the history", :41), history / candidate
views with gradient.  `forward_itm` of the reference's image model is the feature version applied to backbone outputs; it is reached
through `forward_vision_backbone` + `NavPreTrainedModel.forward_itm`.
"""
from __future__ import annotations

import torch

from .vilmodel import NavPreTrainedModel
from .vision_transformer import vit_base_patch16_224


class NavImagePreTrainedModel(NavPreTrainedModel):
    def __init__(self, config, vit_depth: int = 12):
        super().__init__(config)
        # image_vilmodel.py:26-29 (pretrained=True downloads the timm checkpoint in the reference; load one with load_state_dict here)
        self.vision_backbone = vit_base_patch16_224(drop_rate=config.hidden_dropout_prob, attn_drop_rate=config.attention_probs_dropout_prob,
                                                    drop_path_rate=0.0, depth=vit_depth)
        object.__setattr__(self.vision_backbone, "_arena_owner", self)
        # registered first in the reference (:25): same state_dict key order
        vb = self._modules.pop("vision_backbone")
        rest = list(self._modules.items())
        self._modules.clear()
        self._modules["vision_backbone"] = vb
        for k, v in rest:
            self._modules[k] = v

    def forward_vision_backbone(self, images: torch.Tensor, detach: bool = False, _run=None) -> torch.Tensor:
        """image_vilmodel.py:40-59: [N,T,3,H,W] -> [N,T,768]; [N,T,P,3,H,W] (panorama views) -> [N,T,P,768] under no_grad."""
        run = _run or self.begin()
        if images.dim() == 6:
            N, T, P = images.shape[:3]
            with torch.no_grad():
                f = self.vision_backbone.forward_features(images.reshape(N * T * P, *images.shape[3:]), _run=_NoSave(run))
            f = f.view(N, T, P, -1)
        else:
            N, T = images.shape[:2]
            f = self.vision_backbone.forward_features(images.reshape(N * T, *images.shape[2:]), _run=run).view(N, T, -1)
        return f.detach() if detach else f

    def forward(self, txt_ids, txt_masks, hist_images, hist_ang_feats, hist_pano_images, hist_pano_ang_feats, hist_masks,
                ob_images, ob_ang_feats, ob_nav_types, ob_masks, hist_mrc_masks=None, ob_v_exists=None, _run=None):
        """image_vilmodel.py:61-123.  Feature tensors ([B,T,F] / [B,T,P,F] / [B,O,F]) in the image slots take the feature path of the
        parent class unchanged (the pretraining heads call it that way after `MultiStepNavImagePreTraining` ran the backbone)."""
        is_feat = lambda t, nd: t is None or t.dim() == nd      # noqa: E731
        if is_feat(hist_images, 3) and is_feat(hist_pano_images, 4) and is_feat(ob_images, 3):
            return super().forward(txt_ids, txt_masks, hist_images, hist_ang_feats, hist_pano_images, hist_pano_ang_feats, hist_masks,
                                   ob_images, ob_ang_feats, ob_nav_types, ob_masks, _run=_run)
        run = _run or self.begin()
        B = txt_ids.shape[0]
        hist_img_feats = hist_pano_img_feats = None
        if hist_images is not None:
            hist_img_feats = self.forward_vision_backbone(hist_images, _run=run)
            hist_pano_img_feats = self.forward_vision_backbone(hist_pano_images, detach=True, _run=run)
            if hist_mrc_masks is not None:          # (N, T)
                hist_img_feats = hist_img_feats.masked_fill(hist_mrc_masks.unsqueeze(-1), 0)
                hist_pano_img_feats = hist_pano_img_feats.masked_fill(hist_mrc_masks.unsqueeze(-1).unsqueeze(-1), 0)
        ob_img_feats = None
        if ob_images is not None:
            ob_img_feats = self.forward_vision_backbone(ob_images, _run=run)
            if ob_v_exists is not None:
                ob_img_feats = ob_img_feats.masked_fill(ob_v_exists.logical_not().unsqueeze(-1).unsqueeze(-1) if ob_v_exists.dim() == 1
                                                        else ob_v_exists.logical_not().unsqueeze(-1), 0)
            # add the STOP token (:103-106)
            ob_img_feats = torch.cat([ob_img_feats, torch.zeros(B, 1, ob_img_feats.size(2), dtype=ob_img_feats.dtype, device=ob_img_feats.device)], 1)
        return super().forward(txt_ids, txt_masks, hist_img_feats, hist_ang_feats, hist_pano_img_feats, hist_pano_ang_feats, hist_masks,
                               ob_img_feats, ob_ang_feats, ob_nav_types, ob_masks, _run=run)


class _NoSave:
    """View of a Run that saves nothing for backward (panorama views: forward only)."""

    def __init__(self, run):
        object.__setattr__(self, "_run", run)

    def __getattr__(self, name):
        if name == "save":
            return False
        return getattr(self._run, name)

    def __setattr__(self, name, value):
        setattr(self._run, name, value)
