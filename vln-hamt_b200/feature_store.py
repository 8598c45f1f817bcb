"""Device-resident view-feature store and on-device batch assembly.

The reference keeps the precomputed ViT features of every panorama (36 views x (768 features + 1000 class logits), fp32) in an HDF5 file,
and builds each sample on the CPU: `MultiStepNavData.get_history_feature` / `get_ob_pano_view` (pretrain_src/data/r2r_data.py:187-208,
:264-329) stack per-step rows, the collate functions pad them (`pad_tensors`, data/common.py:5-20; r2r_tasks.py:343-379) and the
PrefetchLoader copies ~106 MB per batch to the GPU (data/loader.py:78-125).

Here the whole table lives in HBM once, as bf16 [V * 36, D] (10.5 k panoramas x 36 x 768 x 2 B = 0.6 GB of 180 GB), and a batch is described
by INDICES: which panorama each history step / observation comes from and which view the agent faced.  One gather kernel per tensor
(`hamt_gather_rows_pad_bf16`: coalesced 16-byte loads of the 768-d rows, zero rows for padding) writes the model inputs directly in the
layout and dtype the kernels consume (bf16), so the per-step host -> device traffic drops from ~103 MB to a few KB of indices and the
fp32 -> bf16 cast pass over the features disappears.

Semantics restated from the reference:
  * history step t of a sample: `hist_img_fts[t] = fts[vp_t][view_t]`, `hist_pano_img_fts[t] = fts[vp_t][:, :D]`,
    `hist_pano_ang_fts[t] = angle_features[view_t]` (r2r_data.py:283-296); steps beyond the sample's length are zero (pad_tensors);
  * observation (pano view mode): 36 views of the current panorama + an all-zero STOP row, angle features relative to the current view
    index + a zero row (r2r_data.py:187-189, :316-329); a sample whose observation was "killed" by the augmentation
    (r2r_tasks.py:322-327) is all zeros -> index -1;
  * `hist_img_probs[t] = softmax(fts[vp_t][view_t, D:])` (r2r_data.py:298-299, :307-308) when the class-logit table is loaded.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch

from . import ops

N_VIEWS = 36


def point_angle_features(angle_feat_size: int = 4) -> torch.Tensor:
    """[36 base views, 36 views, A] fp32: sin/cos of heading / elevation of every discretised view relative to the heading of the
    base view -- `get_all_point_angle_feature` (r2r_data.py:19-35, angle_feature :14-17)."""
    out = torch.empty(N_VIEWS, N_VIEWS, angle_feat_size, dtype=torch.float32)
    for base in range(N_VIEWS):
        base_heading = (base % 12) * math.radians(30)
        heading, elevation = 0.0, 0.0
        for ix in range(N_VIEWS):
            if ix == 0:
                heading, elevation = 0.0, math.radians(-30)
            elif ix % 12 == 0:
                heading = 0.0
                elevation += math.radians(30)
            else:
                heading += math.radians(30)
            h = heading - base_heading
            row = [math.sin(h), math.cos(h), math.sin(elevation), math.cos(elevation)] * (angle_feat_size // 4)
            out[base, ix] = torch.tensor(row, dtype=torch.float64).to(torch.float32)
    return out


class FeatureStore:
    """keys: one '<scan>_<viewpoint>' string per panorama (the HDF5 keys, r2r_data.py:311); feats: [V, 36, >= D] (any float dtype, host
    or device): columns [:D] are the image features, columns [D:] (optional) the class logits used by MRC."""

    def __init__(self, keys: Sequence[str], feats: torch.Tensor, device, image_feat_size: int = 768, angle_feat_size: int = 4,
                 keep_logits: bool = True):
        if feats.dim() != 3 or feats.shape[0] != len(keys) or feats.shape[1] != N_VIEWS or feats.shape[2] < image_feat_size:
            raise ValueError("FeatureStore: feats must be [len(keys), 36, >= image_feat_size]")
        if image_feat_size % 8:
            raise ValueError("FeatureStore: image_feat_size must be a multiple of 8 (16-byte rows)")
        self.device = torch.device(device)
        self.D, self.A = image_feat_size, angle_feat_size
        self.index: Dict[str, int] = {k: i for i, k in enumerate(keys)}
        if len(self.index) != len(keys):
            raise ValueError("FeatureStore: duplicate keys")
        self.n_pano = len(keys)
        self.table = feats[:, :, :image_feat_size].to(self.device).to(torch.bfloat16).reshape(self.n_pano * N_VIEWS, image_feat_size).contiguous()
        self.logits = None
        if keep_logits and feats.shape[2] > image_feat_size:
            self.logits = feats[:, :, image_feat_size:].to(self.device, torch.float32).reshape(self.n_pano * N_VIEWS, -1).contiguous()
        self.angle_table = point_angle_features(angle_feat_size).to(self.device)                   # [36, 36, A]
        self._views = torch.arange(N_VIEWS, device=self.device)

    # -- host side: names -> indices ---------------------------------------------------------------------------------
    def lookup(self, scan: str, viewpoint: str) -> int:
        return self.index["%s_%s" % (scan, viewpoint)]

    def _check(self, pano: torch.Tensor, view: Optional[torch.Tensor]):
        if pano.dtype != torch.int64 or (view is not None and view.dtype != torch.int64):
            raise ValueError("FeatureStore: indices must be int64")
        if pano.device.type == "cpu":       # indices are normally built on the host: validate there, never on the device
            if pano.numel() and (int(pano.max()) >= self.n_pano or int(pano.min()) < -1):
                raise IndexError("FeatureStore: panorama index out of range")
            if view is not None and view.numel() and (int(view.max()) >= N_VIEWS or int(view.min()) < -1):
                raise IndexError("FeatureStore: view index out of range")

    # -- device side: indices -> model inputs ------------------------------------------------------------------------
    def assemble_history(self, hist_pano: torch.Tensor, hist_view: torch.Tensor, with_pano: bool = True, with_probs: bool = False,
                         mrc_mask: Optional[torch.Tensor] = None) -> Dict:
        """hist_pano / hist_view: int64 [B, T]; -1 = padding (step >= the sample's history length).  Returns the reference's batch
        entries `hist_img_fts` [B,T,D], `hist_pano_img_fts` [B,T,36,D], `hist_pano_ang_fts` [B,T,36,A] (features in bf16) and, with
        with_probs, `hist_img_probs` [B,T,P] fp32.  T == 0 -> every entry is None (r2r_tasks.py:360-366).
        mrc_mask (bool [B,T], optional): the MRC augmentation (`_mask_img_feat` / `_mask_pano_img_feat`, r2r_tasks.py:183-190) zeroes
        ONLY the image features of the masked steps (`hist_img_fts`, `hist_pano_img_fts`) and keeps their angle features and their
        class probabilities (the prediction targets)."""
        self._check(hist_pano, hist_view)
        B, T = hist_pano.shape
        keys = ["hist_img_fts"] + (["hist_pano_img_fts", "hist_pano_ang_fts"] if with_pano else []) + (["hist_img_probs"] if with_probs else [])
        if T == 0:
            return {k: None for k in keys}
        pano = hist_pano.to(self.device, non_blocking=True)
        view = hist_view.to(self.device, non_blocking=True)
        valid = (pano >= 0) & (view >= 0)
        base = pano * N_VIEWS
        row = torch.where(valid, base + view, torch.full_like(base, -1))
        shown = valid
        if mrc_mask is not None:
            if mrc_mask.shape != hist_pano.shape:
                raise ValueError("FeatureStore: mrc_mask must be [B, T]")
            shown = valid & ~mrc_mask.to(self.device, non_blocking=True).bool()
        img_row = torch.where(shown, row, torch.full_like(row, -1))
        out = {"hist_img_fts": ops.gather_rows_pad(self.table, img_row.reshape(-1).contiguous()).view(B, T, self.D)}
        if with_pano:
            rows36 = torch.where(shown[..., None], base[..., None] + self._views, torch.full((1,), -1, dtype=torch.int64, device=self.device))
            out["hist_pano_img_fts"] = ops.gather_rows_pad(self.table, rows36.reshape(-1).contiguous()).view(B, T, N_VIEWS, self.D)
            ang = self.angle_table[view.clamp(min=0)]                                    # [B,T,36,A]
            out["hist_pano_ang_fts"] = ang * valid[..., None, None].to(ang.dtype)
        if with_probs:
            if self.logits is None:
                raise RuntimeError("FeatureStore: built without the class-logit columns")
            lg = self.logits[row.clamp(min=0)]                                           # [B,T,P]
            out["hist_img_probs"] = torch.softmax(lg, -1) * valid[..., None].to(lg.dtype)
        return out

    def assemble_observation(self, ob_pano: torch.Tensor, ob_view: torch.Tensor) -> Dict:
        """ob_pano / ob_view: int64 [B] (current panorama, current view index); ob_pano = -1 zeroes the image features of the sample,
        ob_view = -1 its angle features (the random_kill augmentation).  Returns `ob_img_fts` [B,37,D] bf16 and `ob_ang_fts` [B,37,A]
        fp32 with the all-zero STOP row last (r2r_data.py:187-189)."""
        self._check(ob_pano, ob_view)
        B = ob_pano.shape[0]
        pano = ob_pano.to(self.device, non_blocking=True)
        view = ob_view.to(self.device, non_blocking=True)
        rows = torch.full((B, N_VIEWS + 1), -1, dtype=torch.int64, device=self.device)
        rows[:, :N_VIEWS] = torch.where((pano >= 0)[:, None], pano[:, None] * N_VIEWS + self._views, rows[:, :N_VIEWS])
        img = ops.gather_rows_pad(self.table, rows.reshape(-1).contiguous()).view(B, N_VIEWS + 1, self.D)
        ang = torch.zeros(B, N_VIEWS + 1, self.A, dtype=torch.float32, device=self.device)
        ang[:, :N_VIEWS] = self.angle_table[view.clamp(min=0)] * (view >= 0)[:, None, None].to(torch.float32)
        return {"ob_img_fts": img, "ob_ang_fts": ang}
