"""Seeded synthetic batches with the layout of the reference's collate functions.

Follows SURVEY.md section 8(d); layout sources: pretrain_src/data/r2r_tasks.py (collate fns),
r2r_data.py:14-17 (angle feature = sin/cos of heading/elevation), :205-208 (STOP = all-zero row
appended last, nav_type 2), r2r_tasks.py:60 (real-word id range), :498-506 (sprel targets).
All tensors are generated on the CPU with a torch.Generator so every machine draws identical
values; ``device`` only says where they are moved afterwards.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch


def _angles(gen, shape):
    h = (torch.rand(shape, generator=gen) * 2 - 1) * math.pi
    e = (torch.randint(0, 3, shape, generator=gen).float() - 1.0) * (math.pi / 6)
    return torch.stack([torch.sin(h), torch.cos(h), torch.sin(e), torch.cos(e)], -1)


def sprel_target_table() -> torch.Tensor:
    """[36,36,2] table built like SprelDataset.__init__ (r2r_tasks.py:498-506) with angles
    standardised to (-pi, pi]."""
    t = torch.zeros(36, 36, 2)
    for i in range(36):
        ah, ae = (i % 12) * math.radians(30), (i // 12 - 1) * math.radians(30)
        for j in range(36):
            ch, ce = (j % 12) * math.radians(30), (j // 12 - 1) * math.radians(30)
            for c, v in enumerate((ch - ah, ce - ae)):
                v = (v + math.pi) % (2 * math.pi) - math.pi
                t[i, j, c] = v
    return t


def make_batch(task: str, batch_size: int = 64, txt_len: int = 80, hist_len: int = 15, n_pano: int = 36,
               n_ob: int = 37, feat: int = 768, prob_size: int = 1000, seed: int = 0, ragged: bool = False,
               device: str = "cpu", vocab_hi: int = 29611, dtype=torch.float32) -> Dict[str, Optional[torch.Tensor]]:
    """One collated batch for ``task`` in {mlm,sap,sar,sprel,mrc,itm}.  hist_len == 0 reproduces the
    'all samples at step 0' case where hist_*_fts are None (r2r_tasks.py:360-366)."""
    g = torch.Generator().manual_seed(seed)
    B, L, T, P, O = batch_size, txt_len, hist_len, n_pano, n_ob
    b: Dict[str, Optional[torch.Tensor]] = {}
    ids = torch.randint(1996, vocab_hi, (B, L), generator=g)
    ids[:, 0] = 101
    if ragged:
        lens = torch.randint(min(20, L), L + 1, (B,), generator=g)
        lens[0] = L
    else:
        lens = torch.full((B,), L)
    txt_masks = torch.arange(L)[None] < lens[:, None]
    ids = ids * txt_masks
    b["txt_ids"], b["txt_masks"] = ids, txt_masks

    if T > 0:
        b["hist_img_fts"] = torch.randn(B, T, feat, generator=g)
        b["hist_ang_fts"] = _angles(g, (B, T))
        b["hist_pano_img_fts"] = torch.randn(B, T, P, feat, generator=g)
        b["hist_pano_ang_fts"] = _angles(g, (B, T, P))
        if ragged:
            hl = torch.randint(1, T + 2, (B,), generator=g)
            hl[0] = T + 1
        else:
            hl = torch.full((B,), T + 1)
        b["hist_masks"] = torch.arange(T + 1)[None] < hl[:, None]        # includes the CLS slot
    else:
        for k in ("hist_img_fts", "hist_ang_fts", "hist_pano_img_fts", "hist_pano_ang_fts"):
            b[k] = None
        b["hist_masks"] = torch.ones(B, 1, dtype=torch.bool)

    if task in ("sap", "sar", "sprel"):
        ob = torch.randn(B, O, feat, generator=g)
        ob[:, -1] = 0                                                     # STOP row
        b["ob_img_fts"] = ob
        b["ob_ang_fts"] = _angles(g, (B, O))
        nav = torch.zeros(B, O, dtype=torch.long)
        labels = torch.zeros(B, dtype=torch.long)
        for i in range(B):
            cand = torch.randperm(O - 1, generator=g)[:4]
            nav[i, cand] = 1
            nav[i, -1] = 2
            pool = torch.cat([cand, torch.tensor([O - 1])])
            labels[i] = pool[torch.randint(0, len(pool), (1,), generator=g)]
        b["ob_nav_types"] = nav
        b["ob_masks"] = torch.ones(B, O, dtype=torch.bool)
        if task == "sap":
            b["ob_action_viewindex"] = labels
        if task == "sar":
            b["ob_action_angles"] = (torch.rand(B, 2, generator=g) * 2 - 1) * math.pi
            b["ob_progress"] = torch.rand(B, generator=g)
        if task == "sprel":
            anchor = torch.randint(0, 36, (B,), generator=g)
            b["sp_anchor_idxs"] = anchor
            b["sp_targets"] = sprel_target_table()[anchor]
    if task == "mlm":
        lab = torch.full((B, L), -1, dtype=torch.long)
        pick = (torch.rand(B, L, generator=g) < 0.15) & txt_masks
        pick[:, 0] = False
        for i in range(B):
            if not pick[i].any():
                pick[i, 1] = True
        lab[pick] = ids[pick]
        ids2 = ids.clone()
        ids2[pick] = 103
        b["txt_ids"], b["txt_labels"] = ids2, lab
    if task == "mrc":
        assert T > 0
        m = (torch.rand(B, T, generator=g) < 0.15) & b["hist_masks"][:, 1:]
        for i in range(B):
            if not m[i].any():
                m[i, 0] = True
        b["hist_mrc_masks"] = m
        b["hist_img_probs"] = torch.softmax(torch.randn(B, T, prob_size, generator=g), -1)
    out = {}
    for k, v in b.items():
        if v is None:
            out[k] = None
        elif v.is_floating_point():
            out[k] = v.to(device=device, dtype=dtype)
        else:
            out[k] = v.to(device=device)
    return out


def seeded_state_dict(model: torch.nn.Module, seed: int = 0, std: float = 0.02, perturb_ln: bool = True):
    """Deterministic BERT-style initialisation of every tensor of ``model.state_dict()`` in key
    order (weights ~ N(0, std); Linear biases and LayerNorm affine get a small random perturbation
    when ``perturb_ln`` so that bias / gamma / beta code paths are actually exercised by parity tests;
    a plain BERT init has them at exactly 0 / 1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    seen = {}
    for k, v in model.state_dict().items():
        if v.data_ptr() in seen:                       # tied weights share one draw
            sd[k] = sd[seen[v.data_ptr()]]
            continue
        seen[v.data_ptr()] = k
        shape = tuple(v.shape)
        is_ln = ("LayerNorm" in k) or ("layer_norm" in k) or (".net.2." in k)
        if k.endswith("weight") and is_ln:
            t = torch.ones(shape) + (0.1 * torch.randn(shape, generator=g) if perturb_ln else 0)
        elif k.endswith("bias") or is_ln:
            t = 0.02 * torch.randn(shape, generator=g) if perturb_ln else torch.zeros(shape)
        else:
            t = std * torch.randn(shape, generator=g)
        sd[k] = t.to(v.dtype)
    return sd


def seeded_vit_state_dict(model: torch.nn.Module, seed: int = 0, std: float = 0.02):
    """Deterministic initialisation of a ViT backbone's state_dict in key order (timm key names: `norm1` / `norm2` / `norm` are the
    LayerNorms): weights ~ N(0, std), LayerNorm gamma = 1 + 0.1 N(0,1), every bias / beta / cls / pos ~ 0.02 N(0,1), so that no code
    path is exercised at a trivial value.  Used for the reference class, the oracle and the CUDA module alike."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        shape = tuple(v.shape)
        leaf = k.split(".")[-2] if "." in k else k
        is_ln = leaf in ("norm", "norm1", "norm2")
        if is_ln and k.endswith("weight"):
            t = torch.ones(shape) + 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("bias") or k in ("cls_token", "pos_embed"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            t = std * torch.randn(shape, generator=g)
        sd[k] = t.to(v.dtype)
    return sd


def make_images(n: int, seed: int = 0, size: int = 224) -> torch.Tensor:
    """Synthetic normalised RGB views, fp32 [n, 3, size, size]: smooth low-frequency content + noise (ImageNet-normalised pixels are O(1))."""
    g = torch.Generator().manual_seed(seed)
    low = torch.nn.functional.interpolate(torch.randn(n, 3, 14, 14, generator=g), size=(size, size), mode="bilinear", align_corners=False)
    return (low + 0.3 * torch.randn(n, 3, size, size, generator=g)).contiguous()


def make_image_batch(task: str, batch_size: int = 1, txt_len: int = 60, hist_len: int = 5, n_pano: int = 36, n_ob: int = 37, seed: int = 0,
                     size: int = 224, device: str = "cpu") -> Dict[str, Optional[torch.Tensor]]:
    """End-to-end stage batch (image collate of the reference's stage 2: `hist_images` [B,T,3,S,S], `hist_pano_images`
    [B,T,P,3,S,S], `ob_images` [B,O-1,3,S,S], `ob_v_exists`; image_pretrain.py:47-90 keys) on top of make_batch's text / angle / label
    tensors.  Images are normalised fp32 pixels; they are drawn directly on `device` (seeded) because a batch is gigabytes."""
    b = make_batch(task, batch_size=batch_size, txt_len=txt_len, hist_len=hist_len, n_pano=n_pano, n_ob=n_ob, feat=8, seed=seed)
    for k in ("hist_img_fts", "hist_pano_img_fts", "ob_img_fts"):
        b.pop(k, None)
    out = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in b.items()}
    g = torch.Generator(device=device).manual_seed(seed + 77)
    B, T, P, O = batch_size, hist_len, n_pano, n_ob
    if T > 0:
        out["hist_images"] = torch.randn(B, T, 3, size, size, generator=g, device=device)
        out["hist_pano_images"] = torch.randn(B, T, P, 3, size, size, generator=g, device=device)
    if task in ("sap", "sar", "sprel"):
        out["ob_images"] = torch.randn(B, O - 1, 3, size, size, generator=g, device=device)
        out["ob_v_exists"] = torch.ones(B, O - 1, dtype=torch.bool, device=device)
    return out
