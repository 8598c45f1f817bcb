"""ViT-B/16 vision backbone of the end-to-end stage (BASELINE config 3).

Mirrors the reference's vendored timm model -- pretrain_src/model/vision_transformer.py:226-361 (`VisionTransformer`), :181-198
(`Block`), :154-178 (`Attention`), :132-151 (`Mlp`), :201-223 (`PatchEmbed`) -- as PARAMETER HOLDERS with the same attribute names, so
`state_dict()` keys, shapes and order are the reference's (`patch_embed.proj.weight`, `cls_token`, `pos_embed`,
`blocks.{i}.norm1 / attn.qkv / attn.proj / norm2 / mlp.fc1 / mlp.fc2`, `norm`, `head`) and a timm / reference checkpoint loads
unchanged.  The arithmetic is `functional.VitFn`: patchify -> tcgen05 GEMM (the Conv2d with kernel = stride = 16) -> cls / pos
assembly -> 12 pre-LN blocks on the same GEMM / attention / LayerNorm kernels as the feature-based path (S = 197) -> final norm ->
class token.  There is no eager fallback.

Not ported (not on the HAMT path): distillation token, representation layer, stochastic depth (the reference passes
drop_path_rate = 0, image_vilmodel.py:26-29), the npz / hub checkpoint loaders.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as Fn
from .arena import ParamArena

VIT_LN_EPS = 1e-6       # vision_transformer.py:265


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - parameter holder
        raise RuntimeError("hamt_b200: this module only holds parameters; call VisionTransformer.forward_features")


class Mlp(_Holder):
    def __init__(self, in_features, hidden_features, drop=0.0):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, in_features)
        self.drop = nn.Dropout(drop)


class Attention(_Holder):
    def __init__(self, dim, num_heads, qkv_bias=True, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)


class Block(_Holder):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=True, drop=0.0, attn_drop=0.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=VIT_LN_EPS)
        self.attn = Attention(dim, num_heads, qkv_bias, attn_drop, drop)
        self.drop_path = nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=VIT_LN_EPS)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), drop)


class PatchEmbed(_Holder):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.patch_grid = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.patch_grid[0] * self.patch_grid[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.Identity()


class VisionTransformer(nn.Module):
    """`vit_base_patch16_224` geometry by default (vision_transformer.py:507-513)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 qkv_bias=True, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0):
        super().__init__()
        if drop_path_rate:
            raise ValueError("hamt_b200: stochastic depth is not part of the HAMT path (the reference passes drop_path_rate = 0)")
        if embed_dim // num_heads != 64 or embed_dim not in (512, 768, 1024) or not qkv_bias:
            raise ValueError("hamt_b200: the ViT kernels are built for head_dim 64, embed_dim 512/768/1024 and a biased qkv projection")
        self.num_classes, self.num_features, self.embed_dim, self.num_tokens, self.num_heads = num_classes, embed_dim, embed_dim, 1, num_heads
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = None
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio, qkv_bias, drop_rate, attn_drop_rate) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=VIT_LN_EPS)
        self.pre_logits = nn.Identity()
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()      # not used by the HAMT path
        nn.init.trunc_normal_(self.pos_embed, std=0.02, a=-0.04, b=0.04)
        nn.init.trunc_normal_(self.cls_token, std=0.02, a=-0.04, b=0.04)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02, a=-0.04, b=0.04)
                nn.init.zeros_(m.bias)
        self._arena = None
        self._arena_owner = None          # set by an enclosing model (image_vilmodel): its arena holds these parameters too

    # ---- arena / run plumbing (stand-alone use; inside NavImagePreTrainedModel the parent's Run is passed in) -------------------------
    def arena(self) -> ParamArena:
        if self._arena_owner is not None:
            return self._arena_owner.arena()
        if self._arena is None:
            self._arena = ParamArena(self)
        return self._arena

    def begin(self) -> Fn.Run:
        arena = self.arena()
        training = self.training and torch.is_grad_enabled()
        arena.step_begin(training)
        seed = None
        if self.training:
            arena.next_seed()
            seed = arena.run_seed()
        return Fn.Run(arena, self.training, self.num_heads, VIT_LN_EPS, seed=seed)

    def forward_features(self, x: torch.Tensor, _run=None) -> torch.Tensor:
        """images fp32 [N, 3, H, W] -> class-token features fp32 [N, embed_dim] (vision_transformer.py:335-348)."""
        if not x.is_cuda:
            raise RuntimeError("hamt_b200: the compute path needs CUDA tensors (no CPU fallback)")
        if tuple(x.shape[-2:]) != self.patch_embed.img_size:
            raise ValueError(f"Input image size ({x.shape[-2]}*{x.shape[-1]}) doesn't match model ({self.patch_embed.img_size[0]}*{self.patch_embed.img_size[1]}).")
        run = _run or self.begin()
        return Fn.VitFn.apply(run.arena.anchor, x.float().contiguous(), run, self)

    def forward(self, x):
        f = self.forward_features(x)
        if isinstance(self.head, nn.Identity):
            return f
        raise RuntimeError("hamt_b200: the classification head is not part of the HAMT path; use forward_features")


def vit_base_patch16_224(pretrained=False, **kwargs) -> VisionTransformer:
    """vision_transformer.py:507-513.  `pretrained=True` in the reference downloads the timm ImageNet checkpoint; there is no
    network here, load a checkpoint with load_state_dict instead."""
    kw = dict(patch_size=16, embed_dim=768, depth=12, num_heads=12)
    kw.update(kwargs)
    return VisionTransformer(**kw)
