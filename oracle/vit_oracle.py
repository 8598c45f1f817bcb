"""CPU oracle for the end-to-end ViT-B/16 stage: a functional restatement of the reference's vision backbone.

TEST INFRASTRUCTURE ONLY -- imported by tests/, never by the product path.

Restates  pretrain_src/model/vision_transformer.py  (a vendored timm copy):
  PatchEmbed.forward          :217-223   Conv2d(3, 768, k = s = 16) -> flatten(2).transpose(1, 2)
  VisionTransformer.forward_features :335-348   [cls ; patches] + pos_embed -> pos_drop -> 12 Blocks -> norm -> x[:, 0]
  Block.forward               :195-198   x = x + attn(norm1(x));  x = x + mlp(norm2(x))        (pre-LN, LayerNorm eps 1e-6 :265)
  Attention.forward           :166-178   fused qkv Linear, softmax(q k^T * head_dim^-0.5) v, proj
  Mlp.forward                 :145-151   fc1 -> nn.GELU (erf) -> drop -> fc2 -> drop
and the caller  pretrain_src/model/image_vilmodel.py:40-59  (forward_vision_backbone: [N,T(,P),3,224,224] -> [N,T(,P),768], panorama
images under no_grad).

Pinned: against the UNMODIFIED reference class imported through oracle/ref_shim.py (timm is not installed here; the shim provides the
five helper names vision_transformer.py imports from it -- none of them contributes forward arithmetic with drop_path = 0) in
tests/test_oracle.py and through the golden vectors oracle/make_golden_vit.py writes to tests/golden/.

State-dict keys are timm's: patch_embed.proj.{weight,bias}, cls_token, pos_embed, blocks.{i}.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,
mlp.fc2}.{weight,bias}, norm.{weight,bias}.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .hamt_oracle import FP32, Regime, gelu_erf

Tensor = torch.Tensor
VIT_LN_EPS = 1e-6        # vision_transformer.py:265  partial(nn.LayerNorm, eps=1e-6)


def _linear(sd: Dict[str, Tensor], name: str, x: Tensor, rg: Regime, quant_out: bool = True) -> Tensor:
    y = F.linear(rg.q(x), rg.q(sd[name + ".weight"]), None) + sd[name + ".bias"]
    return rg.q(y) if quant_out else y


def _ln(sd: Dict[str, Tensor], name: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], VIT_LN_EPS)


def patch_embed(sd: Dict[str, Tensor], images: Tensor, rg: Regime, prefix: str = "") -> Tensor:
    """vision_transformer.py:217-223.  Conv2d with kernel = stride = 16 is a GEMM over non-overlapping patches whose 768 inputs are
    ordered (channel, row, column) -- the flattening of the conv weight [768, 3, 16, 16]."""
    w = sd[prefix + "patch_embed.proj.weight"]
    E, C, ph, pw = w.shape
    N, _, H, W = images.shape
    gh, gw = H // ph, W // pw
    cols = images.view(N, C, gh, ph, gw, pw).permute(0, 2, 4, 1, 3, 5).reshape(N, gh * gw, C * ph * pw)
    y = F.linear(rg.q(cols), rg.q(w.reshape(E, -1)), None) + sd[prefix + "patch_embed.proj.bias"]
    return rg.q(y)


def block(sd: Dict[str, Tensor], pre: str, x: Tensor, heads: int, rg: Regime) -> Tensor:
    """vision_transformer.py:195-198 with Attention :166-178 and Mlp :145-151 (eval mode: every Dropout / DropPath is the identity).
    The residual stream x stays fp32 in both regimes (torch.autocast keeps the adds in fp32; the CUDA path carries an fp32 stream)."""
    N, S, C = x.shape
    d = C // heads
    y = rg.q(_ln(sd, pre + "norm1", x))
    qkv = _linear(sd, pre + "attn.qkv", y, rg).view(N, S, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = torch.softmax((q @ k.transpose(-2, -1)) * (d ** -0.5), dim=-1)
    ctx = rg.q((rg.q(att) @ v).transpose(1, 2).reshape(N, S, C))
    x = x + _linear(sd, pre + "attn.proj", ctx, rg)
    y = rg.q(_ln(sd, pre + "norm2", x))
    h = _linear(sd, pre + "mlp.fc1", y, rg, quant_out=False)
    a = rg.q(gelu_erf(h))
    x = x + _linear(sd, pre + "mlp.fc2", a, rg)
    return x


def forward_features(sd: Dict[str, Tensor], images: Tensor, heads: int = 12, rg: Regime = FP32, prefix: str = "",
                     depth: Optional[int] = None) -> Tensor:
    """vision_transformer.py:335-348 (dist_token is None, pre_logits = Identity for vit_base_patch16_224): images [N,3,H,W] -> [N,768]."""
    x = patch_embed(sd, images, rg, prefix)
    N = x.shape[0]
    cls = sd[prefix + "cls_token"].expand(N, -1, -1)
    x = torch.cat([cls, x], dim=1) + sd[prefix + "pos_embed"]
    if depth is None:
        depth = 1 + max(int(k[len(prefix) + 7:].split(".")[0]) for k in sd if k.startswith(prefix + "blocks."))
    for i in range(depth):
        x = block(sd, f"{prefix}blocks.{i}.", x, heads, rg)
    x = _ln(sd, prefix + "norm", x)
    return x[:, 0]


def forward_vision_backbone(sd: Dict[str, Tensor], images: Tensor, heads: int = 12, rg: Regime = FP32, prefix: str = "",
                            detach: bool = False) -> Tensor:
    """image_vilmodel.py:40-59: [N,T,3,H,W] -> [N,T,768] (with grad), [N,T,P,3,H,W] -> [N,T,P,768] (panorama views, no_grad)."""
    if images.dim() == 6:
        N, T, P = images.shape[:3]
        with torch.no_grad():
            f = forward_features(sd, images.reshape(N * T * P, *images.shape[3:]), heads, rg, prefix)
        f = f.view(N, T, P, -1)
    else:
        N, T = images.shape[:2]
        f = forward_features(sd, images.reshape(N * T, *images.shape[2:]), heads, rg, prefix).view(N, T, -1)
    return f.detach() if detach else f
