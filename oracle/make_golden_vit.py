"""Golden vectors of the end-to-end ViT stage from the UNMODIFIED reference class (run in the build container):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_vit.py

`VisionTransformer.forward_features` of pretrain_src/model/vision_transformer.py (imported through oracle/ref_shim.py, timm helper
stubs only) at ViT-B/16 geometry, seeded weights (hamt_b200.synth.seeded_vit_state_dict) and seeded synthetic images, eval mode:
outputs in fp32 and under torch.autocast(bfloat16) (the reference's own bf16 error = yardstick of the parity test), plus the fp32
gradient of a fixed scalar loss wrt a few parameters.  Inputs / weights are regenerated from seeds on the test machine, the fixture
only holds outputs.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
import hamt_b200  # noqa: E402,F401
from hamt_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = [("vit_d12_n3", 12, 3, 21, 5), ("vit_d2_n5", 2, 5, 22, 6)]        # name, depth, images, weight seed, image seed
GRAD_KEYS = ["patch_embed.proj.weight", "cls_token", "pos_embed", "blocks.0.attn.qkv.weight", "blocks.0.attn.qkv.bias", "blocks.0.norm1.weight",
             "blocks.1.mlp.fc1.weight", "blocks.1.mlp.fc2.bias", "blocks.1.norm2.bias", "norm.weight"]


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    out = {}
    for name, depth, n, wseed, iseed in CASES:
        m = ref_shim.load_reference_vit(depth=depth).eval()
        m.load_state_dict(synth.seeded_vit_state_dict(m, wseed))
        x = synth.make_images(n, iseed)
        f = m.forward_features(x)
        w = torch.linspace(-1, 1, f.numel()).view_as(f)
        (f * w).sum().backward()
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            fa = m.forward_features(x)
        params = dict(m.named_parameters())
        out[name] = dict(depth=depth, n=n, wseed=wseed, iseed=iseed, feats=f.detach().float().clone(), feats_autocast=fa.float().clone(),
                         grads={k: params[k].grad.float().clone() for k in GRAD_KEYS if k in params})
        print(name, "max |f|", f.abs().max().item(), "autocast gap", (fa.float() - f).abs().max().item())
    # keep the fixture small: gradients of the big tensors as a slice + norm
    for c in out.values():
        g2 = {}
        for k, g in c["grads"].items():
            flat = g.reshape(-1)
            g2[k] = dict(norm=flat.norm().item(), head=flat[:512].clone(), shape=tuple(g.shape))
        c["grads"] = g2
    torch.save(out, os.path.join(GOLD, "vit.pt"))
    print("wrote", os.path.join(GOLD, "vit.pt"), os.path.getsize(os.path.join(GOLD, "vit.pt")), "bytes")


if __name__ == "__main__":
    main()
