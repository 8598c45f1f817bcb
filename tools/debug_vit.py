"""Debug aid: per-parameter gradient comparison of the CUDA ViT against autograd through the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hamt_b200  # noqa
from hamt_b200 import synth
from hamt_b200.vision_transformer import VisionTransformer
from oracle import vit_oracle as V

for depth in (0, 1, 2):
    m = VisionTransformer(depth=depth, num_classes=0)
    sd = synth.seeded_vit_state_dict(m, 5)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = synth.make_images(3, 2)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    fo = V.forward_features(sdo, x, depth=depth)
    w = torch.linspace(-1, 1, fo.numel()).view_as(fo)
    (fo * w).sum().backward()
    f = m.forward_features(x.cuda())
    (f * w.cuda()).sum().backward()
    torch.cuda.synchronize()
    print(f"depth {depth}: fwd err {(f.detach().cpu() - fo.detach()).abs().max().item():.3e}")
    for k, p in m.named_parameters():
        g, go = p.grad.detach().float().cpu(), sdo[k].grad
        rel = (g - go).norm().item() / max(go.norm().item(), 1e-12)
        print(f"   {k:34s} rel {rel:8.4f}  |ours| {g.norm().item():.4e} |oracle| {go.norm().item():.4e}")
