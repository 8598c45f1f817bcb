import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hamt_b200
from hamt_b200 import ops
a = torch.randn(34560, 768, device="cuda").to(torch.bfloat16); w = torch.randn(3072, 768, device="cuda").to(torch.bfloat16)
out = torch.empty(34560, 3072, device="cuda", dtype=torch.bfloat16)
for tn in (256, 512):
    ops.gemm(a, w, out=out, tile_n=tn); torch.cuda.synchronize()
for tn in (256, 512):
    ops.gemm(a, w, out=out, tile_n=tn); torch.cuda.synchronize()
