"""Tied MLM decoder GEMMs (804 masked tokens x 30522 classes x 768): forward, dgrad (split-K) and wgrad across tile modes / split counts."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hamt_b200  # noqa
from hamt_b200 import ops
from kbench import timeit
M, N, K = 804, 30522, 768
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
dl = torch.zeros(M, (N + 7) // 8 * 8, device="cuda", dtype=torch.bfloat16)
dl[:, :N] = (torch.randn(M, N, device="cuda") * 0.01).to(torch.bfloat16)
dlv = dl[:, :N]
g = torch.zeros(N, K, device="cuda")
out = torch.empty(M, (N + 7) // 8 * 8, device="cuda")[:, :N]
row = {}
for tn in (0, 128, 256, 512):
    row[f"fwd tile {tn}"] = round(timeit(lambda: ops.gemm(x, w, out=out, out_dtype=torch.float32, tile_n=tn)) * 1e3, 1)
    for sp in (0, 1, 2, 4):
        row[f"wgrad tile {tn} splits {sp}"] = round(timeit(lambda: ops.gemm(dlv, x, a_mn=True, b_mn=True, out=g, accumulate=True, tile_n=tn, splits=sp)) * 1e3, 1)
print(json.dumps(row))
