"""ncu driver: a few launches of the attention kernels at the pano / text / cross shapes of the batch-64 step (dropout on, as in training).
    ncu --set full --clock-control none --import-source on -k regex:attn_ -o gpurun_out/prof_attn python tools/ncu_attn.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hamt_b200  # noqa
from hamt_b200 import ops

drop = ops.Drop(torch.tensor([1], dtype=torch.int64, device="cuda"), 1, 0.1)
for (B, Sq, Sk, masked) in [(960, 36, 36, False), (64, 80, 80, True), (64, 53, 80, True), (64, 16, 80, True)]:
    qkv = torch.randn(B * max(Sq, Sk), 2304, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:B * Sq, :768], qkv[:B * Sk, 768:1536], qkv[:B * Sk, 1536:]
    mask = torch.zeros(B, Sk, device="cuda") if masked else None
    for _ in range(2):
        out, lse = ops.attn_fwd(q, k, v, B, Sq, Sk, 12, mask, drop)
    dout = torch.randn_like(out)
    dqkv = torch.zeros_like(qkv)
    db = torch.zeros(2304, device="cuda")
    for _ in range(2):
        ops.attn_bwd(q, k, v, out, lse, dout, dqkv[:B * Sq, :768], dqkv[:B * Sk, 768:1536], dqkv[:B * Sk, 1536:], B, Sq, Sk, 12, mask, drop, dbias=db)
torch.cuda.synchronize()
print("done")
