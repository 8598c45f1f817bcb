import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hamt_b200
from hamt_b200 import ops
n = 26_000_000
h = torch.randn(n).pin_memory(); d = torch.empty(n, device="cuda")
a = torch.randn(34560, 768, device="cuda").to(torch.bfloat16); w = torch.randn(3072, 768, device="cuda").to(torch.bfloat16)
side = torch.cuda.Stream()
def ev(): return torch.cuda.Event(enable_timing=True)
def compute(k=60):
    for _ in range(k): ops.gemm(a, w)
compute(5); d.copy_(h, non_blocking=True); torch.cuda.synchronize()
s, e = ev(), ev(); s.record()
for _ in range(5): d.copy_(h, non_blocking=True)
e.record(); torch.cuda.synchronize(); t = s.elapsed_time(e) / 5
print(f"H2D alone: {t:.2f} ms = {n*4/t/1e6:.1f} GB/s")
s, e = ev(), ev(); s.record(); compute(); e.record(); torch.cuda.synchronize(); tc = s.elapsed_time(e)
print(f"compute alone: {tc:.2f} ms")
s, e = ev(), ev(); s2, e2 = ev(), ev()
s.record()
with torch.cuda.stream(side):
    s2.record(side)
    for _ in range(5): d.copy_(h, non_blocking=True)
    e2.record(side)
compute(); e.record(); torch.cuda.synchronize()
print(f"concurrent: compute {s.elapsed_time(e):.2f} ms, 5 copies {s2.elapsed_time(e2):.2f} ms ({n*4*5/s2.elapsed_time(e2)/1e6:.1f} GB/s)")
# same with the copy enqueued from the main thread but graph-captured compute
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    compute()
g.replay(); torch.cuda.synchronize()
s, e = ev(), ev(); s2, e2 = ev(), ev()
s.record()
with torch.cuda.stream(side):
    s2.record(side)
    for _ in range(5): d.copy_(h, non_blocking=True)
    e2.record(side)
g.replay(); e.record(); torch.cuda.synchronize()
print(f"concurrent (graph): compute {s.elapsed_time(e):.2f} ms, 5 copies {s2.elapsed_time(e2):.2f} ms")
