"""Launch a handful of in-step GEMM shapes once each (after one warm-up) for an `ncu --set full` capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hamt_b200
from hamt_b200 import ops
torch.manual_seed(0)
def t(*s): return torch.randn(*s, device="cuda").to(torch.bfloat16)
cases = []
# wgrad [768,3072] <- red 5120 ; wgrad [768,768] <- red 5120 ; dgelu dgrad ; gelu fwd ; plain fwd
dy, x = t(5120, 768), t(5120, 3072)
cases.append(("wgrad_768x3072_r5120", lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, out=torch.zeros(768, 3072, device="cuda"), accumulate=True)))
x2 = t(5120, 768)
cases.append(("wgrad_768x768_r5120", lambda: ops.gemm(dy, x2, a_mn=True, b_mn=True, out=torch.zeros(768, 768, device="cuda"), accumulate=True)))
dt, w2, pre = t(3392, 768), t(768, 3072), t(3392, 3072)
cases.append(("dgelu_3392x3072_k768", lambda: ops.gemm(dt, w2, b_mn=True, aux_mode=ops.AUX_MUL_DGELU, aux=pre)))
xa, w1, b1 = t(5120, 768), t(3072, 768), torch.randn(3072, device="cuda")
h = torch.empty(5120, 3072, device="cuda", dtype=torch.bfloat16)
cases.append(("gelu_5120x3072_k768", lambda: ops.gemm(xa, w1, bias=b1, act=ops.ACT_GELU, aux_mode=ops.AUX_STORE_PRE, aux=h)))
cases.append(("fwd_5120x3072_k768", lambda: ops.gemm(xa, w1, bias=b1)))
wq = t(768, 768)
cases.append(("fwd_5120x768_k768", lambda: ops.gemm(xa, wq, bias=b1[:768].contiguous())))
for name, fn in cases:
    fn(); torch.cuda.synchronize()
for name, fn in cases:
    fn(); torch.cuda.synchronize()
print("order:", [c[0] for c in cases])
