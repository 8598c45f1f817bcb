"""Launch the FFN GEMMs of the pano encoder (M=34560) once each for an `ncu --set full` capture: plain forward, GELU + pre-activation
store, dGELU dgrad with fused column sums."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hamt_b200
from hamt_b200 import ops
torch.manual_seed(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 34560
def t(*s): return torch.randn(*s, device="cuda").to(torch.bfloat16)
x, w1, b1 = t(M, 768), t(3072, 768), torch.randn(3072, device="cuda")
h = torch.empty(M, 3072, device="cuda", dtype=torch.bfloat16)
dt, w2 = t(M, 768), t(768, 3072)
cs = torch.zeros(3072, device="cuda")
cases = [("plain", lambda: ops.gemm(x, w1, bias=b1)),
         ("gelu", lambda: ops.gemm(x, w1, bias=b1, act=ops.ACT_GELU, aux_mode=ops.AUX_STORE_PRE, aux=h)),
         ("dgelu", lambda: ops.gemm(dt, w2, b_mn=True, aux_mode=ops.AUX_MUL_DGELU, aux=h, colsum=cs)),
         ("gelu+derivative (round 2)", lambda: ops.gemm(x, w1, bias=b1, act=ops.ACT_GELU, aux_mode=ops.AUX_STORE_DGELU, aux=h)),
         ("mul-aux (round 2)", lambda: ops.gemm(dt, w2, b_mn=True, aux_mode=ops.AUX_MUL, aux=h, colsum=cs))]
for _ in range(2):
    for name, fn in cases:
        fn(); torch.cuda.synchronize()
print("order:", [c[0] for c in cases])
