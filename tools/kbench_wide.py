"""Round-2 starter: time the FFN GEMMs of the pano / text encoders with the default (8-warp) and the EXPERIMENTAL wide (16-warp)
epilogue (hamt_gemm_set_wide_epilogue).  Run the opt-in parity test first:
    HAMT_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gemm_gpu.py -m gpu -q -k wide
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hamt_b200  # noqa
from hamt_b200 import _lib, ops
from kbench import timeit

lib = _lib.load()
for M in (34560, 5120):
    def t(*s): return torch.randn(*s, device="cuda").to(torch.bfloat16)
    x, w1, b1 = t(M, 768), t(3072, 768), torch.randn(3072, device="cuda")
    h = torch.empty(M, 3072, device="cuda", dtype=torch.bfloat16)
    dt, w2, w3 = t(M, 768), t(768, 3072), t(3072, 768)
    dh, dx = t(M, 3072), t(M, 768)
    cs = torch.zeros(3072, device="cuda")
    cases = {"store N=3072": lambda: ops.gemm(x, w1, bias=b1), "store N=768 K=3072": lambda: ops.gemm(dh, w2, bias=b1[:768].contiguous()),
             "gelu+pre": lambda: ops.gemm(x, w1, bias=b1, act=ops.ACT_GELU, aux_mode=ops.AUX_STORE_PRE, aux=h),
             "dgelu+colsum": lambda: ops.gemm(dt, w2, b_mn=True, aux_mode=ops.AUX_MUL_DGELU, aux=h, colsum=cs),
             "gelu+derivative (r2)": lambda: ops.gemm(x, w1, bias=b1, act=ops.ACT_GELU, aux_mode=ops.AUX_STORE_DGELU, aux=h),
             "mul-aux+colsum (r2)": lambda: ops.gemm(dt, w2, b_mn=True, aux_mode=ops.AUX_MUL, aux=h, colsum=cs),
             "accum K=3072": lambda: ops.gemm(dh, w3, b_mn=True, out=dx, accumulate=True)}
    row = {"M": M}
    for name, fn in cases.items():
        for wide in (0, 1):
            lib.hamt_gemm_set_wide_epilogue(wide)
            row[f"{name} {'wide' if wide else 'base'} us"] = round(timeit(fn) * 1e3, 1)
    lib.hamt_gemm_set_wide_epilogue(0)
    print(json.dumps(row), flush=True)

