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
             "accum K=3072": lambda: ops.gemm(dh, w3, b_mn=True, out=dx, accumulate=True)}
    row = {"M": M}
    for name, fn in cases.items():
        for wide in (0, 1):
            lib.hamt_gemm_set_wide_epilogue(wide)
            row[f"{name} {'wide' if wide else 'base'} us"] = round(timeit(fn) * 1e3, 1)
    lib.hamt_gemm_set_wide_epilogue(0)
    print(json.dumps(row), flush=True)

# LayerNorm backward: default kernel vs the experimental variants (hamt_ln_set_variant)
for M in (34560, 8512, 5120, 3392):
    x = torch.randn(M, 768, device="cuda").to(torch.bfloat16)
    r = torch.randn(M, 768, device="cuda").to(torch.bfloat16)
    gm, bt = torch.ones(768, device="cuda"), torch.zeros(768, device="cuda")
    y, z, mean, rstd = ops.ln_fwd(x.clone(), r, gm, bt, 1e-12)
    dy, dri = torch.randn(M, 768, device="cuda").to(torch.bfloat16), torch.randn(M, 768, device="cuda").to(torch.bfloat16)
    dg, db, dbias = (torch.zeros(768, device="cuda") for _ in range(3))
    row = {"ln_bwd_M": M}
    for v in (0, 1, 2, 3):
        lib.hamt_ln_set_variant(v)
        row[f"variant{v}_us"] = round(timeit(lambda: ops.ln_bwd(dy, z, mean, rstd, gm, dg, db, dbias, dres_in=dri)) * 1e3, 1)
    lib.hamt_ln_set_variant(0)
    print(json.dumps(row), flush=True)
