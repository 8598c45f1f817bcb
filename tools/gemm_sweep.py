"""GEMM time against K at fixed (M, N) for every tile mode, next to cuBLAS: the intercept is the per-launch fixed cost (pipeline fill,
last-tile epilogue, grid drain, launch gap), the slope the steady-state tensor-pipe rate.  Development aid (round 2)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hamt_b200  # noqa
from hamt_b200 import ops
from kbench import timeit


def main():
    Ks = (64, 256, 768, 1536, 3072)
    for (M, N) in [(5120, 768), (5120, 2304), (5120, 3072), (34560, 768), (1024, 768), (2560, 3072)]:
        row = {"M": M, "N": N}
        for name, kw in (("tn128", dict(tile_n=128)), ("tn256", dict(tile_n=256)), ("pair", dict(tile_n=512)), ("auto", dict()), ("cublas", None)):
            ts = []
            for K in Ks:
                a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
                w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
                out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
                if kw is None:
                    t = timeit(lambda: torch.mm(a, w.t(), out=out))
                else:
                    t = timeit(lambda: ops.gemm(a, w, out=out, **kw))
                ts.append(t * 1e3)
            slope, icpt = np.polyfit(np.array(Ks[2:], dtype=float), np.array(ts[2:]), 1)
            row[name] = {"us": [round(x, 1) for x in ts], "intercept_us": round(float(icpt), 1), "slope_TF": round(2 * M * N / slope / 1e6, 0)}
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
