"""Kernel micro-benchmarks (CUDA events, warm-up, L2-sized rotation of inputs) -- development aid, not the graded bench."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hamt_b200  # noqa
from hamt_b200 import ops


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def timeit(fn, n=10, warm=2):          # noqa: F811 -- graph-timed: no CPU launch overhead between the n launches
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (3 * n)


def main():
    res = []
    shapes = [] if "--no-gemm" in sys.argv else [(34560, 768, 768), (34560, 2304, 768), (34560, 3072, 768), (34560, 768, 3072), (5120, 768, 768), (5120, 2304, 768),
                      (5120, 3072, 768), (5120, 768, 3072), (8512, 2304, 768), (3392, 3072, 768)]
    for (M, N, K) in shapes:
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        row = {"M": M, "N": N, "K": K}
        for tn in (128, 256, 512):
            t = timeit(lambda: ops.gemm(a, w, bias=bias, tile_n=tn))
            row[f"fwd_tn{tn}_tflops"] = round(2 * M * N * K / t / 1e9, 1)
        bb = bias.to(torch.bfloat16)
        t = timeit(lambda: torch.nn.functional.linear(a, w, bb))
        row["torch_tflops"] = round(2 * M * N * K / t / 1e9, 1)
        dy = torch.randn(M, N, device="cuda").to(torch.bfloat16)
        dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
        for tn in (0, 512):
            t = timeit(lambda: ops.gemm(dy, w, b_mn=True, out=dx, tile_n=tn))
            row[f"dgrad_tn{tn}_tflops"] = round(2 * M * N * K / t / 1e9, 1)
        g = torch.zeros(N, K, device="cuda")
        for tn in (0, 512):
            t = timeit(lambda: ops.gemm(dy, a, a_mn=True, b_mn=True, out=g, accumulate=True, tile_n=tn))
            row[f"wgrad_tn{tn}_tflops"] = round(2 * M * N * K / t / 1e9, 1)
        res.append(row)
        print(json.dumps(row), flush=True)
    # attention (pano / text / cross / vision shapes of the batch-64 step, plus the 16-token history-only shapes of MLM / MRC / ITM)
    from hamt_b200 import _lib
    shapes_attn = [(960, 36, 36), (64, 80, 80), (64, 80, 53), (64, 53, 80), (64, 53, 53), (64, 16, 80), (64, 80, 16), (64, 16, 16), (2560, 36, 36)]
    if "--attn-long" in sys.argv:       # RxR instructions (L = 300) and the ViT sequence (S = 197): legacy kernels (key axis > 128)
        shapes_attn = [(64, 300, 300), (64, 300, 58), (64, 58, 300), (82, 197, 197)]
    for (B, Sq, Sk) in shapes_attn:
      for legacy in ((1, 2) if "--attn-ab" in sys.argv else (0,)):       # 1 = legacy mma.sync kernels, 2 = tcgen05 forced wherever it fits, 0 = shipped dispatch
        _lib.load().hamt_attn_set_impl(legacy)
        qkv = torch.randn(B * max(Sq, Sk), 2304, device="cuda").to(torch.bfloat16)
        q, k, v = qkv[:B * Sq, :768], qkv[:B * Sk, 768:1536], qkv[:B * Sk, 1536:]
        if "--attn-drop" in sys.argv:
            drop = ops.Drop(torch.tensor([1], dtype=torch.int64, device="cuda"), 1, 0.1)
            t = timeit(lambda: ops.attn_fwd(q, k, v, B, Sq, Sk, 12, None, drop))
            print(json.dumps({"attn_dropout": [B, Sq, Sk], "impl": {0: "auto", 1: "legacy", 2: "tcgen05"}[legacy], "fwd_us": round(t * 1e3, 1)}), flush=True)
        t = timeit(lambda: ops.attn_fwd(q, k, v, B, Sq, Sk, 12, None))
        out, lse = ops.attn_fwd(q, k, v, B, Sq, Sk, 12, None)
        dout = torch.randn_like(out)
        dqkv = torch.zeros_like(qkv)
        t2 = timeit(lambda: ops.attn_bwd(q, k, v, out, lse, dout, dqkv[:B * Sq, :768], dqkv[:B * Sk, 768:1536], dqkv[:B * Sk, 1536:], B, Sq, Sk, 12, None))
        db = torch.zeros(2304, device="cuda")
        t3 = timeit(lambda: ops.attn_bwd(q, k, v, out, lse, dout, dqkv[:B * Sq, :768], dqkv[:B * Sk, 768:1536], dqkv[:B * Sk, 1536:], B, Sq, Sk, 12, None, dbias=db))
        print(json.dumps({"attn": [B, Sq, Sk], "impl": {0: "auto", 1: "legacy", 2: "tcgen05"}[legacy], "fwd_us": round(t * 1e3, 1), "bwd_us": round(t2 * 1e3, 1),
                          "bwd_dbias_us": round(t3 * 1e3, 1), "fwd_GBs": round((B * (Sq + 2 * Sk) * 768 * 2 + B * Sq * 768 * 2) / t / 1e6, 1),
                          "bwd_GBs": round((B * (2 * Sq + 2 * Sk) * 768 * 2 + B * (Sq + 2 * Sk) * 768 * 2) / t3 / 1e6, 1)}), flush=True)
    _lib.load().hamt_attn_set_impl(0)
    if "--attn-only" in sys.argv:
        return
    # layernorm
    for M in (34560, 5120):
        x = torch.randn(M, 768, device="cuda").to(torch.bfloat16)
        r = torch.randn(M, 768, device="cuda").to(torch.bfloat16)
        gm, bt = torch.ones(768, device="cuda"), torch.zeros(768, device="cuda")
        t = timeit(lambda: ops.ln_fwd(x, r, gm, bt, 1e-12))
        print(json.dumps({"ln_fwd_M": M, "us": round(t * 1e3, 1), "GBs": round(M * 768 * 2 * 4 / t / 1e6, 1)}), flush=True)
        y, z, mean, rstd = ops.ln_fwd(x.clone(), r, gm, bt, 1e-12)
        dy = torch.randn(M, 768, device="cuda").to(torch.bfloat16)
        dg, db, dbias = (torch.zeros(768, device="cuda") for _ in range(3))
        t = timeit(lambda: ops.ln_bwd(dy, z, mean, rstd, gm, dg, db, dbias))
        print(json.dumps({"ln_bwd_M": M, "us": round(t * 1e3, 1), "GBs": round(M * 768 * 2 * 4 / t / 1e6, 1)}), flush=True)
        q = torch.randn(M, 2304, device="cuda").to(torch.bfloat16)
        acc = torch.zeros(2304, device="cuda")
        t = timeit(lambda: ops.colsum(q, acc))
        print(json.dumps({"colsum_M": M, "N": 2304, "us": round(t * 1e3, 1), "GBs": round(M * 2304 * 2 / t / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
