"""tcgen05 GEMM parity (C-ABI hamt_gemm_bf16 through ops.gemm) against fp32 torch matmul on the same
bf16-rounded operands.  Covers the three operand-major combinations the hot path uses (forward
K-major x K-major, dgrad K-major x MN-major, wgrad MN-major x MN-major), both tile widths, ragged
edges (M, N, K not multiples of the tile), every fused epilogue, and split-K accumulation."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dgelu_ref(x):
    return 0.5 * (1.0 + torch.erf(x * 0.7071067811865476)) + x * 0.3989422804014327 * torch.exp(-0.5 * x * x)


def _ops():
    import hamt_b200  # noqa: F401
    from hamt_b200 import ops
    return ops


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _tol(K):
    # fp32 accumulation of K bf16 products; output rounded to bf16 -> 2^-8 relative + accumulation noise
    return 2e-2, 1e-2


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 768, 768), (5120, 768, 768), (1024, 2304, 768), (640, 3072, 768),
                                    (333, 1000, 768), (256, 768, 3072), (200, 768, 1000), (77, 37 * 8, 1536)])
@pytest.mark.parametrize("tile_n", [128, 256, 512])
def test_gemm_forward_bias(M, N, K, tile_n):
    ops = _ops()
    a, b = _rand((M, K), 1), _rand((N, K), 2, 0.05)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).cuda()
    out = ops.gemm(a, b, bias=bias, tile_n=tile_n)
    ref = a.float() @ b.float().t() + bias
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item() + 1e-3, f"max err {err}"


def test_gemm_fp32_out_exact_small():
    """fp32 output, integer-valued operands: the tensor-core result must be exact."""
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    a = torch.randint(-4, 5, (256, 192), generator=g).to(torch.bfloat16).cuda()
    b = torch.randint(-4, 5, (384, 192), generator=g).to(torch.bfloat16).cuda()
    out = ops.gemm(a, b, out_dtype=torch.float32)
    assert torch.equal(out, a.float() @ b.float().t())


@pytest.mark.parametrize("act", ["gelu", "relu"])
def test_gemm_activation_and_preact(act):
    ops = _ops()
    M, N, K = 512, 3072, 768
    a, b = _rand((M, K), 4), _rand((N, K), 5, 0.05)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(6)).cuda() * 0.1
    aux = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    out = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU if act == "gelu" else ops.ACT_RELU, aux_mode=ops.AUX_STORE_PRE, aux=aux)
    pre = a.float() @ b.float().t() + bias
    ref = torch.nn.functional.gelu(pre) if act == "gelu" else torch.relu(pre)
    assert (aux.float() - pre).abs().max().item() <= 2e-2 * pre.abs().max().item()
    assert (out.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item() + 1e-3


@pytest.mark.parametrize("M,N,K", [(512, 768, 768), (300, 768, 3072), (5120, 3072, 768), (129, 768, 1000)])
def test_gemm_dgrad_mn_major_b(M, N, K):
    """dx[M,N] = dy[M,K] @ W[K,N]: B operand is MN-major (W stored [K_red, N_out])."""
    ops = _ops()
    dy, w = _rand((M, K), 7), _rand((K, N), 8, 0.05)
    out = ops.gemm(dy, w, b_mn=True)
    ref = dy.float() @ w.float()
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item() + 1e-3, f"max err {err}"


def test_gemm_dgrad_dgelu_epilogue():
    ops = _ops()
    M, N, K = 384, 3072, 768
    dy, w = _rand((M, K), 9), _rand((K, N), 10, 0.05)
    pre = _rand((M, N), 11)
    out = ops.gemm(dy, w, b_mn=True, aux_mode=ops.AUX_MUL_DGELU, aux=pre)
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    ref = (dy.float() @ w.float()) * x.grad
    assert (out.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item() + 1e-3


@pytest.mark.parametrize("M,tile_n", [(384, 0), (5120, 0), (777, 128), (1024, 256), (1024, 512)])
def test_gemm_gelu_derivative_store_and_multiply_epilogues(M, tile_n):
    """Round-2 FFN pair: the forward epilogue stores gelu'(pre) next to gelu(pre) (aux_mode 4), the backward epilogue multiplies the
    dgrad by that saved derivative (aux_mode 5) with the fused bias column sums -- against torch's erf GELU and its autograd."""
    ops = _ops()
    N, K = 3072, 768
    a, b = _rand((M, K), 21), _rand((N, K), 22, 0.05)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(23)).cuda() * 0.1
    der = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    out = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU, aux_mode=ops.AUX_STORE_DGELU, aux=der, tile_n=tile_n)
    pre = (a.float() @ b.float().t() + bias).requires_grad_(True)
    ref = torch.nn.functional.gelu(pre)
    ref.sum().backward()
    assert (out.float() - ref.detach()).abs().max().item() <= 2e-2 * ref.abs().max().item() + 1e-3
    # the derivative is O(1) (range [-0.13, 1.13]); bf16 storage + the bf16-level uncertainty of pre itself
    assert (der.float() - pre.grad).abs().max().item() <= 1.5e-2
    # identical activation output as the pre-activation-saving epilogue
    aux = torch.empty_like(der)
    out1 = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU, aux_mode=ops.AUX_STORE_PRE, aux=aux, tile_n=tile_n)
    assert torch.equal(out, out1)
    dy, w = _rand((M, K), 24), _rand((K, N), 25, 0.05)
    cs = torch.zeros(N, dtype=torch.float32, device="cuda")
    dh = ops.gemm(dy, w, b_mn=True, aux_mode=ops.AUX_MUL, aux=der, colsum=cs, tile_n=tile_n)
    want = (dy.float() @ w.float()) * der.float()
    assert (dh.float() - want).abs().max().item() <= 1e-2 * want.abs().max().item() + 1e-3
    assert (cs - dh.float().sum(0)).abs().max().item() <= 2e-3 * dh.float().abs().sum(0).max().item() + 1e-3


@pytest.mark.parametrize("Mred,N,K", [(5120, 768, 768), (3392, 3072, 768), (1000, 768, 3072), (160, 2304, 768), (34560, 768, 768)])
def test_gemm_wgrad_mn_major_both_splitk(Mred, N, K):
    """dW[N,K] += dy[Mred,N]^T @ x[Mred,K]: both operands MN-major, fp32 output, split-K atomics."""
    ops = _ops()
    dy, x = _rand((Mred, N), 12, 0.1), _rand((Mred, K), 13)
    out = torch.full((N, K), 0.5, dtype=torch.float32, device="cuda")
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=out, accumulate=True)
    ref = dy.float().t() @ x.float() + 0.5
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3, f"max err {err}"


def test_gemm_wgrad_forced_splits_match():
    ops = _ops()
    dy, x = _rand((4096, 768), 14, 0.1), _rand((4096, 768), 15)
    outs = []
    for s in (1, 4, 16):
        o = torch.zeros((768, 768), dtype=torch.float32, device="cuda")
        ops.gemm(dy, x, a_mn=True, b_mn=True, out=o, accumulate=True, splits=s)
        outs.append(o)
    assert (outs[0] - outs[1]).abs().max().item() < 1e-2
    assert (outs[0] - outs[2]).abs().max().item() < 1e-2


@pytest.mark.parametrize("M,N,K", [(512, 768, 768), (5120, 3072, 768), (300, 768, 3072), (129, 768, 1000)])
def test_gemm_pair_dgrad_and_epilogues(M, N, K):
    """CTA-pair kernel (tile_n=512, tcgen05 cta_group::2): MN-major B, dGELU epilogue with prefetched aux, bf16 accumulate."""
    ops = _ops()
    dy, w = _rand((M, K), 7), _rand((K, N), 8, 0.05)
    out = ops.gemm(dy, w, b_mn=True, tile_n=512)
    ref = dy.float() @ w.float()
    assert (out.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item() + 1e-3
    pre = _rand((M, N), 11)
    out2 = ops.gemm(dy, w, b_mn=True, aux_mode=ops.AUX_MUL_DGELU, aux=pre, tile_n=512)
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    ref2 = ref * x.grad
    assert (out2.float() - ref2).abs().max().item() <= 2e-2 * ref2.abs().max().item() + 1e-3
    base = _rand((M, N), 12)
    acc = base.clone()
    ops.gemm(dy, w, b_mn=True, out=acc, accumulate=True, tile_n=512)
    ref3 = ref + base.float()
    assert (acc.float() - ref3).abs().max().item() <= 3e-2 * ref3.abs().max().item() + 1e-3


@pytest.mark.parametrize("Mred,N,K", [(5120, 768, 768), (3392, 3072, 768), (34560, 768, 768), (160, 2304, 768)])
def test_gemm_pair_wgrad(Mred, N, K):
    ops = _ops()
    dy, x = _rand((Mred, N), 12, 0.1), _rand((Mred, K), 13)
    out = torch.full((N, K), 0.5, dtype=torch.float32, device="cuda")
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=out, accumulate=True, tile_n=512)
    ref = dy.float().t() @ x.float() + 0.5
    assert (out - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3


def test_gemm_strided_views():
    """Operands / outputs that are column slices of wider buffers (fused QKV layout)."""
    ops = _ops()
    big = _rand((512, 2304), 16)
    w = _rand((768, 768), 17, 0.05)
    a = big[:, 768:1536]
    outbuf = torch.zeros((512, 2304), dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, w, out=outbuf[:, 1536:])
    ref = a.float() @ w.float().t()
    assert (outbuf[:, 1536:].float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    assert outbuf[:, :1536].abs().max().item() == 0


def test_gemm_rejects_bad_pitch():
    ops = _ops()
    a = _rand((64, 100), 18)[:, :99]          # pitch 100 elements = 200 B: not a multiple of 16 B
    b = _rand((64, 99), 19)
    with pytest.raises((RuntimeError, ValueError)):
        ops.gemm(a, b)


def test_gemm_gelu_epilogue_accuracy():
    """The single-ex2 Gaussian-cdf polynomial of the GELU epilogue: |gelu error| well below bf16 resolution.  fp32 accumulate of
    identity-like operands isolates the activation: out = gelu(bias) exactly up to the final bf16 rounding."""
    ops = _ops()
    N = 4096
    xs = torch.linspace(-9.0, 9.0, N).cuda()
    a = torch.zeros((128, 64), dtype=torch.bfloat16, device="cuda")
    b = torch.zeros((N, 64), dtype=torch.bfloat16, device="cuda")
    out = ops.gemm(a, b, bias=xs, act=ops.ACT_GELU, out_dtype=torch.bfloat16)
    ref = torch.nn.functional.gelu(xs.double()).float()
    got = out[0].float()
    # within half a bf16 ulp of the exact value (ulp <= |v| * 2^-7) + 1e-5 absolute for the polynomial
    assert ((got - ref).abs() <= ref.abs() * 2.0 ** -8 + 1e-5).all(), f"max |err| {(got - ref).abs().max().item()}"


@pytest.mark.parametrize("M,N,K,tile_n", [(384, 3072, 768, 0), (5120, 3072, 768, 512), (333, 1000, 768, 256), (129, 776, 192, 128)])
def test_gemm_fused_colsum(M, N, K, tile_n):
    """colsum: the dgrad epilogue accumulates the column sums of the stored bf16 output (bias gradient of the producing Linear,
    vilmodel.py:168-171 backward) -- must equal a separate pass over the output."""
    ops = _ops()
    dy, w = _rand((M, K), 21), _rand((K, N), 22, 0.05)
    pre = _rand((M, N), 23)
    cs = torch.full((N,), 0.25, dtype=torch.float32, device="cuda")
    out = ops.gemm(dy, w, b_mn=True, aux_mode=ops.AUX_MUL_DGELU, aux=pre, colsum=cs, tile_n=tile_n)
    ref = out.float().sum(0) + 0.25
    assert (cs - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item()) + 1e-3
    out2 = ops.gemm(dy, w, b_mn=True, aux_mode=ops.AUX_MUL_DGELU, aux=pre, tile_n=tile_n)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("M,tile_n", [(1024, 256), (2048, 512)])
def test_gemm_wide_dgelu_epilogue_matches_8warp_epilogue(M, tile_n):
    """The dGELU dgrad runs a 16-warp epilogue on aligned 256-wide tiles (default); hamt_gemm_set_wide_epilogue(0) selects the 8-warp
    epilogue: same math -> bit-identical outputs, column sums equal up to fp32 summation order."""
    import hamt_b200  # noqa: F401
    from hamt_b200 import _lib
    ops = _ops()
    lib = _lib.load()
    N = 768
    dy, w2 = _rand((M, 512), 33), _rand((512, N), 34, 0.05)
    pre = _rand((M, N), 35)

    def run():
        cs = torch.zeros(N, device="cuda")
        o3 = ops.gemm(dy, w2, b_mn=True, aux_mode=ops.AUX_MUL_DGELU, aux=pre, colsum=cs, tile_n=tile_n)
        torch.cuda.synchronize()
        return o3, cs

    wide = run()
    lib.hamt_gemm_set_wide_epilogue(0)
    try:
        base = run()
    finally:
        lib.hamt_gemm_set_wide_epilogue(1)
    assert torch.equal(base[0], wide[0])
    assert (base[1] - wide[1]).abs().max().item() <= 1e-3 * max(1.0, base[1].abs().max().item())
    ref = (dy.float() @ w2.float()) * _dgelu_ref(pre.float())
    assert (wide[0].float() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())


def test_gemm_mlm_decoder_shapes_row_tiles_fastest_and_splitk_dgrad():
    """Tied MLM decoder (pretrain_cmt.py:96-99, vilmodel.py:280-284): logits [n, 30522] = x [n, 768] W^T with a 47 MB weight (row tiles
    iterate fastest so that it is read once) and its dgrad dx [n, 768] = dlogits [n, 30522] W (split-K into fp32)."""
    ops = _ops()
    M, N, K = 804, 30522, 768
    x, w = _rand((M, K), 31), _rand((N, K), 32, 0.05)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(33)).cuda() * 0.1
    out = torch.empty((M, (N + 7) // 8 * 8), dtype=torch.float32, device="cuda")[:, :N]
    ops.gemm(x, w, bias=bias, out=out, out_dtype=torch.float32)
    ref = x.float() @ w.float().t() + bias
    assert (out - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-4
    dl = torch.zeros((M, (N + 7) // 8 * 8), dtype=torch.bfloat16, device="cuda")
    dl[:, :N] = _rand((M, N), 34, 0.01)
    dx = ops.gemm(dl[:, :N], w, b_mn=True, out_dtype=torch.float32, accumulate=True)
    refd = dl[:, :N].float() @ w.float()
    assert (dx - refd).abs().max().item() <= 2e-3 * refd.abs().max().item() + 1e-4
