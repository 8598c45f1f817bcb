"""Fused attention forward / backward parity against an fp32 torch restatement of
BertSelfAttention / BertOutAttention (pretrain_src/model/vilmodel.py:96-129, :322-349): scores / sqrt(d) THEN + mask,
softmax, (dropout), P V.  Q/K/V are read in place from a fused [tokens, 3*768] projection buffer."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
HEADS, D = 12, 64


def _ops():
    import hamt_b200  # noqa: F401
    from hamt_b200 import ops
    return ops


@pytest.fixture(params=["tcgen05", "legacy"])
def impl(request):
    """Every test runs twice: with the TMA + tcgen05 packed-tile kernels (default dispatch; shapes outside their envelope fall
    through to the legacy kernels) and with the legacy mma.sync kernels forced (hamt_attn_set_impl(1))."""
    import hamt_b200  # noqa: F401
    from hamt_b200 import _lib
    lib = _lib.load()
    lib.hamt_attn_set_impl(1 if request.param == "legacy" else 2)
    yield request.param
    lib.hamt_attn_set_impl(0)


def _ref(q, k, v, mask, B, Sq, Sk):
    qh = q.float().view(B, Sq, HEADS, D).permute(0, 2, 1, 3)
    kh = k.float().view(B, Sk, HEADS, D).permute(0, 2, 1, 3)
    vh = v.float().view(B, Sk, HEADS, D).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(D)
    if mask is not None:
        s = s + mask[:, None, None, :]
    p = torch.softmax(s, -1)
    return (p @ vh).permute(0, 2, 1, 3).reshape(B * Sq, HEADS * D), p


@pytest.mark.parametrize("B,Sq,Sk,masked", [(3, 36, 36, False), (4, 80, 80, True), (4, 80, 53, True), (4, 53, 80, True), (2, 16, 5, True),
                                             (2, 1, 17, True), (2, 128, 128, True), (5, 17, 100, True),
                                             # long sequences (RxR instructions, BASELINE config 4: L = 300, 58 vision tokens): recompute backward
                                             (2, 300, 300, True), (2, 300, 58, True), (2, 58, 300, True), (1, 129, 40, False), (1, 250, 512, True),
                                             # tile-packing regimes of the tcgen05 kernels: 4 / 3 (+ shared remainder warp) / 2 / 1 problems per
                                             # 128-row tile, packed key axis cut to fit TMEM, several query tiles per problem
                                             (6, 16, 16, True), (7, 16, 80, True), (3, 40, 40, True), (3, 33, 36, False), (2, 41, 41, True),
                                             (2, 64, 24, True), (2, 78, 78, True), (2, 36, 80, True), (1, 32, 128, True), (5, 53, 53, True)])
def test_attention_fwd_bwd(B, Sq, Sk, masked, impl):
    ops = _ops()
    g = torch.Generator().manual_seed(B * 1000 + Sq * 10 + Sk)
    self_attn = Sq == Sk
    if self_attn:
        qkv = torch.randn(B * Sq, 3 * HEADS * D, generator=g).to(torch.bfloat16).cuda()
        q, k, v = qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:]
    else:
        qb = torch.randn(B * Sq, 3 * HEADS * D, generator=g).to(torch.bfloat16).cuda()
        kb = torch.randn(B * Sk, 3 * HEADS * D, generator=g).to(torch.bfloat16).cuda()
        q, k, v = qb[:, :768], kb[:, 768:1536], kb[:, 1536:]
    mask = None
    if masked:
        lens = torch.randint(1, Sk + 1, (B,), generator=g)
        lens[0] = Sk
        mask = ((torch.arange(Sk)[None] >= lens[:, None]).float() * -10000.0).cuda()
    out, lse = ops.attn_fwd(q, k, v, B, Sq, Sk, HEADS, mask)
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    ref, p = _ref(qr, kr, vr, mask, B, Sq, Sk)
    err = (out.float() - ref).abs().max().item()
    assert err < 3e-2, f"fwd err {err}"
    dout = torch.randn(B * Sq, HEADS * D, generator=g).to(torch.bfloat16).cuda()
    ref.backward(dout.float())
    if self_attn:
        dqkv = torch.zeros_like(qkv)
        dq, dk, dv = dqkv[:, :768], dqkv[:, 768:1536], dqkv[:, 1536:]
    else:
        dqb, dkb = torch.zeros_like(qb), torch.zeros_like(kb)
        dq, dk, dv = dqb[:, :768], dkb[:, 768:1536], dkb[:, 1536:]
    dbias = torch.full((3 * HEADS * D,), 0.5, dtype=torch.float32, device="cuda")
    ops.attn_bwd(q, k, v, out, lse, dout, dq, dk, dv, B, Sq, Sk, HEADS, mask, dbias=dbias)
    for got, want, name in ((dq, qr.grad, "dq"), (dk, kr.grad, "dk"), (dv, vr.grad, "dv")):
        e = (got.float() - want).abs().max().item()
        assert e < 4e-2 * max(1.0, want.abs().max().item()), f"{name} err {e}"
    # fused bias gradients: fp32 column sums of the unrounded dq / dk / dv accumulators, ACCUMULATED onto the buffer.  Against the fp32
    # autograd sums with the tolerance of the gradients themselves, and against the sums of the stored bf16 values up to their
    # rounding noise (2^-9 relative per element, random sign: ~ sqrt(rows) * 2^-9 * |element|).
    want_b = torch.cat([qr.grad.sum(0), kr.grad.sum(0), vr.grad.sum(0)]) + 0.5
    assert (dbias - want_b).abs().max().item() <= 4e-2 * max(1.0, want_b.abs().max().item())
    stored_b = torch.cat([dq.float().sum(0), dk.float().sum(0), dv.float().sum(0)]) + 0.5
    rows = B * max(Sq, Sk)
    elem = max(dq.float().abs().max().item(), dk.float().abs().max().item(), dv.float().abs().max().item())
    assert (dbias - stored_b).abs().max().item() <= 4 * (rows ** 0.5) * 2.0 ** -9 * elem + 1e-3
    # and without the pointers the gradients are bit-identical
    if self_attn:
        g2 = torch.zeros_like(qkv)
        dq2, dk2, dv2 = g2[:, :768], g2[:, 768:1536], g2[:, 1536:]
    else:
        gq2, gk2 = torch.zeros_like(qb), torch.zeros_like(kb)
        dq2, dk2, dv2 = gq2[:, :768], gk2[:, 768:1536], gk2[:, 1536:]
    ops.attn_bwd(q, k, v, out, lse, dout, dq2, dk2, dv2, B, Sq, Sk, HEADS, mask)
    assert torch.equal(dq2, dq) and torch.equal(dk2, dk) and torch.equal(dv2, dv)


def test_attention_fully_masked_rows_match_reference(impl):
    """-10000 masks (not -inf): a fully masked key row still yields the reference's softmax over the raw scores."""
    ops = _ops()
    B, S = 2, 20
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(B * S, 2304, generator=g).to(torch.bfloat16).cuda()
    mask = torch.full((B, S), -10000.0).cuda()
    out, _ = ops.attn_fwd(qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:], B, S, S, HEADS, mask)
    ref, _ = _ref(qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:], mask, B, S, S)
    assert (out.float() - ref).abs().max().item() < 3e-2


def test_attention_dropout_statistics_and_backward_consistency(impl):
    """With V = identity-like probes the output exposes dropout(P); check keep-rate and that the backward
    (which regenerates the mask) matches autograd through the explicitly recovered mask."""
    ops = _ops()
    B, S, p = 2, 64, 0.1
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(B * S, 2304, generator=g).to(torch.bfloat16).cuda() * 0.5
    eye = torch.eye(S, D).to(torch.bfloat16).cuda()           # S == D: V_h = I  ->  out_h = dropout(P_h)
    qkv[:, 1536:] = eye.repeat(B, HEADS)
    q, k, v = qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:]
    seed = torch.tensor([99], dtype=torch.int64, device="cuda")
    drop = ops.Drop(seed, 3, p)
    out, lse = ops.attn_fwd(q, k, v, B, S, S, HEADS, None, drop)
    _, P = _ref(q, k, v, None, B, S, S)                       # [B,H,S,S]
    Pd = out.float().view(B, S, HEADS, D).permute(0, 2, 1, 3)  # dropout(P) (bf16-rounded)
    keep = Pd > 0
    big = P > 1e-3
    frac = keep[big].float().mean().item()
    assert abs(frac - (1 - p)) < 1e-2, frac
    assert (Pd[keep & big] / P[keep & big] - 1 / (1 - p)).abs().max().item() < 5e-2
    # backward with the recovered mask
    mult = torch.where(keep | ~big, torch.full_like(P, 1 / (1 - p)), torch.zeros_like(P))
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    qh = qr.view(B, S, HEADS, D).permute(0, 2, 1, 3); kh = kr.view(B, S, HEADS, D).permute(0, 2, 1, 3); vh = vr.view(B, S, HEADS, D).permute(0, 2, 1, 3)
    pr = torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, -1) * mult
    ref = (pr @ vh).permute(0, 2, 1, 3).reshape(B * S, 768)
    dout = torch.randn(B * S, 768, generator=g).to(torch.bfloat16).cuda()
    ref.backward(dout.float())
    dqkv = torch.zeros_like(qkv)
    ops.attn_bwd(q, k, v, out, lse, dout, dqkv[:, :768], dqkv[:, 768:1536], dqkv[:, 1536:], B, S, S, HEADS, None, drop)
    for got, want in ((dqkv[:, :768], qr.grad), (dqkv[:, 768:1536], kr.grad), (dqkv[:, 1536:], vr.grad)):
        assert (got.float() - want).abs().max().item() < 6e-2 * max(1.0, want.abs().max().item())


def test_attention_implementations_agree_bitwise_on_the_dropout_mask():
    """The tcgen05 and the legacy kernels draw the SAME dropout mask (one counter hash over (sequence, head, query, key)), so a forward
    of one and a backward of the other stay consistent; outputs agree to bf16 rounding."""
    import hamt_b200  # noqa: F401
    from hamt_b200 import _lib
    ops = _ops()
    lib = _lib.load()
    B, S, p = 3, 36, 0.3
    g = torch.Generator().manual_seed(11)
    qkv = torch.randn(B * S, 2304, generator=g).to(torch.bfloat16).cuda()
    q, k, v = qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:]
    drop = ops.Drop(torch.tensor([4242], dtype=torch.int64, device="cuda"), 7, p)
    outs = []
    for mode in (2, 1):
        lib.hamt_attn_set_impl(mode)
        try:
            outs.append(ops.attn_fwd(q, k, v, B, S, S, HEADS, None, drop))
        finally:
            lib.hamt_attn_set_impl(0)
    (o0, l0), (o1, l1) = outs
    assert (o0.float() - o1.float()).abs().max().item() < 2e-2
    assert (l0 - l1).abs().max().item() < 1e-3
