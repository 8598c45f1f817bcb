"""Parity of the row-wise CUDA kernels (LayerNorm fwd/bwd with fused dropout+residual, embedders,
mean-pool, column sums, head pieces) against fp32 torch autograd on the same inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

EPS = 1e-12


def _ops():
    import hamt_b200  # noqa: F401
    from hamt_b200 import ops
    return ops


def G(seed):
    return torch.Generator().manual_seed(seed)


def bf(x):
    return x.to(torch.bfloat16).cuda()


@pytest.mark.parametrize("M", [1, 7, 300, 5120])
@pytest.mark.parametrize("with_res", [False, True])
def test_ln_fwd_bwd(M, with_res):
    ops = _ops()
    H = 768
    x = bf(torch.randn(M, H, generator=G(1)))
    res = bf(torch.randn(M, H, generator=G(2))) if with_res else None
    gamma = (1 + 0.1 * torch.randn(H, generator=G(3))).cuda()
    beta = (0.1 * torch.randn(H, generator=G(4))).cuda()
    dy = bf(torch.randn(M, H, generator=G(5)))
    y, z, mean, rstd = ops.ln_fwd(x.clone(), res, gamma, beta, EPS, inplace_z=True)
    xr = x.float().requires_grad_(True)
    rr = res.float().requires_grad_(True) if with_res else None
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    zz = xr + rr if with_res else xr
    yr = F.layer_norm(zz, (H,), gr, br, EPS)
    assert (y.float() - yr).abs().max().item() < 3e-2
    assert (z.float() - zz).abs().max().item() < 3e-2
    yr.backward(dy.float())
    dgamma, dbeta, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
    dx, dres = ops.ln_bwd(dy, z, mean, rstd, gamma, dgamma, dbeta, dbias, want_dres=with_res)
    tol = 4e-2 * max(1.0, xr.grad.abs().max().item())
    assert (dx.float() - xr.grad).abs().max().item() < tol
    if with_res:
        assert (dres.float() - rr.grad).abs().max().item() < tol
    scale = max(1.0, math.sqrt(M))
    assert (dgamma - gr.grad).abs().max().item() < 5e-2 * scale
    assert (dbeta - br.grad).abs().max().item() < 5e-2 * scale
    assert (dbias - xr.grad.sum(0)).abs().max().item() < 5e-2 * scale


def test_ln_dropout_mask_consistency():
    """The mask regenerated in backward equals the forward one; keep fraction ~ 1-p; new seed -> new mask."""
    ops = _ops()
    M, H, p = 2048, 768, 0.1
    seed = torch.tensor([1234], dtype=torch.int64, device="cuda")
    drop = ops.Drop(seed, site=7, p=p)
    ones = torch.ones(M, H, dtype=torch.bfloat16, device="cuda")
    gamma, beta = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
    _, z, mean, rstd = ops.ln_fwd(ones.clone(), None, gamma, beta, EPS, drop)
    keep = z.float() > 0
    frac = keep.float().mean().item()
    assert abs(frac - (1 - p)) < 5e-3
    assert torch.allclose(z.float()[keep], torch.full((1,), 1 / (1 - p), device="cuda").to(torch.bfloat16).float().expand(int(keep.sum())))
    # backward: dx must be zero exactly where the forward dropped
    dy = bf(torch.randn(M, H, generator=G(1)))
    dx, _ = ops.ln_bwd(dy, z, mean, rstd, gamma, None, None, None, want_dres=False, drop=drop)
    assert (dx.float()[~keep] == 0).all()
    assert (dx.float()[keep] != 0).float().mean().item() > 0.99
    _, z2, _, _ = ops.ln_fwd(ones.clone(), None, gamma, beta, EPS, drop)
    assert torch.equal(z, z2)
    seed.add_(1)
    _, z3, _, _ = ops.ln_fwd(ones.clone(), None, gamma, beta, EPS, drop)
    assert not torch.equal(z, z3)
    _, z4, _, _ = ops.ln_fwd(ones.clone(), None, gamma, beta, EPS, ops.Drop(seed, site=8, p=p))
    assert not torch.equal(z3, z4)


def test_embed_text_fwd_bwd():
    ops = _ops()
    B, L, H, V = 5, 23, 768, 1000
    ids = torch.randint(0, V, (B, L), generator=G(1)).cuda()
    word = (0.02 * torch.randn(V, H, generator=G(2))).cuda()
    pos = (0.02 * torch.randn(64, H, generator=G(3))).cuda()
    typ = (0.02 * torch.randn(2, H, generator=G(4))).cuda()
    gamma = (1 + 0.1 * torch.randn(H, generator=G(5))).cuda()
    beta = (0.1 * torch.randn(H, generator=G(6))).cuda()
    out = ops.embed_text_fwd(ids, word, pos, typ[0], gamma, beta, EPS)
    ws = [t.clone().requires_grad_(True) for t in (word, pos, typ, gamma, beta)]
    ref = F.layer_norm(ws[0][ids] + ws[1][:L][None] + ws[2][0], (H,), ws[3], ws[4], EPS)
    assert (out.view(B, L, H).float() - ref).abs().max().item() < 3e-2
    dy = bf(torch.randn(B * L, H, generator=G(7)))
    ref.backward(dy.float().view(B, L, H))
    dword, dpos, dtyp = torch.zeros_like(word), torch.zeros_like(pos), torch.zeros_like(typ)
    dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    ops.embed_text_bwd(dy, ids, word, pos, typ[0], gamma, dword, dpos, dtyp[0], dg, db, EPS)
    for got, want, name in ((dword, ws[0].grad, "word"), (dpos, ws[1].grad, "pos"), (dtyp, ws[2].grad, "type"), (dg, ws[3].grad, "gamma"),
                            (db, ws[4].grad, "beta")):
        assert (got - want).abs().max().item() < 2e-3 * max(1.0, want.abs().max().item()), name


@pytest.mark.parametrize("variant", ["pano", "hist", "ob"])
def test_embed_feat_fwd_bwd(variant):
    ops = _ops()
    H, A = 768, 4
    M = {"pano": 36 * 20, "hist": 6 * 15, "ob": 4 * 37}[variant]
    t = bf(torch.randn(M, H, generator=G(1)))
    ang = torch.randn(M, A, generator=G(2)).cuda()
    P = {k: v.cuda() for k, v in dict(
        w_ang=0.5 * torch.randn(H, A, generator=G(3)), b_ang=0.1 * torch.randn(H, generator=G(4)),
        g_img=1 + 0.1 * torch.randn(H, generator=G(5)), b_img=0.1 * torch.randn(H, generator=G(6)),
        g_ang=1 + 0.1 * torch.randn(H, generator=G(7)), be_ang=0.1 * torch.randn(H, generator=G(8)),
        add_vec=0.1 * torch.randn(H, generator=G(9)), nav_table=0.1 * torch.randn(3, H, generator=G(10)),
        pos_table=0.1 * torch.randn(100, H, generator=G(11)), g_f=1 + 0.1 * torch.randn(H, generator=G(12)),
        b_f=0.1 * torch.randn(H, generator=G(13)), extra=torch.randn(M, H, generator=G(14))).items()}
    nav_ids = torch.randint(0, 3, (M,), generator=G(15)).cuda()
    kw = {}
    if variant == "hist":
        kw = dict(add_vec=P["add_vec"], extra=P["extra"], pos_table=P["pos_table"], pos_mod=15, g_f=P["g_f"], b_f=P["b_f"])
    if variant == "ob":
        kw = dict(add_vec=P["add_vec"], nav_table=P["nav_table"], nav_ids=nav_ids, g_f=P["g_f"], b_f=P["b_f"])
    base = (P["w_ang"], P["b_ang"], P["g_img"], P["b_img"], P["g_ang"], P["be_ang"])
    out = ops.embed_feat_fwd(t, ang, *base, eps=EPS, **kw)

    R = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    tr = t.float().requires_grad_(True)
    s = F.layer_norm(tr, (H,), R["g_img"], R["b_img"], EPS) + F.layer_norm(ang @ R["w_ang"].t() + R["b_ang"], (H,), R["g_ang"], R["be_ang"], EPS)
    if variant != "pano":
        s = s + R["add_vec"]
    if variant == "hist":
        s = s + R["extra"] + R["pos_table"][torch.arange(M, device="cuda") % 15]
    if variant == "ob":
        s = s + R["nav_table"][nav_ids]
    ref = F.layer_norm(s, (H,), R["g_f"], R["b_f"], EPS) if variant != "pano" else s
    assert (out.float() - ref).abs().max().item() < 4e-2
    dy = bf(torch.randn(M, H, generator=G(20)))
    ref.backward(dy.float())
    names = ["dw_ang", "db_ang", "dg_img", "db_img", "dg_ang", "dbe_ang", "db_lin"]
    shapes = dict(dw_ang=(H, A))
    if variant != "pano":
        names += ["dadd_vec", "dg_f", "db_f"]
    if variant == "hist":
        names += ["dpos_table"]; shapes["dpos_table"] = (100, H)
    if variant == "ob":
        names += ["dnav_table"]; shapes["dnav_table"] = (3, H)
    grads = {n: torch.zeros(shapes.get(n, (H,)), device="cuda") for n in names}
    dt, dextra = ops.embed_feat_bwd(dy, t, ang, *base, grads, eps=EPS, want_dextra=(variant == "hist"), **kw)
    sc = max(1.0, math.sqrt(M) / 4)
    assert (dt.float() - tr.grad).abs().max().item() < 4e-2 * max(1.0, tr.grad.abs().max().item())
    want = dict(dw_ang=R["w_ang"].grad, db_ang=R["b_ang"].grad, dg_img=R["g_img"].grad, db_img=R["b_img"].grad, dg_ang=R["g_ang"].grad,
                dbe_ang=R["be_ang"].grad, db_lin=tr.grad.sum(0), dadd_vec=R["add_vec"].grad, dg_f=R["g_f"].grad, db_f=R["b_f"].grad,
                dpos_table=R["pos_table"].grad, dnav_table=R["nav_table"].grad)
    for n in names:
        err = (grads[n] - want[n]).abs().max().item()
        assert err < 3e-2 * sc * max(1.0, want[n].abs().max().item() / 10), (n, err)
    if variant == "hist":
        assert (dextra - R["extra"].grad).abs().max().item() < 1e-3 * max(1.0, R["extra"].grad.abs().max().item())


def test_misc_streaming_kernels():
    ops = _ops()
    x = torch.randn(1000, 777, generator=G(1)).cuda()
    assert torch.equal(ops.cast_bf16(x), x.to(torch.bfloat16))
    a = bf(torch.randn(3000, 2304, generator=G(2)))
    out = torch.zeros(768, device="cuda")
    ops.colsum(a[:, 768:1536], out)
    assert (out - a[:, 768:1536].float().sum(0)).abs().max().item() < 1e-2
    out2 = torch.ones(1000, device="cuda")
    b = bf(torch.randn(77, 1000, generator=G(3)))
    ops.colsum(b, out2)
    assert (out2 - 1 - b.float().sum(0)).abs().max().item() < 1e-3
    t = bf(torch.randn(12 * 36, 768, generator=G(4)))
    m = ops.mean_pool_fwd(t, 12, 36)
    assert (m - t.float().view(12, 36, 768).mean(1)).abs().max().item() < 1e-5
    dm = torch.randn(12, 768, generator=G(5)).cuda()
    dx = ops.mean_pool_bwd(dm, 12, 36)
    assert (dx.float().view(12, 36, 768) - (dm / 36)[:, None]).abs().max().item() < 1e-3
    u, v = bf(torch.randn(500, 768, generator=G(6))), bf(torch.randn(500, 768, generator=G(7)))
    assert torch.equal(ops.add(u, v), (u.float() + v.float()).to(torch.bfloat16))
    w = bf(torch.randn(4, 768, generator=G(8)))
    mr = ops.mul_rows(u[:4 * 37], w, 4, 37)
    assert torch.equal(mr, (u[:148].float().view(4, 37, 768) * w.float()[:, None]).to(torch.bfloat16).view(148, 768))


def test_head_pieces():
    ops = _ops()
    M, H = 148, 768
    x = bf(torch.randn(M, H, generator=G(1)))
    for N in (1, 2, 3):
        w = (0.05 * torch.randn(N, H, generator=G(2))).cuda()
        b = torch.randn(N, generator=G(3)).cuda()
        y = ops.rowdot_fwd(x, w, b)
        xr, wr, br = x.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        yr = xr @ wr.t() + br
        assert (y - yr).abs().max().item() < 1e-3
        dy = torch.randn(M, N, generator=G(4)).cuda()
        yr.backward(dy)
        dw, db = torch.zeros_like(w), torch.zeros_like(b)
        dx = ops.rowdot_bwd(dy, x, w, dw, db)
        assert (dx.float() - xr.grad).abs().max().item() < 2e-2 * max(1.0, xr.grad.abs().max().item())
        assert (dw - wr.grad).abs().max().item() < 1e-2 and (db - br.grad).abs().max().item() < 1e-3
    # cross entropy with -inf entries (masked_fill_(nav_type==0, -inf), pretrain_cmt.py:177)
    for Mr, N in ((64, 37), (50, 30522)):
        logits = torch.randn(Mr, N, generator=G(5)).cuda()
        labels = torch.randint(0, N, (Mr,), generator=G(6)).cuda()
        if N == 37:
            msk = torch.rand(Mr, N, generator=G(7)).cuda() < 0.5
            msk[torch.arange(Mr), labels] = False
            logits = logits.masked_fill(msk, -float("inf"))
        lr = logits.clone().requires_grad_(True)
        ref = F.cross_entropy(lr, labels, reduction="none")
        loss, lse = ops.ce_fwd(logits, labels)
        assert (loss - ref).abs().max().item() < 1e-4
        g = torch.rand(Mr, generator=G(8)).cuda()
        ref.backward(g)
        d = ops.ce_bwd(logits, labels, lse, g)
        assert (d - lr.grad).abs().max().item() < 1e-5
        dbf = ops.ce_bwd(logits, labels, lse, g, bf16_padded=True)
        assert dbf.shape[1] % 8 == 0 and (dbf[:, :N].float() - lr.grad).abs().max().item() < 1e-2
        assert dbf[:, N:].abs().max().item() == 0 if dbf.shape[1] > N else True
    idx = torch.tensor([5, 1, 99, 42], device="cuda")
    gth = ops.gather_rows(x, idx)
    assert torch.equal(gth, x[idx])
    sc = ops.scatter_rows(gth, idx, M)
    assert torch.equal(sc[idx], x[idx]) and sc.abs().sum().item() == x[idx].abs().sum().item()


@pytest.mark.parametrize("M", [1, 37, 1500, 34560])
@pytest.mark.parametrize("H", [768, 512])
def test_ln_bwd_persistent_grid_all_row_counts(M, H):
    """ln_bwd is a one-wave persistent kernel (rows strided over all warps of the chip, next row prefetched): fewer rows than
    warps, ragged tails and the pano-sized M all against fp32 torch, with dropout and an incoming residual-path gradient."""
    ops = _ops()
    x, res, dy, dri = (bf(torch.randn(M, H, generator=G(s))) for s in (1, 2, 3, 4))
    gamma, beta = (1 + 0.1 * torch.randn(H, generator=G(5))).cuda(), torch.zeros(H, device="cuda")
    seed = torch.tensor([77], dtype=torch.int64, device="cuda")
    drop = ops.Drop(seed, site=3, p=0.1)
    _, z, mean, rstd = ops.ln_fwd(x.clone(), res, gamma, beta, EPS, drop)
    ones = torch.ones_like(x)
    _, keep, _, _ = ops.ln_fwd(ones, None, torch.ones(H, device="cuda"), torch.zeros(H, device="cuda"), EPS, drop)   # z = dropout(1)
    sums = [torch.full((H,), 0.5, device="cuda") for _ in range(3)]
    dx, dres = ops.ln_bwd(dy, z, mean, rstd, gamma, *sums, dres_in=dri, drop=drop)
    torch.cuda.synchronize()
    zf, dyf = z.float(), dy.float()
    xh = (zf - mean[:, None]) * rstd[:, None]
    gd = dyf * gamma
    dz = rstd[:, None] * (gd - gd.mean(1, keepdim=True) - xh * (gd * xh).mean(1, keepdim=True))
    tol = 2e-2
    assert (dres.float() - (dz + dri.float())).abs().max().item() <= tol * max(1.0, dz.abs().max().item())
    dxr = dz * keep.float()
    assert (dx.float() - dxr).abs().max().item() <= tol * max(1.0, dxr.abs().max().item())
    for got, want in zip(sums, ((dyf * xh).sum(0), dyf.sum(0), dxr.sum(0))):
        assert (got - 0.5 - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item()) * (1 + M ** 0.5 / 8)
