"""End-to-end ViT stage.

CPU: the oracle restatement (oracle/vit_oracle.py) against the golden vectors generated from the UNMODIFIED reference class
(oracle/make_golden_vit.py) and, where the reference checkout exists, against the reference run live; state_dict keys of the CUDA
module against the reference's.
GPU (-m gpu): the row kernels against torch, the backbone forward against the fp32 golden bounded by the reference's own autocast
error, its parameter gradients against autograd through the oracle, and the image model (backbone -> cross-modal transformer)
against the feature model fed with the oracle's features.
"""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ref_vit(depth):
    from oracle import ref_shim
    return ref_shim.load_reference_vit(depth=depth).eval()


def _case(name):
    return torch.load(os.path.join(GOLD, "vit.pt"))[name]


def _weights(depth, wseed, num_classes=0):
    import hamt_b200  # noqa: F401
    from hamt_b200 import synth
    from hamt_b200.vision_transformer import VisionTransformer
    m = VisionTransformer(depth=depth, num_classes=num_classes)
    return m, synth.seeded_vit_state_dict(m, wseed)


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("name", ["vit_d2_n5", "vit_d12_n3"])
def test_vit_oracle_reproduces_reference_golden(name):
    from hamt_b200 import synth
    from oracle import vit_oracle as V
    c = _case(name)
    _, sd = _weights(c["depth"], c["wseed"])
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = synth.make_images(c["n"], c["iseed"])
    f = V.forward_features(sd, x)
    assert (f - c["feats"]).abs().max().item() < 5e-5
    w = torch.linspace(-1, 1, f.numel()).view_as(f)
    (f * w).sum().backward()
    for k, g in c["grads"].items():
        got = sd[k].grad.reshape(-1)
        assert abs(got.norm().item() - g["norm"]) <= 1e-4 * g["norm"] + 1e-6, k
        assert (got[:512] - g["head"]).abs().max().item() <= 1e-4 * g["head"].abs().max().item() + 1e-6, k


def test_vit_oracle_matches_reference_live_and_state_dict_keys():
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference checkout not available")
    from hamt_b200 import synth
    from oracle import vit_oracle as V
    ref = _ref_vit(3)
    mine, sd = _weights(3, 4)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    assert [tuple(v.shape) for v in ref.state_dict().values()] == [tuple(v.shape) for v in mine.state_dict().values()]
    ref.load_state_dict(sd)
    x = synth.make_images(2, 9)
    with torch.no_grad():
        want = ref.forward_features(x)
        got = V.forward_features(sd, x)
    assert (got - want).abs().max().item() < 2e-5
    # the 6-D panorama path of forward_vision_backbone (image_vilmodel.py:40-59) is forward_features over the flattened views
    imgs = synth.make_images(4, 3).view(1, 2, 2, 3, 224, 224)
    with torch.no_grad():
        want6 = ref.forward_features(imgs.view(4, 3, 224, 224)).view(1, 2, 2, -1)
    got6 = V.forward_vision_backbone(sd, imgs)
    assert not got6.requires_grad and (got6 - want6).abs().max().item() < 2e-5


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_patchify_and_embed_kernels_match_torch():
    import hamt_b200  # noqa: F401
    from hamt_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(3, 3, 224, 224, device="cuda")
    got = ops.patchify(x, 16)
    want = x.view(3, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(3 * 196, 768).to(torch.bfloat16)
    assert torch.equal(got, want)
    # conv == patchify + GEMM
    w = torch.randn(768, 3, 16, 16, device="cuda") * 0.02
    b = torch.randn(768, device="cuda") * 0.1
    y = ops.gemm(got, w.view(768, -1).to(torch.bfloat16), bias=b).float()
    ref = torch.nn.functional.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), b, stride=16).flatten(2).transpose(1, 2).reshape(-1, 768)
    assert (y - ref).abs().max().item() < 2e-2 * ref.abs().max().item()
    # embed: [cls ; tokens] + pos
    N, S, H = 3, 197, 768
    t0 = torch.randn(N * (S - 1), H, device="cuda").to(torch.bfloat16)
    cls, pos = torch.randn(H, device="cuda"), torch.randn(S * H, device="cuda")
    x32, x16 = ops.vit_embed_fwd(t0, cls, pos, N, S)
    want32 = torch.cat([cls.view(1, 1, H).expand(N, 1, H), t0.float().view(N, S - 1, H)], 1) + pos.view(1, S, H)
    assert torch.equal(x32.view(N, S, H), want32) and torch.equal(x16, want32.view(-1, H).to(torch.bfloat16))
    dx = torch.randn(N * S, H, device="cuda").to(torch.bfloat16)
    dfull, dt0 = ops.vit_embed_bwd(dx, N, S)
    assert torch.equal(dfull, dx) and torch.equal(dt0.view(N, S - 1, H), dx.view(N, S, H)[:, 1:])
    # dropout: forward and backward use the same mask, keep rate ~ 0.9, scale 1 / 0.9
    d = ops.Drop(torch.tensor([12345], dtype=torch.int64, device="cuda"), 3, 0.1)
    x32d, _ = ops.vit_embed_fwd(t0, cls, pos, N, S, d)
    kept = x32d != 0
    assert 0.88 < kept.float().mean().item() < 0.92
    assert torch.allclose(x32d[kept], (want32.view(-1, H) / 0.9)[kept], rtol=1e-6, atol=1e-6)
    dfull_d, _ = ops.vit_embed_bwd(dx, N, S, d)
    assert torch.equal(dfull_d != 0, kept & (dx != 0))


@pytest.mark.gpu
def test_ln_prenorm_kernel_matches_torch():
    import hamt_b200  # noqa: F401
    from hamt_b200 import ops
    torch.manual_seed(1)
    M, H = 1000, 768
    t = torch.randn(M, H, device="cuda").to(torch.bfloat16)
    x32 = torch.randn(M, H, device="cuda") * 3
    g, b = torch.rand(H, device="cuda") + 0.5, torch.randn(H, device="cuda") * 0.1
    want_z = t.float() + x32
    want_y = torch.nn.functional.layer_norm(want_z, (H,), g, b, 1e-6)
    y, y32, z16, z32, mean, rstd = ops.ln_fwd_prenorm(t.clone(), x32, g, b, 1e-6, want_y32=True)
    assert torch.equal(z32, want_z) and torch.equal(z16, want_z.to(torch.bfloat16))
    assert (y32 - want_y).abs().max().item() < 2e-5 and torch.equal(y, y32.to(torch.bfloat16))
    assert (mean - want_z.mean(1)).abs().max().item() < 1e-5
    # x = None: plain LayerNorm of the fp32 stream
    y0, y032, z0, z032, _, _ = ops.ln_fwd_prenorm(None, x32, g, b, 1e-6, want_y32=True)
    assert z0 is None and z032 is None
    assert (y032 - torch.nn.functional.layer_norm(x32, (H,), g, b, 1e-6)).abs().max().item() < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["vit_d2_n5", "vit_d12_n3"])
def test_vit_forward_and_gradients_vs_reference_golden(name):
    """Forward: |ours - reference fp32| <= 1.5 x |reference autocast(bf16) - reference fp32| (+ 2e-3).  Gradients of
    sum(features * w) wrt a spread of parameters against the reference's fp32 autograd: relative error of the norm and of the leading
    512 entries."""
    from hamt_b200 import synth
    c = _case(name)
    m, sd = _weights(c["depth"], c["wseed"])
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = synth.make_images(c["n"], c["iseed"]).cuda()
    f = m.forward_features(x)
    assert f.dtype == torch.float32 and f.shape == (c["n"], 768)
    err = (f.detach().cpu() - c["feats"]).abs().max().item()
    yard = (c["feats_autocast"] - c["feats"]).abs().max().item()
    assert err <= 1.5 * yard + 2e-3, (name, err, yard)
    w = torch.linspace(-1, 1, f.numel()).view_as(f).cuda()
    (f * w).sum().backward()
    torch.cuda.synchronize()
    params = dict(m.named_parameters())
    worst, report = 0.0, []
    for k, g in c["grads"].items():
        got = params[k].grad.detach().float().cpu().reshape(-1)
        assert torch.isfinite(got).all(), k
        rel_norm = abs(got.norm().item() - g["norm"]) / g["norm"]
        rel_head = (got[:512] - g["head"]).norm().item() / max(g["head"].norm().item(), 1e-12)
        worst = max(worst, rel_norm, rel_head)
        report.append((k, round(rel_norm, 4), round(rel_head, 4)))
    assert all(rn < 0.05 and rh < 0.08 for _, rn, rh in report), (name, report)
    os.makedirs(os.path.join(os.path.dirname(GOLD), "..", "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(GOLD), "..", "gpurun_out", f"parity_{name}.txt"), "w") as fh:
        fh.write(f"{name}: features max|ours - ref_fp32| = {err:.3e}, reference autocast yardstick = {yard:.3e}, ratio {err / yard:.2f}; "
                 f"worst relative gradient error = {worst:.3e}\n")


@pytest.mark.gpu
def test_vit_train_mode_dropout_runs_and_no_grad_path():
    """Train mode with the reference's drop rates (image_vilmodel.py:26-29: 0.1 / 0.1): finite outputs, fresh masks per call, finite
    gradients for every parameter on the path; no_grad call saves nothing and matches eval within dropout-free arithmetic."""
    from hamt_b200 import synth
    from hamt_b200.vision_transformer import VisionTransformer
    m = VisionTransformer(depth=2, num_classes=0, drop_rate=0.1, attn_drop_rate=0.1)
    m.load_state_dict(synth.seeded_vit_state_dict(m, 3))
    m = m.cuda().train()
    x = synth.make_images(4, 1).cuda()
    f1 = m.forward_features(x)
    f1.square().mean().backward()
    f2 = m.forward_features(x)
    assert torch.isfinite(f1).all() and not torch.equal(f1, f2)
    for k, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum().item() > 0, k
    m.eval()
    with torch.no_grad():
        e1 = m.forward_features(x)
    e2 = m.forward_features(x)
    assert torch.equal(e1, e2.detach())


@pytest.mark.gpu
def test_image_model_matches_feature_model_on_backbone_features():
    """NavImagePreTrainedModel (image_vilmodel.py:61-123) = backbone + the feature model: its outputs must equal the feature model's
    on the backbone's own features (STOP row appended, MRC-masked steps zeroed), and the loss gradient must reach the backbone."""
    from hamt_b200 import synth
    from hamt_b200.config import HamtConfig
    from hamt_b200.image_vilmodel import NavImagePreTrainedModel
    cfg = HamtConfig(num_l_layers=1, num_x_layers=1, num_h_pano_layers=1, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    torch.manual_seed(0)
    m = NavImagePreTrainedModel(cfg, vit_depth=1).cuda().eval()
    B, L, T, P, O = 2, 12, 2, 3, 4
    g = torch.Generator().manual_seed(5)
    txt_ids = torch.randint(1000, 5000, (B, L), generator=g).cuda()
    txt_masks = torch.ones(B, L, dtype=torch.bool).cuda()
    hist_images = synth.make_images(B * T, 1).view(B, T, 3, 224, 224).cuda()
    pano_images = synth.make_images(B * T * P, 2).view(B, T, P, 3, 224, 224).cuda()
    ob_images = synth.make_images(B * (O - 1), 3).view(B, O - 1, 3, 224, 224).cuda()
    hist_ang, pano_ang, ob_ang = torch.randn(B, T, 4, generator=g).cuda(), torch.randn(B, T, P, 4, generator=g).cuda(), torch.randn(B, O, 4, generator=g).cuda()
    hist_masks = torch.ones(B, T + 1, dtype=torch.bool).cuda()
    ob_nav = torch.randint(0, 3, (B, O), generator=g).cuda()
    ob_masks = torch.ones(B, O, dtype=torch.bool).cuda()
    mrc = torch.tensor([[True, False], [False, False]]).cuda()
    txt, hist, ob = m(txt_ids, txt_masks, hist_images, hist_ang, pano_images, pano_ang, hist_masks, ob_images, ob_ang, ob_nav, ob_masks, hist_mrc_masks=mrc)
    assert txt.shape == (B, L, 768) and hist.shape == (B, T + 1, 768) and ob.shape == (B, O, 768)
    loss = txt.float().square().mean() + ob.float().square().mean() + hist.float().square().mean()
    loss.backward()
    gw = m.vision_backbone.blocks[0].attn.qkv.weight.grad
    assert gw is not None and torch.isfinite(gw).all() and gw.abs().sum().item() > 0
    assert m.vision_backbone.patch_embed.proj.weight.grad.abs().sum().item() > 0
    with torch.no_grad():
        hf = m.forward_vision_backbone(hist_images).masked_fill(mrc.unsqueeze(-1), 0)
        pf = m.forward_vision_backbone(pano_images).masked_fill(mrc.unsqueeze(-1).unsqueeze(-1), 0)
        of = torch.cat([m.forward_vision_backbone(ob_images), torch.zeros(B, 1, 768, device="cuda")], 1)
        from hamt_b200.vilmodel import NavPreTrainedModel
        t2, h2, o2 = NavPreTrainedModel.forward(m, txt_ids, txt_masks, hf, hist_ang, pf, pano_ang, hist_masks, of, ob_ang, ob_nav, ob_masks)
    assert torch.equal(t2, txt.detach()) and torch.equal(h2, hist.detach()) and torch.equal(o2, ob.detach())


@pytest.mark.gpu
def test_image_pretraining_model_sap_and_mrc_steps():
    """MultiStepNavImagePreTraining (image_pretrain.py:18-90): image batch -> backbone -> heads.  SAP logits must equal the feature model's
    logits on the backbone's own features; a train-mode MRC step (history views zeroed where masked, image_vilmodel.py:80-82) and a SAP step
    give finite losses and reach the backbone's parameters."""
    from hamt_b200 import synth
    from hamt_b200.config import HamtConfig
    from hamt_b200.image_pretrain import MultiStepNavImagePreTraining
    cfg = HamtConfig(num_l_layers=1, num_x_layers=1, num_h_pano_layers=1)
    m = MultiStepNavImagePreTraining(cfg, vit_depth=1)
    sd = synth.seeded_state_dict(m, seed=2)
    sd.update({"bert.vision_backbone." + k: v for k, v in synth.seeded_vit_state_dict(m.bert.vision_backbone, 3).items()})
    m.load_state_dict(sd)
    m = m.cuda().eval()
    b = synth.make_image_batch("sap", batch_size=2, txt_len=16, hist_len=2, n_pano=3, n_ob=5, seed=4, device="cuda")
    with torch.no_grad():
        logits = m(b, "sap", compute_loss=False)
        fb = {k: v for k, v in b.items() if not k.endswith("images") and k != "ob_v_exists"}
        fb["hist_img_fts"] = m.bert.forward_vision_backbone(b["hist_images"])
        fb["hist_pano_img_fts"] = m.bert.forward_vision_backbone(b["hist_pano_images"])
        of = m.bert.forward_vision_backbone(b["ob_images"])
        fb["ob_img_fts"] = torch.cat([of, torch.zeros(2, 1, 768, device="cuda")], 1)
        want = m(fb, "sap", compute_loss=False)
    assert logits.shape == (2, 5) and torch.equal(logits, want)
    m.train()
    for task in ("sap", "mrc"):
        bt = synth.make_image_batch(task, batch_size=2, txt_len=16, hist_len=2, n_pano=3, n_ob=5, seed=6, device="cuda")
        loss = m(bt, task, compute_loss=True)
        assert torch.isfinite(loss).all()
        loss.mean().backward()
        g = m.bert.vision_backbone.blocks[0].mlp.fc1.weight.grad
        assert g is not None and torch.isfinite(g).all() and g.abs().sum().item() > 0
        m.zero_grad(set_to_none=True)
