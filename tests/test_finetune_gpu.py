"""Finetune path on the GPU: `NavCMT.forward(mode)` in all its variants -- forward AND the backward the
agent trains through (finetune_src/r2r/agent_cmt.py:562 runs dozens of 'language' / 'history' / 'visual' forwards and then ONE
loss.backward()) -- and the `VLNBertCMT` wrapper with the `get_vlnbert_models` checkpoint-key remap
(finetune_src/models/model_HAMT.py:20-65, vlnbert_init.py:22-31), against the CPU oracle (pinned to the unmodified reference
NavCMT in tests/test_oracle.py)."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEPTH = dict(num_l_layers=2, num_x_layers=2, num_h_pano_layers=1)
VARIANTS = {
    "ob_txt": dict(),
    "no_lang_ca": dict(no_lang_ca=True),
    "ob": dict(act_pred_token="ob"),
    "ob_hist": dict(act_pred_token="ob_hist"),
    "ob_txt_hist": dict(act_pred_token="ob_txt_hist"),
}


def _cfg(**over):
    import hamt_b200  # noqa: F401
    from hamt_b200.config import HamtConfig
    return HamtConfig(**dict(DEPTH, hist_enc_pano=True, output_attentions=True, **over))


def _navcmt(cfg, seed=11):
    from hamt_b200 import synth
    from hamt_b200.vilmodel_cmt import NavCMT
    model = NavCMT(cfg)
    sd = synth.seeded_state_dict(model, seed=seed)
    model.load_state_dict(sd)
    return model.cuda(), sd


def _dev(b):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}


def _episode(model, b, T):
    """language -> history x (T + 1) -> visual, exactly the agent's call sequence (agent_cmt.py:275-474)."""
    B = b["txt_ids"].shape[0]
    txt = model("language", txt_ids=b["txt_ids"], txt_masks=b["txt_masks"])
    hs = [model("history").expand(B, -1)]
    for t in range(T):
        hs.append(model("history", hist_img_feats=b["hist_img_fts"][:, t], hist_ang_feats=b["hist_ang_fts"][:, t],
                        ob_step_ids=torch.LongTensor([t]).to(b["txt_ids"].device), hist_pano_img_feats=b["hist_pano_img_fts"][:, t],
                        hist_pano_ang_feats=b["hist_pano_ang_fts"][:, t]))
    hist = torch.stack(hs, 1)
    hm = torch.ones(B, T + 1, dtype=torch.bool, device=hist.device)
    return model("visual", txt_embeds=txt, txt_masks=b["txt_masks"], hist_embeds=hist, hist_masks=hm, ob_img_feats=b["ob_img_fts"],
                 ob_ang_feats=b["ob_ang_fts"], ob_nav_types=b["ob_nav_types"], ob_masks=b["ob_masks"])


def _oracle_episode(sd, cfg, b, T, rg):
    from oracle import hamt_oracle as O
    B = b["txt_ids"].shape[0]
    txt = O.navcmt_language(sd, cfg, b["txt_ids"], b["txt_masks"], rg=rg)
    hs = [O.navcmt_history(sd, cfg, None, None, None, rg=rg).expand(B, -1)]
    for t in range(T):
        hs.append(O.navcmt_history(sd, cfg, b["hist_img_fts"][:, t], b["hist_ang_fts"][:, t], torch.LongTensor([t]),
                                   b["hist_pano_img_fts"][:, t], b["hist_pano_ang_fts"][:, t], rg=rg))
    hist = torch.stack(hs, 1)
    hm = torch.ones(B, T + 1, dtype=torch.bool)
    return O.navcmt_visual(sd, cfg, txt, b["txt_masks"], hist, hm, b["ob_img_fts"], b["ob_ang_fts"], b["ob_nav_types"], b["ob_masks"], rg=rg)


def _proj(logits, seed=100):
    w = torch.randn(logits.shape, generator=torch.Generator().manual_seed(seed)).to(logits.device)
    fin = torch.isfinite(logits)
    return (torch.where(fin, logits.float(), torch.zeros_like(logits, dtype=torch.float32)) * w).sum() / max(1, int(fin.sum()))


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_navcmt_episode_forward_and_backward_vs_oracle(variant):
    """Every `act_pred_token` variant and `no_lang_ca`: eval logits vs the fp32 oracle (argmax equal), then the train-mode episode
    (dropout probabilities 0) differentiated through 'visual', the history steps and 'language': every parameter gradient -- x_layers,
    img_embeddings, hist_embeddings (incl. the pano encoder), the text layers, next_action -- vs autograd through the oracle."""
    from hamt_b200 import synth
    from oracle import hamt_oracle as O
    cfg = _cfg(**VARIANTS[variant])
    model, sd = _navcmt(cfg)
    B, L, T, Ob = 3, 16, 2, 11
    b = synth.make_batch("sap", batch_size=B, txt_len=L, hist_len=T, n_ob=Ob, seed=7, ragged=True)
    model.eval()
    with torch.no_grad():
        out = _episode(model, _dev(b), T)
        ref = _oracle_episode(sd, cfg, b, T, O.FP32)
    got, want = out[0].float().cpu(), ref[0]
    fin = torch.isfinite(want)
    assert torch.equal(fin, torch.isfinite(got))
    err = (got[fin] - want[fin]).abs().max().item()
    assert err < 3e-2, (variant, err)
    top2 = want.topk(2, dim=1).values
    ok = (top2[:, 0] - top2[:, 1]) > 2 * err
    assert torch.equal(got.argmax(1)[ok], want.argmax(1)[ok])

    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.zero_grad(set_to_none=True)
    loss = _proj(_episode(model, _dev(b), T)[0])
    loss.backward()
    torch.cuda.synchronize()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref_loss = _proj(_oracle_episode(sdr, cfg, b, T, O.BF16)[0])
    assert abs(float(loss) - float(ref_loss)) < 1e-2 * max(1.0, abs(float(ref_loss)))
    ref_loss.backward()
    bad, checked, groups = [], 0, set()
    for k, p in model.named_parameters():
        g_ref = sdr[k].grad
        if g_ref is None or g_ref.abs().max().item() == 0:
            assert p.grad is None or p.grad.abs().max().item() < 1e-6, f"{k}: gradient where the oracle has none"
            continue
        assert p.grad is not None, f"{k}: missing gradient"
        floor = 1e-3 * (g_ref.numel() ** 0.5)
        rel = (p.grad.float().cpu() - g_ref).norm().item() / max(g_ref.norm().item(), floor)
        checked += 1
        groups.add(k.split(".")[0] + "." + k.split(".")[1])
        if rel > 0.10:
            bad.append((k, round(rel, 4)))
    assert checked > 40 and {"encoder.x_layers", "img_embeddings.img_linear", "next_action.net", "hist_embeddings.pano_encoder"} <= groups, groups
    if not VARIANTS[variant].get("no_lang_ca"):
        assert "encoder.layer" in groups           # the gradient reaches the text layers through txt_embeds
    assert not bad, f"{variant}: gradient mismatch ({len(bad)} of {checked}): {sorted(bad, key=lambda t: -t[1])[:8]}"


def test_backward_after_later_forwards_regenerates_the_same_dropout_masks():
    """every forward owns a private seed cell.  forward A -> forward B -> backward A must give the gradients of
    forward A -> backward A (dropout ON): the masks regenerated in A's backward are A's, not the ones of the last forward."""
    from hamt_b200 import synth
    cfg = _cfg()
    model, _ = _navcmt(cfg)
    model.train()
    b1 = _dev(synth.make_batch("sap", batch_size=2, txt_len=16, hist_len=1, seed=1))
    b2 = _dev(synth.make_batch("sap", batch_size=2, txt_len=16, hist_len=1, seed=2))

    def run(extra_forward):
        model.zero_grad(set_to_none=True)
        model.arena().ensure()
        model.arena().seed.fill_(12345)
        a = model("language", txt_ids=b1["txt_ids"], txt_masks=b1["txt_masks"])
        if extra_forward:
            model("language", txt_ids=b2["txt_ids"], txt_masks=b2["txt_masks"])
            model("history")
        a.float().square().mean().backward()
        torch.cuda.synchronize()
        return a.detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

    a0, g0 = run(False)
    a1, g1 = run(True)
    assert torch.equal(a0, a1)
    assert set(g0) <= set(g1) and len(g0) > 20
    for n in g0:
        assert (g0[n] - g1[n]).abs().max().item() <= 1e-3 * g0[n].abs().max().item() + 1e-7, n
    for n in set(g1) - set(g0):          # parameters only the extra forwards touched (history CLS): attached, never back-propagated into
        assert float(g1[n].abs().max()) == 0.0, n


def _args(ckpt, **over):
    d = dict(bert_ckpt_file=ckpt, dataset="r2r", tokenizer="bert", image_feat_size=768, angle_feat_size=4, num_l_layers=2, num_h_layers=0,
             num_x_layers=2, hist_enc_pano=True, hist_pano_num_layers=1, fix_lang_embedding=False, fix_hist_embedding=False,
             fix_obs_embedding=False, no_lang_ca=False, act_pred_token="ob_txt", feat_dropout=0.4, dropout=0.5)
    d.update(over)
    return types.SimpleNamespace(**d)


@pytest.mark.parametrize("flavour", ["plain", "ddp_module_prefix"])
def test_vlnbert_cmt_checkpoint_remap_and_modes(flavour, tmp_path):
    """`get_vlnbert_models` (vlnbert_init.py:22-31): a pretrain checkpoint saved from MultiStepNavCMTPreTraining (keys `bert.*`,
    `next_action.*`, ...; or all prefixed `module.` when saved from DDP) lands on the NavCMT attributes; then `VLNBertCMT.forward` in
    its three modes: `length2mask`, `torch.stack` of the history list, `states = txt[:, 0] * hist[:, 0]` (model_HAMT.py:20-65)."""
    from hamt_b200 import synth
    from hamt_b200.config import HamtConfig
    from hamt_b200.model_HAMT import Critic, VLNBertCMT
    from hamt_b200.pretrain_cmt import MultiStepNavCMTPreTraining
    from oracle import hamt_oracle as O
    pre = MultiStepNavCMTPreTraining(HamtConfig(**DEPTH))
    sd_pre = synth.seeded_state_dict(pre, seed=5)
    ckpt = {("module." + k if flavour == "ddp_module_prefix" else k): v for k, v in sd_pre.items()}
    path = os.path.join(tmp_path, "model_step_1.pt")
    torch.save(ckpt, path)
    vln = VLNBertCMT(_args(path)).cuda()
    nav = vln.vln_bert
    want = {"embeddings.word_embeddings.weight": "bert.embeddings.word_embeddings.weight",
            "encoder.x_layers.1.visual_attention.att.key.bias": "bert.encoder.x_layers.1.visual_attention.att.key.bias",
            "hist_embeddings.pano_encoder.layer.0.output.dense.weight": "bert.hist_embeddings.pano_encoder.layer.0.output.dense.weight",
            "img_embeddings.nav_type_embedding.weight": "bert.img_embeddings.nav_type_embedding.weight",
            "next_action.net.0.weight": "next_action.net.0.weight", "next_action.net.4.bias": "next_action.net.4.bias"}
    got_sd = nav.state_dict()
    for k_nav, k_pre in want.items():
        assert torch.equal(got_sd[k_nav].cpu(), sd_pre[k_pre]), (k_nav, k_pre)
    assert not any(k.startswith(("mlm_head", "itm_head", "bert.")) for k in got_sd)

    # ---- the three modes, eval (drop_env inactive), against the oracle on the loaded weights
    sd_nav = {k: v.detach().cpu().clone() for k, v in got_sd.items()}
    cfg = nav.config
    B, L, T, Ob = 3, 16, 2, 11
    b = synth.make_batch("sap", batch_size=B, txt_len=L, hist_len=T, n_ob=Ob, seed=9, ragged=True)
    bd = _dev(b)
    vln.eval()
    lens = [3, 1, 2]                                   # history tokens (incl. the CLS slot) valid per sample
    with torch.no_grad():
        txt = vln("language", txt_ids=bd["txt_ids"], txt_masks=bd["txt_masks"])
        hs = [vln("history").expand(B, -1)]
        for t in range(T):
            hs.append(vln("history", hist_img_feats=bd["hist_img_fts"][:, t], hist_ang_feats=bd["hist_ang_fts"][:, t], ob_step=t,
                          hist_pano_img_feats=bd["hist_pano_img_fts"][:, t], hist_pano_ang_feats=bd["hist_pano_ang_fts"][:, t]))
        logits, states = vln("visual", txt_embeds=txt, txt_masks=bd["txt_masks"], hist_embeds=hs, hist_lens=lens, ob_img_feats=bd["ob_img_fts"],
                             ob_ang_feats=bd["ob_ang_fts"], ob_nav_types=bd["ob_nav_types"], ob_masks=bd["ob_masks"], return_states=True)
        (logits_only,) = vln("visual", txt_embeds=txt, txt_masks=bd["txt_masks"], hist_embeds=hs, hist_lens=lens, ob_img_feats=bd["ob_img_fts"],
                             ob_ang_feats=bd["ob_ang_fts"], ob_nav_types=bd["ob_nav_types"], ob_masks=bd["ob_masks"])
        # oracle: same call sequence; hist_masks = not length2mask(lens) (True = valid)
        o_txt = O.navcmt_language(sd_nav, cfg, b["txt_ids"], b["txt_masks"])
        o_hs = [O.navcmt_history(sd_nav, cfg, None, None, None).expand(B, -1)]
        for t in range(T):
            o_hs.append(O.navcmt_history(sd_nav, cfg, b["hist_img_fts"][:, t], b["hist_ang_fts"][:, t], torch.LongTensor([t]),
                                         b["hist_pano_img_fts"][:, t], b["hist_pano_ang_fts"][:, t]))
        hm = torch.arange(T + 1)[None] < torch.tensor(lens)[:, None]
        o_logits, o_txt2, o_hist, _ = O.navcmt_visual(sd_nav, cfg, o_txt, b["txt_masks"], torch.stack(o_hs, 1), hm, b["ob_img_fts"], b["ob_ang_fts"],
                                                      b["ob_nav_types"], b["ob_masks"])
    assert torch.equal(logits, logits_only)
    got, ref = logits.float().cpu(), o_logits
    fin = torch.isfinite(ref)
    assert torch.equal(fin, torch.isfinite(got))
    err = (got[fin] - ref[fin]).abs().max().item()
    assert err < 3e-2, err
    top2 = ref.topk(2, dim=1).values
    ok = (top2[:, 0] - top2[:, 1]) > 2 * err
    assert torch.equal(got.argmax(1)[ok], ref.argmax(1)[ok])
    o_states = o_txt2[:, 0] * o_hist[:, 0]
    assert tuple(states.shape) == (B, 768)
    assert (states.float().cpu() - o_states).abs().max().item() < 8e-2 * max(1.0, o_states.abs().max().item())
    # the mask matters: with every history token visible the logits of the short-history samples change
    with torch.no_grad():
        (full,) = vln("visual", txt_embeds=txt, txt_masks=bd["txt_masks"], hist_embeds=hs, hist_lens=[T + 1] * B, ob_img_feats=bd["ob_img_fts"],
                      ob_ang_feats=bd["ob_ang_fts"], ob_nav_types=bd["ob_nav_types"], ob_masks=bd["ob_masks"])
    fin1 = torch.isfinite(full[1])
    assert not torch.equal(full[1][fin1], logits[1][fin1]) and torch.equal(full[0], logits[0])

    # ---- train mode: feature dropout (drop_env, p = 0.4) is active, the episode trains end to end, the critic consumes `states`
    vln.train()
    critic = Critic(_args(path)).cuda()
    txt = vln("language", txt_ids=bd["txt_ids"], txt_masks=bd["txt_masks"])
    hs = [vln("history").expand(B, -1), vln("history", hist_img_feats=bd["hist_img_fts"][:, 0], hist_ang_feats=bd["hist_ang_fts"][:, 0], ob_step=0,
                                             hist_pano_img_feats=bd["hist_pano_img_fts"][:, 0], hist_pano_ang_feats=bd["hist_pano_ang_fts"][:, 0])]
    logits_t, states_t = vln("visual", txt_embeds=txt, txt_masks=bd["txt_masks"], hist_embeds=hs, hist_lens=[2, 1, 2], ob_img_feats=bd["ob_img_fts"],
                             ob_ang_feats=bd["ob_ang_fts"], ob_nav_types=bd["ob_nav_types"], ob_masks=bd["ob_masks"], return_states=True)
    target = torch.tensor([int(r.nonzero()[0]) for r in (b["ob_nav_types"] > 0)], device="cuda")
    loss = torch.nn.functional.cross_entropy(logits_t.float(), target) + critic(states_t).square().mean()
    loss.backward()
    torch.cuda.synchronize()
    for name in ("encoder.x_layers.0.visn_self_att.self.query.weight", "img_embeddings.img_linear.weight", "next_action.net.0.weight",
                 "hist_embeddings.pano_img_linear.weight", "encoder.layer.0.attention.self.query.weight"):
        g = dict(nav.named_parameters())[name].grad
        assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0, name
    assert critic.state2value[0].weight.grad is not None
