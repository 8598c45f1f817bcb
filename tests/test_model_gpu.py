"""Model-level parity of the sm_100a path, called through the reference-facing module API (which goes through the C ABI).

Comparators
  (a) golden vectors produced by the UNMODIFIED reference in fp32 (tests/golden, oracle/make_golden.py);
  (b) the CPU oracle run in the same dtype regime (bf16 rounding points mirrored, oracle.BF16).

What can and cannot be asserted (measured, see DESIGN.md "Parity"): a 13..15-layer bf16 pipeline is chaotic at the rounding
level -- perturbing the weights by 1e-7 relative moves the bf16-regime logits by ~1.8e-2, the same size as the bf16-vs-fp32 gap
(tests/test_oracle.py::test_bf16_regime_rounding_chaos).  The reference's own torch.autocast(bf16) run differs from its fp32 run by
1.2e-2 (SAP) .. 3e-2 (MLM).  So "< 1e-3 in bf16" is asserted where it is attainable -- per kernel (tests/test_*_gpu.py) -- and
at model level the assertions are:
   * logits:  |ours - reference_fp32| <= TOL_LOGITS, and not worse than RATIO x the bf16-regime oracle's own distance to fp32;
   * losses:  |ours - reference_fp32| <= TOL_LOSS * max(1, |ref|);
   * SAP / finetune action argmax bit-exact wherever the reference's top-2 margin exceeds 2 x TOL_LOGITS, and always equal to
     the reference when the margin is that large;
   * -inf patterns (masked_fill_(nav_type == 0)) identical;  integer outputs bit-exact.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TASKS = ("mlm", "sap", "sar", "sprel", "mrc", "itm")
TOL_LOGITS = 4e-2     # max-abs, logits (std 0.3 .. 0.55) vs the fp32 reference; measured 0.8e-2 .. 2.4e-2
TOL_LOSS = 1e-1       # un-reduced loss entries (MSE on angles amplifies a logit error by 2|pred - target| <= 2 pi)
RATIO = 3.0           # ours may be at most this many times further from fp32 than the bf16-regime oracle (+ 5e-3 slack)


def _build(cfg_over, weight_seed, device="cuda"):
    import hamt_b200  # noqa: F401
    from hamt_b200 import synth
    from hamt_b200.config import HamtConfig
    from hamt_b200.pretrain_cmt import MultiStepNavCMTPreTraining
    cfg = HamtConfig(**cfg_over)
    model = MultiStepNavCMTPreTraining(cfg)
    sd = synth.seeded_state_dict(model, seed=weight_seed)
    model.load_state_dict(sd)
    return cfg, model.to(device), sd


def _to_dev(b, dev="cuda"):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}


def _err(got, want):
    got, want = got.float().cpu(), want.float().cpu()
    fin = torch.isfinite(want)
    assert torch.equal(fin, torch.isfinite(got)), "-inf pattern differs"
    return (got[fin] - want[fin]).abs().max().item() if fin.any() else 0.0


def _mlm_compact(out):
    return dict(head=out[:, :256], lse=torch.logsumexp(out.float(), 1), argmax=out.argmax(1), mean=out.float().mean(1))


@pytest.mark.parametrize("case", ["small_l2x1_b4", "full_ragged_b3", "full_b2"])
def test_pretrain_tasks_vs_reference_golden_and_oracle(case):
    from hamt_b200 import synth
    from oracle import hamt_oracle as O
    rec = torch.load(os.path.join(GOLD, f"pretrain_{case}.pt"))
    meta = rec["meta"]
    cfg, model, sd = _build(meta["cfg"], meta["weight_seed"])
    model.eval()
    report, failures = [], []
    for task in TASKS:
        b = synth.make_batch(task, seed=meta["batch_seed"], **meta["batch"])
        bd = _to_dev(b)
        for cl in (False, True):
            np.random.seed(meta["rng_seed"]); torch.manual_seed(meta["rng_seed"])
            with torch.no_grad():
                out = model(bd, task, compute_loss=cl)
            outs = list(out) if isinstance(out, tuple) else [out]
            np.random.seed(meta["rng_seed"]); torch.manual_seed(meta["rng_seed"])
            with torch.no_grad():
                o16 = O.pretrain_forward(sd, cfg, b, task, compute_loss=cl, rg=O.BF16)
            o16 = list(o16) if isinstance(o16, tuple) else [o16]
            gold = rec[f"{task}_{'loss' if cl else 'logits'}"]
            for i, (g_ours, g_o16, g_ref) in enumerate(zip(outs, o16, gold)):
                key = f"{case}/{task}/{'loss' if cl else 'logits'}[{i}]"
                if isinstance(g_ref, dict):     # compacted MLM logits: column slice + row statistics
                    g_ours, g_o16, g_ref = _mlm_compact(g_ours)["head"], _mlm_compact(g_o16)["head"], g_ref["head"]
                if g_ref.dtype in (torch.int64, torch.bool):
                    assert torch.equal(g_ours.cpu(), g_ref), key
                    continue
                e_ref, e_o16_ref, e_o16 = _err(g_ours, g_ref), _err(g_o16, g_ref), _err(g_ours, g_o16)
                report.append(f"{key}: ours_vs_reference_fp32={e_ref:.3e} bf16oracle_vs_reference_fp32={e_o16_ref:.3e} ours_vs_bf16oracle={e_o16:.3e}")
                tol = TOL_LOSS * max(1.0, g_ref[torch.isfinite(g_ref)].abs().max().item()) if cl else TOL_LOGITS
                if e_ref > tol:
                    failures.append(f"{key}: |ours - reference| = {e_ref:.3e} > {tol:.3e}")
                if not cl and e_ref > RATIO * e_o16_ref + 5e-3:
                    failures.append(f"{key}: ours is {e_ref:.3e} from fp32 but the bf16-regime oracle only {e_o16_ref:.3e}")
            if task == "sap" and not cl:
                ours_l, ref_l = outs[0].float().cpu(), gold[0]
                top2 = ref_l.topk(2, dim=1).values
                margin_ok = (top2[:, 0] - top2[:, 1]) > 2 * TOL_LOGITS
                report.append(f"{case}/sap argmax: ours={ours_l.argmax(1).tolist()} reference={ref_l.argmax(1).tolist()} checked={margin_ok.tolist()}")
                if not torch.equal(ours_l.argmax(1)[margin_ok], ref_l.argmax(1)[margin_ok]):
                    failures.append("SAP argmax actions differ from the reference")
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/parity_{case}.txt", "w") as fh:
        fh.write("\n".join(report) + "\n")
    assert not failures, "\n".join(failures)


def _proj_loss(out, seed):
    """Well-conditioned scalar for gradient checks: fixed random projection of every finite output element."""
    outs = list(out) if isinstance(out, tuple) else [out]
    total = 0.0
    for i, o in enumerate(outs):
        if not o.is_floating_point() or not o.requires_grad:
            continue
        w = torch.randn(o.shape, generator=torch.Generator().manual_seed(seed + i)).to(o.device)
        fin = torch.isfinite(o)
        total = total + (torch.where(fin, o.float(), torch.zeros_like(o, dtype=torch.float32)) * w).sum() / max(1, int(fin.sum()))
    return total


@pytest.mark.parametrize("mode", ["proj", "loss"])
@pytest.mark.parametrize("task", TASKS)
def test_pretrain_gradients_vs_oracle(task, mode):
    """fwd+bwd in train mode with dropout probabilities 0: every parameter gradient the kernels write into the arena
    against fp32 autograd through the oracle.  mode 'proj': random projection of the logits (well conditioned: per-tensor
    relative L2 error <= 10 %, measured 0.3 .. 7.5 %);  mode 'loss': the task's own loss .mean() (what the training loop
    differentiates; CE over near-identical candidates cancels heavily at random init, so the bound is norm ratio + rel 0.5)."""
    from hamt_b200 import synth
    from oracle import hamt_oracle as O
    cfg_over = dict(num_l_layers=2, num_x_layers=2, num_h_pano_layers=1)
    cfg, model, sd = _build(cfg_over, 3)
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    b = synth.make_batch(task, batch_size=4, txt_len=24, hist_len=5, seed=9, ragged=True)
    cl = mode == "loss"
    np.random.seed(1); torch.manual_seed(1)
    out = model(_to_dev(b), task, compute_loss=cl)
    loss = out.mean() if cl else _proj_loss(out, 100)
    loss.backward()
    torch.cuda.synchronize()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    if "mlm_head.predictions.decoder.weight" in sdr:
        sdr["mlm_head.predictions.decoder.weight"] = sdr["bert.embeddings.word_embeddings.weight"]
    np.random.seed(1); torch.manual_seed(1)
    ref_out = O.pretrain_forward(sdr, cfg, b, task, compute_loss=cl, rg=O.BF16)
    ref = ref_out.mean() if cl else _proj_loss(ref_out, 100)
    assert abs(float(loss) - float(ref)) < 1e-2 * max(1.0, abs(float(ref)))
    ref.backward()
    named = dict(model.named_parameters())
    bad, checked, rows = [], 0, []
    for k, p in named.items():
        g_ref = sdr[k].grad
        if g_ref is None or g_ref.abs().max().item() == 0:
            assert p.grad is None or p.grad.abs().max().item() < 1e-6, f"{k}: gradient where the reference has none"
            continue
        assert p.grad is not None, f"{k}: missing gradient"
        g = p.grad.float().cpu()
        # gradients that are mathematically zero (key bias under softmax, LN bias ahead of a shared 1-wide classifier
        # over a softmax group) are pure rounding noise in both implementations: floor the denominator
        floor = 1e-3 * (g_ref.numel() ** 0.5)
        gn, rn = g.norm().item(), g_ref.norm().item()
        rel = (g - g_ref).norm().item() / max(rn, floor)
        checked += 1
        rows.append((k, round(rel, 4), gn, rn))
        ok = rel <= 0.10 if mode == "proj" else (rel <= 0.5 and (rn < floor or abs(gn / rn - 1) < 0.15))
        if not ok:
            bad.append((k, round(rel, 4), gn, rn))
    assert checked > 20
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/grad_{task}_{mode}.txt", "w") as fh:
        fh.write("\n".join(map(str, sorted(rows, key=lambda t: -t[1]))))
    assert not bad, f"gradient mismatch ({len(bad)} of {checked}): {sorted(bad, key=lambda t: -t[1])[:8]}"


def test_grad_accumulation_and_zero_grad_semantics():
    """Two backward passes without zero_grad accumulate; zero_grad(set_to_none=True) restarts; parameters a task does not use
    keep grad None (the reference's AdamW skips them, optim/adamw.py:64-66)."""
    from hamt_b200 import synth
    cfg, model, sd = _build(dict(num_l_layers=1, num_x_layers=1, num_h_pano_layers=1), 3)
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    b = _to_dev(synth.make_batch("sap", batch_size=2, txt_len=16, hist_len=3, seed=2))
    w = model.bert.encoder.layer[0].output.dense.weight
    model(b, "sap").mean().backward()
    g1 = w.grad.clone()
    assert model.mlm_head.predictions.bias.grad is None and model.itm_head.net[0].weight.grad is None
    model(b, "sap").mean().backward()
    assert (w.grad - 2 * g1).abs().max().item() < 2e-2 * g1.abs().max().item() + 1e-6
    model.zero_grad(set_to_none=True)
    model(b, "sap").mean().backward()
    assert (w.grad - g1).abs().max().item() < 2e-2 * g1.abs().max().item() + 1e-6


def test_train_mode_dropout_is_active_and_reseeded():
    from hamt_b200 import synth
    cfg, model, sd = _build(dict(num_l_layers=1, num_x_layers=1, num_h_pano_layers=1), 3)
    b = _to_dev(synth.make_batch("sap", batch_size=2, txt_len=16, hist_len=3, seed=2))
    model.eval()
    with torch.no_grad():
        e1, e2 = model(b, "sap", False), model(b, "sap", False)
    assert torch.equal(e1, e2)
    model.train()
    with torch.no_grad():
        t1, t2 = model(b, "sap", False), model(b, "sap", False)
    fin = torch.isfinite(t1)
    assert not torch.equal(t1[fin], t2[fin]) and not torch.equal(t1[fin], e1[fin])


def test_hist_none_and_single_sample_edge_cases():
    """All samples at step 0: hist_*_fts are None (r2r_tasks.py:360-366); batch of one; ITM with batch 1 (vilmodel.py:690-691)."""
    from hamt_b200 import synth
    from oracle import hamt_oracle as O
    cfg, model, sd = _build(dict(num_l_layers=1, num_x_layers=1, num_h_pano_layers=1), 5)
    model.eval()
    for task, kw in (("sap", dict(batch_size=3, txt_len=20, hist_len=0)), ("mlm", dict(batch_size=1, txt_len=9, hist_len=2)),
                     ("itm", dict(batch_size=1, txt_len=12, hist_len=4)), ("sap", dict(batch_size=2, txt_len=80, hist_len=1, ragged=True))):
        b = synth.make_batch(task, seed=4, **kw)
        np.random.seed(3); torch.manual_seed(3)
        with torch.no_grad():
            out = model(_to_dev(b), task, compute_loss=False)
        np.random.seed(3); torch.manual_seed(3)
        with torch.no_grad():
            ref = O.pretrain_forward(sd, cfg, b, task, compute_loss=False)
        out = out[0] if isinstance(out, tuple) else out
        ref = ref[0] if isinstance(ref, tuple) else ref
        assert _err(out, ref) < TOL_LOGITS, (task, kw)


def test_finetune_navcmt_modes_vs_reference_golden():
    import hamt_b200  # noqa: F401
    from hamt_b200 import synth
    from hamt_b200.config import HamtConfig
    from hamt_b200.vilmodel_cmt import NavCMT
    rec = torch.load(os.path.join(GOLD, "finetune_navcmt.pt"))
    meta = rec["meta"]
    model = NavCMT(HamtConfig(**meta["cfg"]))
    model.load_state_dict(synth.seeded_state_dict(model, seed=meta["weight_seed"]))
    model = model.cuda().eval()
    B, L, O = meta["B"], meta["L"], meta["O"]
    b = _to_dev(synth.make_batch("sap", batch_size=B, txt_len=L, hist_len=2, n_ob=O, seed=meta["batch_seed"], ragged=True))
    HID = 8e-2    # hidden states reach |x| ~ 4 where one bf16 ulp is 3.1e-2
    with torch.no_grad():
        txt = model("language", txt_ids=b["txt_ids"], txt_masks=b["txt_masks"])
        assert _err(txt, rec["language"]) < HID
        h0 = model("history")
        assert tuple(h0.shape) == (1, 768)
        assert _err(h0, rec["history0"]) < HID
        hs = [h0.expand(B, -1)]
        for t in range(2):
            h = model("history", hist_img_feats=b["hist_img_fts"][:, t], hist_ang_feats=b["hist_ang_fts"][:, t],
                      ob_step_ids=torch.LongTensor([t]).cuda(), hist_pano_img_feats=b["hist_pano_img_fts"][:, t],
                      hist_pano_ang_feats=b["hist_pano_ang_fts"][:, t])
            assert _err(h, rec["history"][t]) < HID
            hs.append(h)
        hist = torch.stack(hs, 1)
        hm = torch.ones(B, 3, dtype=torch.bool, device="cuda")
        vis = model("visual", txt_embeds=txt, txt_masks=b["txt_masks"], hist_embeds=hist, hist_masks=hm, ob_img_feats=b["ob_img_fts"],
                    ob_ang_feats=b["ob_ang_fts"], ob_nav_types=b["ob_nav_types"], ob_masks=b["ob_masks"])
        assert _err(vis[0], rec["visual"][0]) < TOL_LOGITS
        for got, want in zip(vis[1:], rec["visual"][1:]):
            assert _err(got, want) < HID
        assert torch.equal(vis[0].float().cpu().argmax(1), rec["visual"][0].argmax(1))


STRESS = {
    # BASELINE config 4 (RxR): text_len 300, multilingual-sized position table, 512-d features, hist 20 x 36; the XLM-R vocabulary is
    # cut to 6000 rows so that the CPU oracle stays fast (the embedding gather / tied decoder are size-agnostic)
    "rxr": (dict(image_feat_size=512, max_position_embeddings=514, vocab_size=6000, num_l_layers=2, num_x_layers=2, num_h_pano_layers=1),
            dict(batch_size=2, txt_len=300, hist_len=20, feat=512, vocab_hi=6000)),
    # BASELINE config 5 (R4R): hist_len 40 x 36 views (41 + 37 = 78 vision tokens)
    "r4r": (dict(num_l_layers=1, num_x_layers=2, num_h_pano_layers=2), dict(batch_size=2, txt_len=80, hist_len=40)),
}


@pytest.mark.parametrize("name", ["rxr", "r4r"])
def test_stress_configs_forward_and_gradients_vs_oracle(name):
    """BASELINE configs 4 and 5 as parity cases: eval logits (SAP, MLM) against the fp32 oracle, then train-mode (dropout 0)
    parameter gradients of a projected SAP output against autograd through the bf16-regime oracle.  RxR exercises the
    long-sequence attention backward (S = 300 > 128)."""
    from hamt_b200 import synth
    from oracle import hamt_oracle as O
    cfg_over, bkw = STRESS[name]
    cfg, model, sd = _build(cfg_over, 11)
    model.eval()
    for task in ("sap", "mlm"):
        b = synth.make_batch(task, seed=21, ragged=True, **bkw)
        with torch.no_grad():
            out = model(_to_dev(b), task, compute_loss=False)
            ref = O.pretrain_forward(sd, cfg, b, task, compute_loss=False)
        assert _err(out, ref) < TOL_LOGITS, (name, task, _err(out, ref))
        if task == "sap":
            top2 = ref.topk(2, dim=1).values
            ok = (top2[:, 0] - top2[:, 1]) > 2 * TOL_LOGITS
            assert torch.equal(out.float().cpu().argmax(1)[ok], ref.argmax(1)[ok])
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    b = synth.make_batch("sap", seed=22, ragged=True, **bkw)
    loss = _proj_loss(model(_to_dev(b), "sap", compute_loss=False), 100)
    loss.backward()
    torch.cuda.synchronize()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    sdr["mlm_head.predictions.decoder.weight"] = sdr["bert.embeddings.word_embeddings.weight"]
    ref = _proj_loss(O.pretrain_forward(sdr, cfg, b, "sap", compute_loss=False, rg=O.BF16), 100)
    assert abs(float(loss.detach()) - float(ref.detach())) < 1e-2 * max(1.0, abs(float(ref.detach())))
    ref.backward()
    bad, checked = [], 0
    for k, p in model.named_parameters():
        g_ref = sdr[k].grad
        if g_ref is None or g_ref.abs().max().item() == 0:
            continue
        assert p.grad is not None, f"{k}: missing gradient"
        floor = 1e-3 * (g_ref.numel() ** 0.5)
        rel = (p.grad.float().cpu() - g_ref).norm().item() / max(g_ref.norm().item(), floor)
        checked += 1
        if rel > 0.10:
            bad.append((k, round(rel, 4)))
    assert checked > 20
    assert not bad, f"{name}: gradient mismatch ({len(bad)} of {checked}): {sorted(bad, key=lambda t: -t[1])[:8]}"


@pytest.mark.parametrize("task", TASKS)
def test_kernels_never_write_outside_a_parameters_gradient(task):
    """Arena hygiene: every parameter's gradient view is followed by an alignment tail (arena.ALIGN = 64 elements); the wgrad /
    column-sum / embedding kernels must leave those tails and the gradients of untouched parameters at exactly zero."""
    from hamt_b200 import synth
    cfg, model, sd = _build(dict(num_l_layers=1, num_x_layers=1, num_h_pano_layers=1), 3)
    model.train()
    b = synth.make_batch(task, batch_size=3, txt_len=20, hist_len=4, seed=5, ragged=True)
    np.random.seed(1); torch.manual_seed(1)
    model(_to_dev(b), task, compute_loss=True).mean().backward()
    torch.cuda.synchronize()
    arena = model.arena()
    inside = torch.zeros_like(arena.flat_grad, dtype=torch.bool)
    names = {id(p): n for n, p in model.named_parameters()}
    for p in arena.params:
        o = arena.offsets[id(p)]
        if p.grad is not None:
            inside[o:o + p.numel()] = True
    stray = (arena.flat_grad != 0) & ~inside
    if stray.any():
        idx = int(stray.nonzero()[0])
        owner = max((arena.offsets[id(p)], names.get(id(p), "?")) for p in arena.params if arena.offsets[id(p)] <= idx)
        raise AssertionError(f"{int(stray.sum())} stray gradient elements outside touched parameter views; first at flat index {idx} "
                             f"(after the start of {owner[1]} by {idx - owner[0]})")
