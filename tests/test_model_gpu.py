"""Model-level parity of the sm_100a path, called through the reference-facing module API (which goes through the C ABI).

Comparators
  (a) golden vectors produced by the UNMODIFIED reference in fp32 (tests/golden, oracle/make_golden.py);
  (b) the CPU oracle run in the same dtype regime (bf16 rounding points mirrored, oracle.BF16).

What can and cannot be asserted (measured, see DESIGN.md "Parity"): a 13..15-layer bf16 pipeline is chaotic at the rounding
level -- perturbing the weights by 1e-7 relative moves the bf16-regime logits by ~1.8e-2, the same size as the bf16-vs-fp32 gap
(tests/test_oracle.py::test_bf16_regime_rounding_chaos).  The reference's OWN torch.autocast(bf16) run differs from its fp32 run by
0.6e-2 .. 1.7e-2 on these inputs; those outputs are stored next to the fp32 ones in the goldens (`*_autocast`) and are the yardstick:
   * logits:  |ours - reference_fp32| <= RATIO (1.5) x |reference_autocast - reference_fp32|, per task and case (max-abs), and
              <= TOL_LOGITS absolutely;
   * losses:  |ours - reference_fp32| <= TOL_LOSS * max(1, |ref|);
   * SAP / finetune action argmax bit-exact on every row whose reference top-2 margin exceeds twice the measured error;
   * -inf patterns (masked_fill_(nav_type == 0)) identical;  integer outputs bit-exact.
"< 1e-3 in bf16" is asserted where it is attainable: per kernel (tests/test_*_gpu.py).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TASKS = ("mlm", "sap", "sar", "sprel", "mrc", "itm")
TOL_LOGITS = 3e-2     # absolute ceiling, max-abs, logits (std 0.3 .. 0.55) vs the fp32 reference
TOL_LOSS = 1e-1       # un-reduced loss entries (MSE on angles amplifies a logit error by 2|pred - target| <= 2 pi)
RATIO = 1.5           # ours may be at most this many times further from fp32 than the reference's own torch.autocast(bf16) run


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


def _comparable(task, out, gold):
    """Bring an output into the (possibly compacted) form its golden counterpart was stored in (oracle/make_golden.py compact())."""
    if isinstance(gold, dict):                  # MLM logits: column slice + row statistics
        return out[:, :gold["head"].shape[1]], gold["head"]
    if out.dim() == 2 and gold.dim() == 2 and out.shape[1] > gold.shape[1]:      # batch-64 MRC: 64 columns
        return out[:, :gold.shape[1]], gold
    return out, gold


def _batch_kwargs(task, bkw):
    return dict(bkw, batch_size=bkw["batch_size"] // 2) if (bkw["batch_size"] >= 64 and task == "itm") else bkw


def _check_case_against_goldens(case, rec, run_model, report, failures, with_oracle=None):
    """Shared by the eval-mode cases: every task, logits and un-reduced losses, against the fp32 goldens with the reference's own
    autocast(bf16) distance as the yardstick."""
    from hamt_b200 import synth
    meta = rec["meta"]
    for task in TASKS:
        b = synth.make_batch(task, seed=meta["batch_seed"], **_batch_kwargs(task, meta["batch"]))
        for cl in (False, True):
            np.random.seed(meta["rng_seed"]); torch.manual_seed(meta["rng_seed"])
            out = run_model(b, task, cl)
            outs = list(out) if isinstance(out, tuple) else [out]
            o16 = with_oracle(b, task, cl) if with_oracle is not None else None
            kind = "loss" if cl else "logits"
            gold, gold_ac = rec[f"{task}_{kind}"], rec[f"{task}_{kind}_autocast"]
            for i, (g_ours, g_ref, g_ac) in enumerate(zip(outs, gold, gold_ac)):
                key = f"{case}/{task}/{kind}[{i}]"
                g_ours, g_ref_t = _comparable(task, g_ours, g_ref)
                g_ac = g_ac["head"] if isinstance(g_ac, dict) else g_ac
                if g_ref_t.dtype in (torch.int64, torch.bool):
                    assert torch.equal(g_ours.cpu(), g_ref_t), key
                    continue
                e_ref, e_ac = _err(g_ours, g_ref_t), _err(g_ac, g_ref_t)
                line = f"{key}: ours_vs_reference_fp32={e_ref:.3e} reference_autocast_vs_fp32={e_ac:.3e} ratio={e_ref / max(e_ac, 1e-12):.2f}"
                if o16 is not None:
                    g_o = _comparable(task, o16[i], g_ref)[0]
                    line += f" bf16oracle_vs_reference_fp32={_err(g_o, g_ref_t):.3e} ours_vs_bf16oracle={_err(g_ours, g_o):.3e}"
                report.append(line)
                if cl:
                    tol = TOL_LOSS * max(1.0, g_ref_t[torch.isfinite(g_ref_t)].abs().max().item())
                    if e_ref > tol:
                        failures.append(f"{key}: |ours - reference| = {e_ref:.3e} > {tol:.3e}")
                else:
                    if e_ref > TOL_LOGITS:
                        failures.append(f"{key}: |ours - reference| = {e_ref:.3e} > {TOL_LOGITS:.1e}")
                    if e_ref > RATIO * e_ac + 1e-6:
                        failures.append(f"{key}: ours is {e_ref:.3e} from the fp32 reference, the reference's own autocast(bf16) run {e_ac:.3e} "
                                        f"(ratio {e_ref / max(e_ac, 1e-12):.2f} > {RATIO})")
            if task == "sap" and not cl:
                ours_l, ref_l = outs[0].float().cpu(), gold[0]
                e_sap = _err(ours_l, ref_l)
                top2 = ref_l.topk(2, dim=1).values
                margin_ok = (top2[:, 0] - top2[:, 1]) > 2 * e_sap
                report.append(f"{case}/sap argmax: rows={ref_l.shape[0]} checked={int(margin_ok.sum())} (margin > 2 x {e_sap:.3e}) "
                              f"equal={int((ours_l.argmax(1) == ref_l.argmax(1)).sum())}")
                if not torch.equal(ours_l.argmax(1)[margin_ok], ref_l.argmax(1)[margin_ok]):
                    failures.append("SAP argmax actions differ from the reference")
                if int(margin_ok.sum()) < (ref_l.shape[0] * 3) // 4:
                    failures.append(f"SAP argmax: only {int(margin_ok.sum())} of {ref_l.shape[0]} rows have a margin above twice the measured error")


@pytest.mark.parametrize("case", ["small_l2x1_b4", "full_ragged_b3", "full_b2"])
def test_pretrain_tasks_vs_reference_golden_and_oracle(case):
    from oracle import hamt_oracle as O
    rec = torch.load(os.path.join(GOLD, f"pretrain_{case}.pt"))
    meta = rec["meta"]
    cfg, model, sd = _build(meta["cfg"], meta["weight_seed"])
    model.eval()

    def run_model(b, task, cl):
        with torch.no_grad():
            return model(_to_dev(b), task, compute_loss=cl)

    def run_oracle(b, task, cl):
        np.random.seed(meta["rng_seed"]); torch.manual_seed(meta["rng_seed"])
        with torch.no_grad():
            o = O.pretrain_forward(sd, cfg, b, task, compute_loss=cl, rg=O.BF16)
        return list(o) if isinstance(o, tuple) else [o]

    report, failures = [], []
    _check_case_against_goldens(case, rec, run_model, report, failures, with_oracle=run_oracle)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/parity_{case}.txt", "w") as fh:
        fh.write("\n".join(report) + "\n")
    assert not failures, "\n".join(failures)


def test_headline_batch64_vs_reference_golden_eager_and_graphed():
    """BASELINE configs[1] shape (batch 64, ITM 32, txt 80, hist 15 x 36, obs 37, full depth): the cost model picks the CTA-pair /
    split-K tiles the bench runs.  (1) eval-mode logits + losses of all six tasks against the goldens of the UNMODIFIED reference,
    SAP argmax on all 64 rows with a sufficient margin; (2) the SAME captured-graph path the bench times (graph.GraphedTrainer,
    train mode, dropout probabilities 0): the loss vector of every task against the reference's fp32 loss."""
    from hamt_b200 import graph, synth
    rec = torch.load(os.path.join(GOLD, "pretrain_full_b64.pt"))
    meta = rec["meta"]
    cfg, model, sd = _build(meta["cfg"], meta["weight_seed"])
    model.eval()

    def run_model(b, task, cl):
        with torch.no_grad():
            return model(_to_dev(b), task, compute_loss=cl)

    report, failures = [], []
    _check_case_against_goldens("full_b64", rec, run_model, report, failures)
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    trainer = graph.GraphedTrainer(model)
    for task in TASKS:
        b = synth.make_batch(task, seed=meta["batch_seed"], **_batch_kwargs(task, meta["batch"]))
        np.random.seed(meta["rng_seed"]); torch.manual_seed(meta["rng_seed"])
        bd = graph.add_sync_free_extras(task, b)
        loss = trainer.step(task, _to_dev(bd)).detach().float().cpu()
        torch.cuda.synchronize()
        g_ref = rec[f"{task}_loss"][0]
        e = _err(loss, g_ref)
        tol = TOL_LOSS * max(1.0, g_ref.abs().max().item())
        report.append(f"full_b64/{task}/graphed_loss: ours_vs_reference_fp32={e:.3e} (tol {tol:.2e}); mean ours={loss.mean():.5f} reference={g_ref.mean():.5f}")
        if e > tol:
            failures.append(f"full_b64/{task}: graphed loss differs from the reference by {e:.3e} > {tol:.3e}")
        g = model.bert.encoder.x_layers[0].visual_attention.att.query.weight.grad
        if g is None or not torch.isfinite(g).all() or float(g.abs().sum()) == 0.0:
            failures.append(f"full_b64/{task}: the captured step left no finite gradient")
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_full_b64.txt", "w") as fh:
        fh.write("\n".join(report) + "\n")
    assert not failures, "\n".join(failures)


def _proj_loss(out, seed, positive=False):
    """Well-conditioned scalar for gradient checks: fixed random projection of every finite output element.
    positive: |w| instead of w.  The five ITM logits of a sample are near-identical functions of the text stream and of the shared
    history CLS token (same instruction, slightly different trajectories), so a random-SIGN combination of them is a small difference
    of large terms: the fp32 oracle's own gradient moves by 7.5 % when only its forward residual stream is rounded to bf16
    (measured) -- an ill-conditioned comparator, not a property of the kernels.  Equal signs keep the sum
    well conditioned; the cancelling case stays covered by mode 'loss' (cross-entropy over the same five logits)."""
    outs = list(out) if isinstance(out, tuple) else [out]
    total = 0.0
    for i, o in enumerate(outs):
        if not o.is_floating_point() or not o.requires_grad:
            continue
        w = torch.randn(o.shape, generator=torch.Generator().manual_seed(seed + i)).to(o.device)
        if positive:
            w = w.abs()
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
    loss = out.mean() if cl else _proj_loss(out, 100, positive=task == "itm")
    loss.backward()
    torch.cuda.synchronize()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    if "mlm_head.predictions.decoder.weight" in sdr:
        sdr["mlm_head.predictions.decoder.weight"] = sdr["bert.embeddings.word_embeddings.weight"]
    np.random.seed(1); torch.manual_seed(1)
    ref_out = O.pretrain_forward(sdr, cfg, b, task, compute_loss=cl, rg=O.BF16)
    ref = ref_out.mean() if cl else _proj_loss(ref_out, 100, positive=task == "itm")
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
