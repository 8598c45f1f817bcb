"""CUDA-graph step (graph.py): a replayed captured step must produce the same loss and the same gradients as the eager
path on the same batch (dropout off), must follow new inputs copied into its static buffers, and must re-seed dropout
on every replay."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build():
    import hamt_b200  # noqa: F401
    from hamt_b200 import synth
    from hamt_b200.config import HamtConfig
    from hamt_b200.pretrain_cmt import MultiStepNavCMTPreTraining
    model = MultiStepNavCMTPreTraining(HamtConfig(num_l_layers=1, num_x_layers=1, num_h_pano_layers=1))
    model.load_state_dict(synth.seeded_state_dict(model, seed=3))
    return model.cuda().train()


@pytest.mark.parametrize("task", ["sap", "mlm", "mrc", "itm", "sprel"])
def test_graph_replay_matches_eager(task):
    from hamt_b200 import graph, synth
    model = _build()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    kw = dict(batch_size=4, txt_len=24, hist_len=5)
    b1 = synth.make_batch(task, seed=1, **kw)
    b2 = synth.make_batch(task, seed=2, **kw)
    if task in ("mlm", "mrc"):      # same number of masked rows -> same graph signature
        key = "txt_labels" if task == "mlm" else "hist_mrc_masks"
        b2[key] = b1[key].clone()
        if task == "mlm":
            b2["txt_ids"] = b1["txt_ids"].clone()

    def eager(b, seed):
        np.random.seed(seed); torch.manual_seed(seed)
        model.zero_grad(set_to_none=True)
        bd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
        loss = model(bd, task, compute_loss=True)
        loss.mean().backward()
        torch.cuda.synchronize()
        return loss.detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

    l1, g1 = eager(b1, 11)
    l2, g2 = eager(b2, 12)
    np.random.seed(11); torch.manual_seed(11)
    trainer = graph.GraphedTrainer(model)
    e1 = graph.add_sync_free_extras(task, b1)
    lg1 = trainer.step(task, e1).detach().clone()
    torch.cuda.synchronize()
    gg1 = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    assert torch.allclose(lg1, l1, atol=2e-3, rtol=1e-3)
    assert set(gg1) == set(g1)
    for n in g1:
        assert (gg1[n] - g1[n]).abs().max().item() <= 2e-2 * g1[n].abs().max().item() + 1e-5, n
    # new inputs through the same graph
    np.random.seed(12); torch.manual_seed(12)
    lg2 = trainer.step(task, graph.add_sync_free_extras(task, b2)).detach().clone()
    torch.cuda.synchronize()
    assert len(trainer.steps) == 1
    assert torch.allclose(lg2, l2, atol=2e-3, rtol=1e-3)
    gg2 = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    for n in g2:
        assert (gg2[n] - g2[n]).abs().max().item() <= 2e-2 * g2[n].abs().max().item() + 1e-5, n
    # parameters the task does not use keep grad None
    unused = model.itm_head.net[0].weight if task != "itm" else model.next_action.net[0].weight
    assert unused.grad is None


def test_graph_replay_reseeds_dropout():
    from hamt_b200 import graph, synth
    model = _build()
    b = graph.add_sync_free_extras("sap", synth.make_batch("sap", batch_size=2, txt_len=16, hist_len=3, seed=4))
    trainer = graph.GraphedTrainer(model)
    a = trainer.step("sap", b).detach().clone()
    c = trainer.step("sap", b).detach().clone()
    assert not torch.equal(a, c)


@pytest.mark.parametrize("task", ["sap", "itm"])
def test_packed_prefetch_matches_dict_batch(task):
    """loader.PackedBatch (one pinned blob, one H2D copy, one D2D copy into the captured step's static inputs) feeds the
    step exactly what the per-tensor path feeds it; loader.LossReader returns each step's loss in order."""
    from hamt_b200 import graph, loader, synth
    model = _build()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    kw = dict(batch_size=4, txt_len=24, hist_len=5)
    np.random.seed(5); torch.manual_seed(5)
    b1 = graph.add_sync_free_extras(task, synth.make_batch(task, seed=1, **kw))
    np.random.seed(6); torch.manual_seed(6)
    b2 = graph.add_sync_free_extras(task, synth.make_batch(task, seed=2, **kw))
    trainer = graph.GraphedTrainer(model)
    want = [trainer.step(task, b).detach().float().mean().item() for b in (b1, b2)]
    pk = loader.PackedBatch(b1, torch.device("cuda"))
    copy_stream = torch.cuda.Stream()
    reader = loader.LossReader(depth=2)
    got = []
    for b in (b1, b2):
        pk.fill(b)
        db = pk.to_device(copy_stream)
        torch.cuda.current_stream().wait_event(pk.ready)
        reader.push(trainer.step(task, db).float().mean())
        torch.cuda.synchronize()        # the single slot is refilled next iteration
    while reader.pending():
        got.append(reader.pop())
    assert len(trainer.steps) == 1, "the packed batch must hit the same captured graph"
    assert np.allclose(got, want, rtol=0, atol=1e-6), (got, want)


def test_graph_accumulate_and_capture_during_accumulation():
    """gradient_accumulation_steps > 1 (main_r2r.py:243-249): `step(..., accumulate=True)` adds to the gradients instead of
    resetting them, and a first-use capture in the middle of an accumulation (its warm-up passes run the model) leaves the
    already accumulated gradients intact."""
    from hamt_b200 import graph, synth
    model = _build()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    b = synth.make_batch("sap", batch_size=4, txt_len=24, hist_len=5, seed=1)
    bd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    model.zero_grad(set_to_none=True)
    model(bd, "sap", compute_loss=True).mean().backward()          # micro-batch 1, eager
    torch.cuda.synchronize()
    g1 = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    seed_before = int(model.arena().seed.item())
    trainer = graph.GraphedTrainer(model)
    trainer.step("sap", graph.add_sync_free_extras("sap", b), accumulate=True)     # micro-batch 2: captured on first use
    torch.cuda.synchronize()
    g2 = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    assert set(g1) == set(g2)
    for n in g1:
        assert (g2[n] - 2 * g1[n]).abs().max().item() <= 3e-2 * g1[n].abs().max().item() + 1e-5, n
    assert int(model.arena().seed.item()) != seed_before            # exactly the replay advanced it (warm-up passes did not) ...
    trainer.step("sap", graph.add_sync_free_extras("sap", b), accumulate=True)
    torch.cuda.synchronize()
    for n, p in model.named_parameters():
        if p.grad is not None:
            assert (p.grad - 3 * g1[n]).abs().max().item() <= 4e-2 * g1[n].abs().max().item() + 1e-5, n
    trainer.step("sap", graph.add_sync_free_extras("sap", b))       # default: a fresh accumulation
    torch.cuda.synchronize()
    for n, p in model.named_parameters():
        if p.grad is not None:
            assert (p.grad - g1[n]).abs().max().item() <= 2e-2 * g1[n].abs().max().item() + 1e-5, n


def test_run_to_run_gradient_drift_is_bounded():
    """The weight-gradient split-K epilogue, the LayerNorm / embedding / bias column sums and the attention bias sums accumulate with
    floating-point atomics (`red.global.add`), and the weight-gradient GEMMs run on a second stream: the summation ORDER varies from run
    to run, the values may not.  Two eager passes and two graph replays over the same batch and weights (dropout off) must agree to
    fp32-reassociation level: per-tensor relative L2 drift <= 2e-5 (bf16-level differences would be 4e-3), losses bit-equal."""
    from hamt_b200 import graph, synth
    model = _build()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    b = synth.make_batch("sap", seed=5, batch_size=8, txt_len=40, hist_len=6)
    bd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}

    def eager():
        model.zero_grad(set_to_none=True)
        loss = model(bd, "sap", compute_loss=True)
        loss.mean().backward()
        torch.cuda.synchronize()
        return loss.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

    l1, g1 = eager()
    l2, g2 = eager()
    trainer = graph.GraphedTrainer(model)
    model.zero_grad(set_to_none=True)
    l3 = trainer.step("sap", bd).detach().clone()
    torch.cuda.synchronize()
    g3 = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    l4 = trainer.step("sap", bd).detach().clone()
    torch.cuda.synchronize()
    g4 = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    assert torch.equal(l1, l2) and torch.equal(l3, l4) and torch.equal(l1, l3)
    assert g1.keys() == g2.keys() == g3.keys() == g4.keys()
    # (gradients that are identically zero in exact arithmetic -- the key bias of every attention: rows of dS sum to zero -- are pure
    # rounding noise, so the drift is measured against the larger of the tensor's own norm and 1e-3 of the typical gradient norm)
    norms = sorted(g1[k].float().norm().item() for k in g1)
    floor = 1e-3 * norms[len(norms) // 2]
    report = []
    for k in g1:
        if k.endswith("key.bias") or k == "next_action.net.4.bias":     # identically zero: rows of dS / of (softmax - onehot) sum to zero
            continue
        n = max(g1[k].float().norm().item(), floor)
        d = max((other[k].float() - g1[k].float()).norm().item() / n for other in (g2, g3, g4))
        report.append((d, k))
    report.sort(reverse=True)
    assert report[0][0] <= 2e-5, report[:6]


def test_itm_device_negative_sampling_distribution_and_graph_capture():
    """config.itm_device_negatives: negatives and position shuffles drawn on the device with the distribution of the
    reference's host loops (vilmodel.py:676-704): a negative is never the sample itself and is uniform over the others; a shuffle is a
    permutation of the valid history steps followed by the padding positions in order.  Inside a captured step the draws are fresh on
    every replay (no host plan in the batch)."""
    import hamt_b200  # noqa: F401
    from hamt_b200 import graph, synth
    from hamt_b200.config import HamtConfig
    from hamt_b200.pretrain_cmt import MultiStepNavCMTPreTraining
    from hamt_b200.vilmodel import itm_negative_plan_device
    torch.manual_seed(0)
    B, T = 6, 9
    lens = torch.tensor([9, 1, 4, 7, 2, 9])
    hm = (torch.arange(T + 1)[None] < (lens + 1)[:, None]).cuda()
    counts = torch.zeros(B, B)
    for _ in range(200):
        neg, shuf = itm_negative_plan_device(B, hm, T, 4)
        assert neg.shape == (B, 2) and len(shuf) == 2
        assert (neg != torch.arange(B, device="cuda")[:, None]).all() and neg.min() >= 0 and neg.max() < B
        for j in range(2):
            counts[torch.arange(B), neg[:, j].cpu()] += 1
        for s in shuf:
            s = s.cpu()
            for i in range(B):
                n = int(lens[i])
                assert sorted(s[i, :n].tolist()) == list(range(n)) and s[i, n:].tolist() == list(range(n, T))
    off = counts[~torch.eye(B, dtype=torch.bool)]
    assert counts.diag().sum() == 0 and off.min() > 0.5 * off.mean() and off.max() < 1.6 * off.mean()      # 400 draws per row over 5 others
    # captured ITM step without a host plan: replays draw new negatives
    model = MultiStepNavCMTPreTraining(HamtConfig(num_l_layers=1, num_x_layers=1, num_h_pano_layers=1, itm_device_negatives=True))
    model.load_state_dict(synth.seeded_state_dict(model, seed=3))
    model = model.cuda().train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    b = synth.make_batch("itm", seed=1, batch_size=4, txt_len=24, hist_len=5)
    bd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    bd["_itm_device_negatives"] = True
    trainer = graph.GraphedTrainer(model)
    e = graph.add_sync_free_extras("itm", bd)
    assert "itm_plan" not in e
    losses = [trainer.step("itm", e).detach().float().clone() for _ in range(4)]
    torch.cuda.synchronize()
    assert len(trainer.steps) == 1 and all(torch.isfinite(l).all() for l in losses)
    assert any(not torch.equal(losses[0], l) for l in losses[1:])
