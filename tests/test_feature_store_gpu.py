"""Device-resident feature store (feature_store.py -> C ABI hamt_gather_rows_pad_bf16) against the reference's own
per-sample assembly + padding (golden file produced from the unmodified reference code, tests/golden/feature_assembly.pt).
Row gathers are copies: bit-exact against the bf16 rounding of the reference's fp32 rows; angle features bit-exact in fp32."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _store_and_indices():
    import hamt_b200  # noqa: F401
    from hamt_b200.feature_store import FeatureStore
    from oracle import make_golden_features as G
    keys, feats = G.scenario()
    store = FeatureStore(keys, torch.from_numpy(feats), "cuda", image_feat_size=G.D, angle_feat_size=G.A)
    B, T = len(G.SAMPLES), max(s["t_cur"] for s in G.SAMPLES)
    hp = torch.full((B, T), -1, dtype=torch.int64)
    hv = torch.full((B, T), -1, dtype=torch.int64)
    op, ov = torch.zeros(B, dtype=torch.int64), torch.zeros(B, dtype=torch.int64)
    for b, s in enumerate(G.SAMPLES):
        for t in range(s["t_cur"]):
            hp[b, t], hv[b, t] = store.lookup(G.SCAN, s["path"][t]), s["views"][t]
        op[b], ov[b] = store.lookup(G.SCAN, s["path"][s["t_cur"]]), s["views"][s["t_cur"]]
    return store, hp, hv, op, ov


def test_history_and_observation_assembly_match_reference():
    rec = torch.load(os.path.join(GOLD, "feature_assembly.pt"))
    store, hp, hv, op, ov = _store_and_indices()
    h = store.assemble_history(hp, hv, with_pano=True, with_probs=True)
    o = store.assemble_observation(op, ov)
    torch.cuda.synchronize()
    for k in ("hist_img_fts", "hist_pano_img_fts"):
        assert h[k].dtype == torch.bfloat16 and torch.equal(h[k].cpu(), rec[k].to(torch.bfloat16)), k
    assert torch.equal(h["hist_pano_ang_fts"].cpu(), rec["hist_pano_ang_fts"])
    assert (h["hist_img_probs"].cpu() - rec["hist_img_probs"]).abs().max().item() < 1e-6
    assert torch.equal(o["ob_img_fts"].cpu(), rec["ob_img_fts"].to(torch.bfloat16))
    assert torch.equal(o["ob_ang_fts"].cpu(), rec["ob_ang_fts"])
    assert float(o["ob_img_fts"][:, -1].abs().max()) == 0.0 and float(o["ob_ang_fts"][:, -1].abs().max()) == 0.0      # STOP row


def test_edge_cases_empty_history_killed_observation_bad_index():
    store, hp, hv, op, ov = _store_and_indices()
    e = store.assemble_history(hp[:, :0], hv[:, :0])
    assert all(v is None for v in e.values())                                  # all samples at step 0 (r2r_tasks.py:360-366)
    op2, ov2 = op.clone(), ov.clone()
    op2[1], ov2[2] = -1, -1                                                    # random_kill_v / random_kill_a (r2r_tasks.py:322-327)
    o = store.assemble_observation(op2, ov2)
    assert float(o["ob_img_fts"][1].abs().max()) == 0.0 and float(o["ob_img_fts"][0].abs().max()) > 0.0
    assert float(o["ob_ang_fts"][2].abs().max()) == 0.0 and float(o["ob_ang_fts"][1].abs().max()) > 0.0
    bad = hp.clone()
    bad[0, 0] = store.n_pano
    with pytest.raises(IndexError):
        store.assemble_history(bad, hv)
    with pytest.raises(KeyError):
        store.lookup("scanA", "nope")


def test_model_consumes_store_output_like_host_features():
    """The backbone takes the store's bf16 tensors as they are: logits are bitwise identical to feeding the same rows as fp32 host
    features (which the model casts to bf16 itself), at the real feature width."""
    import hamt_b200  # noqa: F401
    from hamt_b200 import synth
    from hamt_b200.config import HamtConfig
    from hamt_b200.feature_store import FeatureStore
    from hamt_b200.pretrain_cmt import MultiStepNavCMTPreTraining
    g = torch.Generator().manual_seed(3)
    V, B, T = 9, 3, 4
    feats = torch.randn(V, 36, 768, generator=g)
    store = FeatureStore([f"s_{i}" for i in range(V)], feats, "cuda", keep_logits=False)
    hp = torch.randint(0, V, (B, T), generator=g)
    hv = torch.randint(0, 36, (B, T), generator=g)
    lens = torch.tensor([4, 2, 1])
    pad = torch.arange(T)[None] >= lens[:, None]
    hp[pad], hv[pad] = -1, -1
    op, ov = torch.randint(0, V, (B,), generator=g), torch.randint(0, 36, (B,), generator=g)
    model = MultiStepNavCMTPreTraining(HamtConfig(num_l_layers=1, num_x_layers=1, num_h_pano_layers=1))
    model.load_state_dict(synth.seeded_state_dict(model, seed=2))
    model = model.cuda().eval()
    b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synth.make_batch("sap", batch_size=B, txt_len=12, hist_len=T, seed=1).items()}
    b["hist_masks"] = (torch.arange(T + 1)[None] < (lens + 1)[:, None]).cuda()
    dev = dict(b)
    dev.update(store.assemble_history(hp, hv))
    dev.update(store.assemble_observation(op, ov))
    host = dict(b)
    valid = (~pad).float()
    host["hist_img_fts"] = (feats[hp.clamp(min=0), hv.clamp(min=0)] * valid[..., None]).cuda()
    host["hist_pano_img_fts"] = (feats[hp.clamp(min=0)] * valid[..., None, None]).cuda()
    host["hist_pano_ang_fts"] = dev["hist_pano_ang_fts"]
    host["ob_img_fts"] = torch.cat([feats[op], torch.zeros(B, 1, 768)], 1).cuda()
    host["ob_ang_fts"] = dev["ob_ang_fts"]
    with torch.no_grad():
        a = model(dev, "sap", compute_loss=False)
        c = model(host, "sap", compute_loss=False)
    fin = torch.isfinite(c)
    assert torch.equal(torch.isfinite(a), fin) and torch.equal(a[fin], c[fin])
