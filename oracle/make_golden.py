"""Generate the golden vectors in tests/golden/ from the UNMODIFIED reference (run in the build
container where /root/reference exists):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

For each proxy task (and the finetune NavCMT modes) the reference model is built from
pretrain_src/config/r2r_model_config.json, loaded with the deterministic ``seeded_state_dict``
(vln-hamt_b200/synth.py), run in eval mode on ``make_batch`` inputs, and its outputs are stored --
once in fp32 (the reference's own arithmetic) and once under ``torch.autocast(bfloat16)``
(``*_autocast`` keys): the distance between the two is the reference's OWN bf16 error on these
inputs and is the yardstick of the model-level parity tests (|ours - fp32| <= 1.5 x |autocast - fp32|).
The ``full_b64`` case is the headline benchmark shape (batch 64, ITM 32, txt 80, hist 15 x 36, obs 37).
Everything is regenerated from seeds on the test machine, so the fixtures only hold outputs (a few
hundred KB).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
import hamt_b200  # noqa: E402,F401
from hamt_b200 import synth  # noqa: E402
from hamt_b200.config import HamtConfig  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TASKS = ("mlm", "sap", "sar", "sprel", "mrc", "itm")
# (name, config overrides, batch kwargs)
CASES = [
    ("full_b2", dict(), dict(batch_size=2, txt_len=80, hist_len=15)),
    ("full_ragged_b3", dict(), dict(batch_size=3, txt_len=40, hist_len=6, ragged=True)),
    ("small_l2x1_b4", dict(num_l_layers=2, num_x_layers=1, num_h_pano_layers=1), dict(batch_size=4, txt_len=24, hist_len=5, ragged=True)),
    ("full_b64", dict(), dict(batch_size=64, txt_len=80, hist_len=15)),
]
WEIGHT_SEED, BATCH_SEED, RNG_SEED = 11, 7, 5


def compact(task, out, big=False):
    """Keep fixtures small: the MLM logits are stored as a column slice + row statistics (64 columns in the batch-64 case,
    whose MRC logits / targets are also cut to 64 columns)."""
    if task == "mlm" and out.dim() == 2 and out.shape[1] > 4096:
        out = out.float()
        return dict(head=out[:, :64 if big else 256].clone(), lse=torch.logsumexp(out, 1), argmax=out.argmax(1), mean=out.mean(1))
    if big and task == "mrc" and out.dim() == 2 and out.shape[1] > 64:
        return out[:, :64].float().clone()
    return out.float().clone() if out.is_floating_point() else out.clone()


def batch_kwargs(task, bkw):
    """ITM runs at half the batch (pretrain_src/data/loader.py:130) in the batch-64 case."""
    if bkw["batch_size"] >= 64 and task == "itm":
        return dict(bkw, batch_size=bkw["batch_size"] // 2)
    return bkw


def main():
    os.makedirs(GOLD, exist_ok=True)
    for name, cfg_over, bkw in CASES:
        cfg = ref_shim.pretrain_config(**cfg_over)
        model = ref_shim.load_pretrain_model(cfg).eval()
        ours_cfg = HamtConfig(**cfg_over)
        from hamt_b200.pretrain_cmt import MultiStepNavCMTPreTraining
        ours = MultiStepNavCMTPreTraining(ours_cfg)
        assert list(ours.state_dict().keys()) == list(model.state_dict().keys()), "state_dict key order differs from the reference"
        sd = synth.seeded_state_dict(ours, seed=WEIGHT_SEED)
        model.load_state_dict(sd)
        rec = {"meta": dict(case=name, cfg=cfg_over, batch=bkw, weight_seed=WEIGHT_SEED, batch_seed=BATCH_SEED, rng_seed=RNG_SEED)}
        big = bkw["batch_size"] >= 64
        for task in TASKS:
            b = synth.make_batch(task, seed=BATCH_SEED, **batch_kwargs(task, bkw))
            for cl in (False, True):
                for ac in (False, True):
                    np.random.seed(RNG_SEED)
                    torch.manual_seed(RNG_SEED)
                    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16, enabled=ac):
                        out = model(b, task, compute_loss=cl)
                    outs = out if isinstance(out, tuple) else (out,)
                    rec[f"{task}_{'loss' if cl else 'logits'}{'_autocast' if ac else ''}"] = [compact(task, o, big) for o in outs]
        torch.save(rec, os.path.join(GOLD, f"pretrain_{name}.pt"))
        print("wrote", name, {k: [tuple(t.shape) if torch.is_tensor(t) else "dict" for t in v] for k, v in rec.items() if k != "meta" and not k.endswith("_autocast")})

    # finetune facade (NavCMT modes) on a reduced-depth model
    fcfg = dict(num_l_layers=2, num_x_layers=2, num_h_pano_layers=1, hist_enc_pano=True, no_lang_ca=False, act_pred_token="ob_txt",
                fix_lang_embedding=False, fix_hist_embedding=False, fix_obs_embedding=False, output_attentions=True)
    cfg = ref_shim.pretrain_config(**fcfg)
    model = ref_shim.load_navcmt(cfg).eval()
    from hamt_b200.vilmodel_cmt import NavCMT
    ours = NavCMT(HamtConfig(**fcfg))
    assert list(ours.state_dict().keys()) == list(model.state_dict().keys())
    sd = synth.seeded_state_dict(ours, seed=WEIGHT_SEED)
    model.load_state_dict(sd)
    B, L, O = 3, 16, 11
    b = synth.make_batch("sap", batch_size=B, txt_len=L, hist_len=2, n_ob=O, seed=BATCH_SEED, ragged=True)
    rec = {"meta": dict(cfg=fcfg, B=B, L=L, O=O, weight_seed=WEIGHT_SEED, batch_seed=BATCH_SEED)}
    with torch.no_grad():
        txt = model("language", txt_ids=b["txt_ids"], txt_masks=b["txt_masks"])
        h0 = model("history")
        hs = [h0.expand(B, -1)]
        for t in range(2):
            hs.append(model("history", hist_img_feats=b["hist_img_fts"][:, t], hist_ang_feats=b["hist_ang_fts"][:, t],
                            ob_step_ids=torch.LongTensor([t]), hist_pano_img_feats=b["hist_pano_img_fts"][:, t],
                            hist_pano_ang_feats=b["hist_pano_ang_fts"][:, t]))
        hist = torch.stack(hs, 1)
        hm = torch.ones(B, 3, dtype=torch.bool)
        vis = model("visual", txt_embeds=txt, txt_masks=b["txt_masks"], hist_embeds=hist, hist_masks=hm, ob_img_feats=b["ob_img_fts"],
                    ob_ang_feats=b["ob_ang_fts"], ob_nav_types=b["ob_nav_types"], ob_masks=b["ob_masks"])
    rec.update(language=txt.clone(), history0=h0.clone(), history=[h.clone() for h in hs[1:]], visual=[v.clone() for v in vis])
    torch.save(rec, os.path.join(GOLD, "finetune_navcmt.pt"))
    print("wrote finetune_navcmt")


if __name__ == "__main__":
    main()
