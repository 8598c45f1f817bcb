"""Generate tests/golden/adamw_reference.pt from the UNMODIFIED reference optimizer (pretrain_src/optim/adamw.py, sched.py) and
torch.nn.utils.clip_grad_norm_, driven the way main_r2r.py:252-281 drives them.  Run in the build container:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_optim.py

TEST INFRASTRUCTURE ONLY.  The scenario (shapes, seeds, which parameters have a gradient in which step) is `scenario()` below and
is re-created from seeds by the tests; the fixture stores the reference's trajectory only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(5, 7), (64,), (3, 130), (1,), (257, 3)]
WD = [0.01, 0.0, 0.01, 0.0, 0.01]
STEPS = 7
LR0, WARMUP, TOTAL, MAX_NORM = 5e-5, 3, 10, 5.0


def scenario():
    g = torch.Generator().manual_seed(123)
    params = [torch.randn(s, generator=g) for s in SHAPES]
    grads = []
    for t in range(STEPS):
        row = []
        for i, s in enumerate(SHAPES):
            scale = 30.0 if t == 2 else 0.1                      # step 2 triggers clipping
            gi = torch.randn(s, generator=g) * scale
            if (t + i) % 4 == 3 or (i == 3 and t < 2):           # parameter i unused by this step's task -> grad None
                gi = None
            row.append(gi)
        grads.append(row)
    return params, grads


def main():
    sys.path.insert(0, "/root/reference/pretrain_src")
    from optim.adamw import AdamW
    from optim.sched import warmup_linear
    params, grads = scenario()
    ps = [torch.nn.Parameter(p.clone()) for p in params]
    opt = AdamW([{"params": [p], "weight_decay": w} for p, w in zip(ps, WD)], lr=LR0, betas=(0.9, 0.98))
    norms, traj = [], []
    for t in range(STEPS):
        lr = LR0 * warmup_linear(t + 1, WARMUP, TOTAL)
        if lr <= 0:
            lr = 1e-8
        for gp in opt.param_groups:
            gp["lr"] = lr
        for p, g in zip(ps, grads[t]):
            p.grad = None if g is None else g.clone()
        norms.append(float(torch.nn.utils.clip_grad_norm_(ps, MAX_NORM)))
        opt.step()
        opt.zero_grad()
        traj.append([p.detach().clone() for p in ps])
    rec = dict(norms=norms, params=traj, steps=[opt.state[p].get("step", 0) if p in opt.state else 0 for p in ps],
               exp_avg=[opt.state[p]["exp_avg"].clone() for p in ps], exp_avg_sq=[opt.state[p]["exp_avg_sq"].clone() for p in ps])
    out = os.path.join(ROOT, "tests", "golden", "adamw_reference.pt")
    torch.save(rec, out)
    print("wrote", out, "norms", [round(n, 3) for n in norms], "steps", rec["steps"])


if __name__ == "__main__":
    main()
