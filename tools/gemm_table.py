"""Per-shape table of the GEMM launches inside one step: device duration (CUPTI) matched in launch order with the shapes
recorded on the host.  Development aid."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hamt_b200
from hamt_b200 import synth, ops
from hamt_b200.config import HamtConfig
from hamt_b200.pretrain_cmt import MultiStepNavCMTPreTraining
from torch.profiler import profile, ProfilerActivity
import bench
task = sys.argv[1] if len(sys.argv) > 1 else "sap"
B = 64
m = MultiStepNavCMTPreTraining(HamtConfig()); m.load_state_dict(synth.seeded_state_dict(m, 0, perturb_ln=False)); m = m.cuda().train()
b = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synth.make_batch(task, batch_size=bench.batch_size_of(task, B), seed=1, **bench.SHAPE).items()}
calls = []
orig = ops.gemm
def rec(a, bb, **kw):
    M, K = (a.shape[1], a.shape[0]) if kw.get("a_mn") else a.shape
    N = bb.shape[1] if kw.get("b_mn") else bb.shape[0]
    calls.append((M, N, K, int(bool(kw.get("a_mn"))), int(bool(kw.get("b_mn"))), kw.get("act", 0), kw.get("aux_mode", 0), bool(kw.get("accumulate"))))
    return orig(a, bb, **kw)
def run():
    np.random.seed(0); torch.manual_seed(0)
    m(b, task).mean().backward(); m.zero_grad(set_to_none=True)
run(); torch.cuda.synchronize()
ops.gemm = rec
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
ops.gemm = orig
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "gemm_tcgen05" in e.name], key=lambda e: e.time_range.start)
assert len(evs) == len(calls), (len(evs), len(calls))
agg = collections.OrderedDict()
for c, e in zip(calls, evs):
    a = agg.setdefault(c, [0, 0.0]); a[0] += 1; a[1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"task={task}: {len(calls)} GEMM launches, {tot/1e3:.2f} ms, {sum(2.0*c[0]*c[1]*c[2]*n for c,(n,_) in agg.items())/tot/1e6:.0f} TFLOP/s average")
print(f"{'M':>6} {'N':>6} {'K':>6} aMN bMN act aux acc {'n':>3} {'us/launch':>10} {'TF/s':>7} {'share':>6}")
for c, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{c[0]:6d} {c[1]:6d} {c[2]:6d} {c[3]:3d} {c[4]:3d} {c[5]:3d} {c[6]:3d} {int(c[7]):3d} {n:3d} {us/n:10.1f} {2.0*c[0]*c[1]*c[2]/(us/n)/1e6:7.0f} {us/tot:6.1%}")
