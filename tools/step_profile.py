"""Per-kernel GPU time of one pass over the 12-step task schedule (torch.profiler / CUPTI device timestamps, eager launches).
Development aid; the numbers it prints are not bench values."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hamt_b200
from hamt_b200 import synth
from hamt_b200.config import HamtConfig
from hamt_b200.pretrain_cmt import MultiStepNavCMTPreTraining
from torch.profiler import profile, ProfilerActivity
import bench

tasks = sys.argv[1].split(",") if len(sys.argv) > 1 else bench.SCHEDULE
B = 64
m = MultiStepNavCMTPreTraining(HamtConfig()); m.load_state_dict(synth.seeded_state_dict(m, 0, perturb_ln=False)); m = m.cuda().train()
batches = [{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in synth.make_batch(t, batch_size=bench.batch_size_of(t, B), seed=i, **bench.SHAPE).items()} for i, t in enumerate(tasks)]
def run():
    for i, t in enumerate(tasks):
        np.random.seed(i); torch.manual_seed(i)
        m(batches[i], t).mean().backward(); m.zero_grad(set_to_none=True)
run(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name.replace("hamt::", "").replace("void ", "")
        n = n.split("(")[0][:60]
        agg[n][0] += 1; agg[n][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"tasks={tasks} steps={len(tasks)} GPU kernel time total {tot/1e3:.2f} ms = {tot/1e3/len(tasks):.2f} ms/step")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{k:62s} n/step={n/len(tasks):7.1f} ms/step={us/1e3/len(tasks):7.3f} share={us/tot:6.1%} avg_us={us/n:8.1f}")
