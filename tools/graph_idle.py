"""How busy is the GPU inside one captured step?  CUPTI kernel intervals of graph replays (torch.profiler): span of a step, union of the
kernel intervals, idle gaps by the kernel that FOLLOWS the gap.  Development aid (round 2)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
import hamt_b200  # noqa
from hamt_b200 import graph, synth
from hamt_b200.config import HamtConfig
from hamt_b200.pretrain_cmt import MultiStepNavCMTPreTraining

task = sys.argv[1] if len(sys.argv) > 1 else "sap"
dev = torch.device("cuda:0")
model = MultiStepNavCMTPreTraining(HamtConfig())
model.load_state_dict(synth.seeded_state_dict(model, seed=0, perturb_ln=False))
model = model.to(dev).train()
b = synth.make_batch(task, batch_size=64 if task != "itm" else 32, seed=1)
b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
if task == "itm":
    b["_hist_masks_host"] = b["hist_masks"].cpu()
np.random.seed(0); torch.manual_seed(0)
b = graph.add_sync_free_extras(task, b, device=dev)
tr = graph.GraphedTrainer(model)
for _ in range(3):
    tr.step(task, b)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.step(task, b)
    torch.cuda.synchronize()
ev = sorted([(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start],
            key=lambda t: t[0])
# split into replays by the largest gaps, keep the last replay
t0, t1 = ev[0][0], max(e[1] for e in ev)
n = len(ev) // 3
last = ev[-n:]
span = max(e[1] for e in last) - last[0][0]
busy, cur_end, gaps = 0.0, last[0][0], collections.Counter()
conc = 0.0
for s, e, name in last:
    if s > cur_end:
        gaps[name.split("<")[0][:40]] += s - cur_end
        busy += e - s
        cur_end = e
    else:
        if e > cur_end:
            busy += e - cur_end
            cur_end = e
        conc += min(e, cur_end) - s
print(f"task {task}: {n} kernels per replay, span {span / 1e3:.3f} ms, busy (union) {busy / 1e3:.3f} ms = {100 * busy / span:.1f} %, idle {100 * (1 - busy / span):.1f} %, "
      f"sum of kernel durations {sum(e - s for s, e, _ in last) / 1e3:.3f} ms")
print("idle time by the kernel that follows the gap (us):")
for k, v in gaps.most_common(12):
    print(f"   {k:42s} {v:9.1f}")
