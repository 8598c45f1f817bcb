#!/usr/bin/env python
"""bench.py -- pretrain samples/sec (fwd+bwd) of the HAMT hot path on N B200s of one node.

    python bench.py --gpus 1 --steps 24 --warmup 12
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

Workload = BASELINE.json configs[1]: R2R 6-proxy-task pretrain (cmt-vitbase-6tasks), fixed 768-d features,
txt 80 / hist 15 x 36 views / obs 36+STOP, batch 64 per GPU (ITM 32, pretrain_src/data/loader.py:130), task schedule =
the mix_ratio multiset 5 MLM : 1 SAP : 1 SAR : 1 SPREL : 2 MRC : 2 ITM (pretrain_r2r.json:43-58), train mode (dropout on),
fwd + bwd + gradient zeroing per step (optimizer excluded, as in the metric).  A "step" is one batch of one task.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "pretrain samples/sec (fwd+bwd, R2R 6-task, bsz64, txt80/hist15x36/obs36)"
SCHEDULE = ["mlm", "sap", "mlm", "mrc", "mlm", "itm", "sar", "mlm", "mrc", "sprel", "mlm", "itm"]   # 5:1:1:1:2:2
SHAPE = dict(txt_len=80, hist_len=15, n_pano=36, n_ob=37, feat=768)
# algorithmic forward GFLOP per sample; fwd+bwd = 3x
FWD_GFLOP = dict(mlm=34.37, sap=36.78, sar=36.74, sprel=36.82, mrc=33.80, itm=63.25)

# BASELINE.json configs: [1] = r2r (the headline, default), [3] = rxr (L = 300, 512-d features, XLM-R vocabulary 250 002, hist 20 x 36,
# no MRC: pretrain_rxr.json), [4] = r4r (hist 40 x 36).  The other two are separate bench lines (--config), not the headline.
CONFIGS = {
    "r2r": dict(workload="R2R 6-task pretrain (cmt-vitbase-6tasks), txt80/hist15x36/obs37, schedule 5mlm:1sap:1sar:1sprel:2mrc:2itm",
                shape=SHAPE, schedule=SCHEDULE, cfg={}, batch=64, ref_config="r2r_model_config.json"),
    "rxr": dict(workload="RxR long-instruction stress (rxr_xlm_model_config: 512-d features, vocab 250002), txt300/hist20x36/obs37, schedule 5mlm:1sap:1sar:1sprel:2itm",
                shape=dict(txt_len=300, hist_len=20, n_pano=36, n_ob=37, feat=512, vocab_hi=250000),
                schedule=["mlm", "sap", "mlm", "itm", "mlm", "sar", "mlm", "sprel", "mlm", "itm"],
                cfg=dict(image_feat_size=512, vocab_size=250002, max_position_embeddings=514), batch=64, ref_config="rxr_xlm_model_config.json"),
    "r4r": dict(workload="R4R long-horizon history stress, txt80/hist40x36/obs37, schedule 5mlm:1sap:1sar:1sprel:2mrc:2itm",
                shape=dict(txt_len=80, hist_len=40, n_pano=36, n_ob=37, feat=768), schedule=SCHEDULE, cfg={}, batch=64, ref_config="r2r_model_config.json"),
}


def fwd_gflop_per_sample(task, L, T, O=37, P=36, H=768, I=3072, layers_l=9, layers_x=4, layers_p=2):
    """Algorithmic forward FLOPs of one sample (SURVEY 8a formulae: BertLayer = S(8H^2 + 4HI) + 4 S^2 H; x-layer =
    (L + V)(16H^2 + 4HI) + 8 L V H + 4 L^2 H + 4 V^2 H); used for the non-headline configs (the headline uses SURVEY's table)."""
    def bert(S):
        return S * (8 * H * H + 4 * H * I) + 4 * S * S * H
    has_ob = task in ("sap", "sar", "sprel")
    V = T + 1 + (O if has_ob else 0)
    R = 5 if task == "itm" else 1
    f = layers_l * bert(L) + T * (layers_p * bert(P) + P * 2 * H * 768) + (T + (O if has_ob else 0)) * 2 * H * 768
    f += R * layers_x * ((L + V) * (16 * H * H + 4 * H * I) + 8 * L * V * H + 4 * L * L * H + 4 * V * V * H)
    return f / 1e9


def batch_size_of(task, B):
    return B // 2 if task == "itm" else B


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.1)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
def reference_kind():
    """'reference' when the unmodified reference sources are reachable (baseline/_ref installed by build(), or /root/reference),
    else 'port' (oracle/hamt_oracle.py)."""
    from oracle import ref_shim
    return "reference" if ref_shim.reference_available() else "port"


def reference_samples_per_sec(device, steps, warmup, conf, autocast=False, threads=None, tasks=None):
    """fwd + bwd samples/s of the reference's own implementation of the path: the UNMODIFIED `MultiStepNavCMTPreTraining`
    (pretrain_src/model/pretrain_cmt.py, imported through oracle/ref_shim.py) in train mode -- dropout ON (pretrain_r2r.json:21) --
    at the bench's batch size and task schedule, on the host cores (device cpu, all threads) or as torch eager on one GPU
    (fp32, or under torch.autocast(bfloat16)).  Falls back to the oracle port when the reference sources are absent."""
    import hamt_b200  # noqa: F401
    from hamt_b200 import synth
    from hamt_b200.config import HamtConfig
    from hamt_b200.pretrain_cmt import MultiStepNavCMTPreTraining
    dev = torch.device(device)
    if dev.type == "cpu":
        threads = threads or os.cpu_count()
        torch.set_num_threads(threads)
    kind = reference_kind()
    tasks = tasks or conf["schedule"]
    B = conf["batch"]
    batches = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synth.make_batch(t, batch_size=batch_size_of(t, B), seed=i, **conf["shape"]).items()}
               for i, t in enumerate(tasks)]
    if kind == "reference":
        from oracle import ref_shim
        cfg = ref_shim.pretrain_config(config_name=conf["ref_config"])
        model = ref_shim.load_pretrain_model(cfg)
        model.load_state_dict(synth.seeded_state_dict(model, seed=0, perturb_ln=False))
        model = model.to(dev).train()

        def fwd_bwd(i, t):
            loss = model(batches[i], t, compute_loss=True)
            loss.float().mean().backward()
            model.zero_grad(set_to_none=True)
    else:
        from oracle import hamt_oracle as O
        cfg = HamtConfig(**conf["cfg"])
        m = MultiStepNavCMTPreTraining(cfg)
        sd = {k: v.to(dev).requires_grad_(v.is_floating_point()) for k, v in synth.seeded_state_dict(m, seed=0).items()}
        sd["mlm_head.predictions.decoder.weight"] = sd["bert.embeddings.word_embeddings.weight"]
        del m

        def fwd_bwd(i, t):
            loss = O.pretrain_forward(sd, cfg, batches[i], t, compute_loss=True)
            loss.float().mean().backward()
            for v in sd.values():
                v.grad = None
    gpu = dev.type == "cuda"
    if gpu:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, None
    for i in range(warmup + steps):
        if i == warmup:
            n = 0
            if gpu:
                torch.cuda.synchronize(); e0.record()
            t0 = time.perf_counter()
        j = i % len(tasks)
        np.random.seed(i); torch.manual_seed(i)
        with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=autocast):
            fwd_bwd(j, tasks[j])
        n += batch_size_of(tasks[j], B)
    if gpu:
        e1.record(); torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3
    else:
        dt = time.perf_counter() - t0
    return n / dt, dt, (threads if dev.type == "cpu" else None), kind


def run_reference(args, rank):
    if rank != 0:
        return
    conf = CONFIGS[args.config]
    warm = min(args.warmup, 2)       # CPU steps take seconds each: two untimed steps settle the allocator / thread pool
    sps, dt, threads, kind = reference_samples_per_sec("cpu", args.steps, warm, conf)
    B = conf["batch"]
    what = ("unmodified reference MultiStepNavCMTPreTraining (baseline/_ref via oracle/ref_shim.py)" if kind == "reference"
            else "oracle port (oracle/hamt_oracle.py; reference sources not found)")
    gpu_eager = None
    if torch.cuda.is_available() and not args.no_gpu_eager:
        gpu_eager = {"what": f"{what} as torch eager on 1 GPU, train mode (dropout on), batch {B}, same schedule, fwd+bwd, CUDA events, 12 steps after 6 "
                             "(informational; not the reference arm's value)"}
        for name, ac in (("bf16_autocast", True), ("fp32", False)):
            try:
                gpu_eager[name + "_samples_per_s"] = round(reference_samples_per_sec("cuda:0", 12, 6, conf, autocast=ac)[0], 1)
            except Exception as e:  # noqa: BLE001
                gpu_eager[name + "_error"] = f"{type(e).__name__}: {str(e)[:200]}"
            torch.cuda.empty_cache()
    line = {"impl": "reference", "metric": METRIC, "value": round(sps, 3), "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": conf["workload"], "global_batch": B, "per_gpu_batch": B, "itm_batch": B // 2, "parallelism": "cpu",
                       "mode": "train (dropout 0.1)", "note": f"{what}, fp32 on the host cores; {warm} untimed warm-up steps"},
            "cpu_baseline": {"value": round(sps, 3), "unit": "samples/s", "cores": threads, "kind": kind,
                             "sample": f"{args.steps} steps of the task schedule at batch {B} (ITM {B // 2}), fwd+bwd, train mode, fp32 torch CPU, {threads} threads ({dt:.1f} s)"},
            "e2e": {"value": round(sps, 3), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if gpu_eager is not None:
        line["gpu_eager_reference"] = gpu_eager
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
VIT_FWD_GFLOP_PER_IMAGE = 35.1        # 12 blocks x (197 x (8 H^2 + 4 H I) + 4 x 197^2 x H) + the patch projection, H = 768, I = 3072


def run_e2e_config(args, rank, world, local_rank):
    """BASELINE.json configs[2]: end-to-end stage (main_r2r_image.py / pretrain_r2r_e2e.json), the ViT-B/16 backbone trained jointly with
    the cross-modal transformer from raw 224 x 224 views.  Per sample and step: T history views + 36 candidate views through the
    backbone WITH gradient, T x 36 panorama views without (image_vilmodel.py:40-59).  Device-resident synthetic images (a batch is
    gigabytes of fp32 pixels; the reference reads JPEGs from LMDB in its dataloader, out of scope), eager launches, fwd + bwd."""
    import torch.distributed as dist
    import hamt_b200  # noqa: F401
    from hamt_b200 import _lib, dp, synth
    from hamt_b200.config import HamtConfig
    from hamt_b200.image_pretrain import MultiStepNavImagePreTraining
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, L = (args.batch if args.batch != 64 else 2), 5, 60            # pretrain_r2r_e2e.json: train_batch_size 1, max_txt_len 60
    schedule = args.tasks.split(",") if args.tasks else SCHEDULE
    model = MultiStepNavImagePreTraining(HamtConfig())
    sd = synth.seeded_state_dict(model, seed=0, perturb_ln=False)
    sd.update({"bert.vision_backbone." + k: v for k, v in synth.seeded_vit_state_dict(model.bert.vision_backbone, 1).items()})
    model.load_state_dict(sd)
    model = model.to(dev).train()
    arena = model.arena()
    arena.ensure()
    uniq = sorted(set(schedule))
    batches = {t: synth.make_image_batch(t, batch_size=(B // 2 if t == "itm" and B > 1 else B), txt_len=L, hist_len=T, seed=7 + rank, device=str(dev)) for t in uniq}
    n_img = {t: sum(int(np.prod(v.shape[:-3])) for k, v in batches[t].items() if k.endswith("images")) for t in uniq}
    n_img_grad = {t: sum(int(np.prod(v.shape[:-3])) for k, v in batches[t].items() if k in ("hist_images", "ob_images")) for t in uniq}

    use_graphs = not args.no_graphs
    trainer = None
    if use_graphs:
        from hamt_b200 import graph
        trainer = graph.GraphedTrainer(model, post_backward=(lambda: dp.sync_grads(arena)) if world > 1 else None)
        for t in uniq:       # masked-row indices / the ITM negative plan are host-side parts of the batch (graph.py)
            hb = dict(batches[t])
            if t == "itm":
                hb["_hist_masks_host"] = hb["hist_masks"].cpu()
            np.random.seed(0); torch.manual_seed(0)
            batches[t] = graph.add_sync_free_extras(t, hb, device=dev)

    def step(i):
        t = schedule[i % len(schedule)]
        np.random.seed(i); torch.manual_seed(i)
        if trainer is not None:
            trainer.step(t, batches[t])
        else:
            loss = model(batches[t], t, compute_loss=True)
            loss.mean().backward()
            if world > 1:
                dp.sync_grads(arena)
            model.zero_grad(set_to_none=True)
        return batches[t]["txt_ids"].shape[0], n_img[t], n_img_grad[t]

    for i in range(max(args.warmup, 3, len(schedule) if use_graphs else 0)):       # every task's graph is captured in warm-up
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ns = ni = ng = 0
    for i in range(args.steps):
        a, b_, c = step(i)
        ns += a; ni += b_; ng += c
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    clocks = sampler.result()
    if rank == 0:
        peaks = load_peaks()
        vit_tf = VIT_FWD_GFLOP_PER_IMAGE * 1e9 * ((ni - ng) + 3 * ng) * world / (ms * 1e-3) / 1e12
        print(json.dumps({
            "metric": "pretrain samples/sec (fwd+bwd, R2R 6-task END-TO-END: ViT-B/16 on raw 224x224 views)", "value": round(ns * world / (ms * 1e-3), 2),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"R2R end-to-end pretrain (BASELINE configs[2]; pretrain_r2r_e2e.json), per sample {T} history views + 36 candidate views "
                                   f"through ViT-B/16 with gradient, {T} x 36 panorama views without; txt{L}/hist{T}x36/obs37, 6-task schedule",
                       "per_gpu_batch": B, "itm_batch": max(1, B // 2), "global_batch": B * world, "parallelism": f"dp{world}", "mode": "train (dropout 0.1)",
                       "launch": "cuda-graph per task" if use_graphs else "eager", "l2": "activations of a step (GBs) >> 126 MB L2"},
            "images_per_s": round(ni * world / (ms * 1e-3), 1), "images_with_grad_per_s": round(ng * world / (ms * 1e-3), 1),
            "vit_tflops": round(vit_tf, 1), "vit_frac_of_burst_peak": round(vit_tf / peaks["tf_burst"] / world, 3),
            "gpu_launches": int(_lib.launch_count() - l0) if trainer is None else int(sum(trainer.steps[k].native_launches for k in trainer.steps) * args.steps / max(1, len(trainer.steps))),
            "clocks": clocks,
            "e2e": None, "e2e_note": "device-resident images only: the reference's stage-2 dataloader decodes LMDB JPEGs on the host (out of scope, SURVEY row 10); "
                                     "a host leg would time PCIe, not this path"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--config", default="r2r", choices=sorted(CONFIGS) + ["e2e"],
                    help="BASELINE.json config: r2r = configs[1] (headline), e2e = configs[2] (ViT-B/16 end-to-end stage), rxr = configs[3], r4r = configs[4]")
    ap.add_argument("--tasks", default=None, help="comma list overriding the 6-task schedule (e.g. sap)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="--impl reference: skip the informational torch-eager-on-GPU timing of the oracle port")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--dp-mode", default="overlap", choices=["graph", "after", "overlap"],
                    help="gradient exchange: captured at the end of the step graph / eager after the replay / per-layer overlap (eager only)")
    ap.add_argument("--nccl-sms", type=int, default=-1,
                    help="N > 1: SMs the backward GEMMs leave free for the overlapped NCCL all-reduce kernels (= NCCL channel cap); "
                         "-1 = default (0 = no reservation)")
    ap.add_argument("--nccl-channels", type=int, default=0, help="N > 1: cap NCCL's channel count (= CTAs of the all-reduce kernel) without reserving SMs; 0 = NCCL default")
    ap.add_argument("--nccl-high-prio", action="store_true", help="N > 1: NCCL on a high-priority stream (measured at N = 2: no gain, profiles/r02_scale.txt)")
    ap.add_argument("--grad-wire", default="fp32", choices=["fp32", "bf16"],
                    help="N > 1: dtype of the gradient all-reduce on the wire (bf16 = opt-in compressed exchange, unmeasured; default fp32 like DDP)")
    ap.add_argument("--no-store-leg", action="store_true", help="skip the second e2e variant (device-resident FeatureStore, indices from the host)")
    ap.add_argument("--quick", action="store_true", help="device-resident value only (no e2e / roofline / cpu legs): development aid")
    ap.add_argument("--no-graphs", action="store_true", help="eager launches from Python instead of one CUDA graph per (task, batch signature)")
    ap.add_argument("--profile-range", action="store_true",
                    help="bracket the device-resident timed region with cudaProfilerStart/Stop (ncu --profile-from-start off) and exit after it; "
                         "numbers printed in this mode are not bench values")
    ap.add_argument("--diag", action="store_true", help="print the e2e host-time breakdown and the per-signature GEMM table to stderr")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.config == "e2e":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the reference's end-to-end stage does not import as shipped (image_pretrain.py:11, SURVEY row 10)"}))
            return
        run_reference(args, rank)
        return
    if args.config == "e2e":
        run_e2e_config(args, rank, world, local_rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    import hamt_b200  # noqa: F401
    from hamt_b200 import _lib, dp, graph, loader, ops, synth
    from hamt_b200.config import HamtConfig
    from hamt_b200.pretrain_cmt import MultiStepNavCMTPreTraining

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:      # torchrun exports OMP_NUM_THREADS=1: give the host-side batch / weight preparation its share of the cores back
        torch.set_num_threads(max(1, (os.cpu_count() or 8) // world))
    nccl_sms = 0
    if world > 1:
        # measured at N = 2 (profiles/r01_scale_n2.txt): reserving SMs for NCCL costs more GEMM time than the overlap gains
        # (0: 8850, 8: 8481, 16: 8777 samples/s), so the default is no reservation; the knob stays for larger N
        nccl_sms = 0 if args.nccl_sms < 0 else args.nccl_sms
        if nccl_sms > 0:     # the all-reduce kernel gets exactly the SMs the backward GEMMs leave free
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(nccl_sms))
            os.environ.setdefault("NCCL_MIN_NCHANNELS", str(nccl_sms))
        if args.nccl_channels > 0:
            os.environ.setdefault("NCCL_MAX_NCHANNELS", str(args.nccl_channels))
        pg_opts = None
        if args.nccl_high_prio:
            try:
                pg_opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            except Exception:  # noqa: BLE001 -- older torch: keep the default stream priority
                pg_opts = None
        dist.init_process_group("nccl", device_id=dev, pg_options=pg_opts)
    conf = CONFIGS[args.config]
    schedule = args.tasks.split(",") if args.tasks else conf["schedule"]
    SHAPE = conf["shape"]

    cfg = HamtConfig(**conf["cfg"])
    model = MultiStepNavCMTPreTraining(cfg)
    model.load_state_dict(synth.seeded_state_dict(model, seed=0, perturb_ln=False))
    model = model.to(dev).train()
    arena = model.arena()
    arena.ensure()
    overlap = None
    if world > 1 and args.dp_mode == "overlap":
        overlap = dp.LayerOverlap(arena, wire_dtype=torch.bfloat16 if args.grad_wire == "bf16" else None)
        arena.layer_hook = overlap.layer_done

    B = args.batch
    host_batches, dev_batches = [], []
    for i, t in enumerate(schedule):
        b = synth.make_batch(t, batch_size=batch_size_of(t, B), seed=1000 * rank + i, **SHAPE)
        hb = dict(b)                                     # collated host batch; packed into pinned blobs by prefetch() in warm-up
        host_batches.append(hb)
        db = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
        if t in ("mlm", "mrc"):      # device-resident leg: the row indices are part of the resident batch
            db = graph.add_sync_free_extras(t, db)
            hb = graph.add_sync_free_extras(t, hb)
            host_batches[-1] = hb
        if t == "itm":
            db["_hist_masks_host"] = b["hist_masks"]
        dev_batches.append(db)

    use_graphs = not args.no_graphs

    def exchange():
        if world > 1:
            (overlap.finish() if overlap else dp.sync_grads(arena, wire_dtype=torch.bfloat16 if args.grad_wire == "bf16" else None))

    in_graph = world > 1 and args.dp_mode in ("graph", "overlap")
    bwd_sm_limit = (torch.cuda.get_device_properties(dev).multi_processor_count - nccl_sms) if nccl_sms > 0 else 0
    trainer = graph.GraphedTrainer(model, post_backward=exchange if in_graph else None, bwd_sm_limit=bwd_sm_limit) if use_graphs else None
    graph_launches = {}

    def step(i, batch, eager=False):
        """One step = one batch of one task: fwd + bwd (+ gradient exchange) + gradient reset.  `batch` may live on the host
        (pinned) or on the device."""
        task = schedule[i % len(schedule)]
        np.random.seed(i); torch.manual_seed(i)          # ITM negative sampling uses the host RNGs, as in the reference
        if use_graphs and not eager:
            # masked-row indices (MLM / MRC) and the ITM negative plan are prepared on the host side of the batch
            b = graph.add_sync_free_extras(task, batch) if ("itm_plan" not in batch and "txt_label_rows" not in batch and "hist_mrc_rows" not in batch) else batch
            loss = trainer.step(task, b)
            if world > 1 and not in_graph:
                exchange()
            graph_launches[0] = graph_launches.get(0, 0) + trainer.steps[graph._signature(task, b)].native_launches
            return loss
        if not torch.is_tensor(batch["txt_ids"]) or batch["txt_ids"].device.type != "cuda":
            batch = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}
        loss = model(batch, task, compute_loss=True)
        if bwd_sm_limit:
            _lib.load().hamt_gemm_set_sm_limit(bwd_sm_limit)
        loss.mean().backward()
        if bwd_sm_limit:
            _lib.load().hamt_gemm_set_sm_limit(0)
        exchange()
        model.zero_grad(set_to_none=True)
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    copy_stream = torch.cuda.Stream(device=dev)

    packed = [None] * len(schedule)      # one PackedBatch (pinned host blob + device blob) per schedule slot, built in warm-up

    def prefetch(i):
        """Host -> device copy of step i's batch from pinned memory on a side stream, like the reference's PrefetchLoader
        (pretrain_src/data/loader.py:90-125): the copy of batch i+1 overlaps the compute of batch i.  One cudaMemcpyAsync of
        the packed blob (hamt_b200.loader); the ITM negative plan is drawn on the host (reference RNG order) into the blob."""
        j = i % len(schedule)
        task = schedule[j]
        if packed[j] is None:
            np.random.seed(i); torch.manual_seed(i)
            hb = graph.add_sync_free_extras(task, host_batches[j]) if use_graphs else host_batches[j]
            packed[j] = loader.PackedBatch(hb, dev)
        elif use_graphs and task == "itm":
            np.random.seed(i); torch.manual_seed(i)
            packed[j].fill(graph.add_sync_free_extras(task, host_batches[j]), only_plan=True)
        db = packed[j].to_device(copy_stream)
        return db, packed[j].ready

    diag_t = {"prefetch": 0.0, "step": 0.0, "item": 0.0}

    def timed(n_steps, from_host, read_loss=True, copy=True):
        for k in diag_t:
            diag_t[k] = 0.0
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = _lib.launch_count()
        e0.record()
        samples, d2h = 0, 0
        DEPTH = 2                                           # batches in flight ahead of the compute (absorbs PCIe / host jitter)
        queue = [prefetch(k) for k in range(min(DEPTH, n_steps))] if from_host else None
        reader = loader.LossReader(depth=4)
        for i in range(n_steps):
            j = i % len(schedule)
            if from_host:
                t0 = time.perf_counter()
                batch, ev = queue.pop(0)
                if i + DEPTH < n_steps:
                    queue.append(prefetch(i + DEPTH) if copy else (packed[(i + DEPTH) % len(schedule)].dev_views, ev))
                torch.cuda.current_stream().wait_event(ev)
                t1 = time.perf_counter()
                lm = step(i, batch).float().mean()         # tiny reduction enqueued behind the step
                t2 = time.perf_counter()
                if read_loss:
                    reader.push(lm)                        # async device -> host copy of this step's result + event
                    if reader.pending() > 1:               # read the PREVIOUS step's loss while this step runs
                        assert np.isfinite(reader.pop())
                        d2h += 4
                t3 = time.perf_counter()
                diag_t["prefetch"] += t1 - t0; diag_t["step"] += t2 - t1; diag_t["item"] += t3 - t2
            else:
                step(i, dev_batches[j])
            samples += batch_size_of(schedule[j], B)
        while from_host and reader.pending():
            assert np.isfinite(reader.pop())
            d2h += 4
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        n_launch = _lib.launch_count() - launches0 + graph_launches.pop(0, 0)
        return ms, samples, n_launch, d2h

    for i in range(max(args.warmup, len(schedule) if use_graphs else 0)):      # every (task, signature) graph is captured in warm-up
        step(i, dev_batches[i % len(schedule)])
    for i in range(len(schedule)):                       # build the packed slots + run every step once from the host path
        db, ev = prefetch(i)
        torch.cuda.current_stream().wait_event(ev)
        step(i, db)
    graph_launches.clear()
    h2d_bytes = float(np.mean([pb.layout.payload_bytes for pb in packed]))     # bytes of the tensors copied per step (padding excluded)
    sampler = ClockSampler(local_rank)
    sampler.start()
    n_graphs0 = len(trainer.steps) if trainer else 0
    if args.profile_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        ms, samples, launches, _ = timed(args.steps, from_host=False)
        torch.cuda.profiler.stop()
        if rank == 0:
            print(json.dumps({"profile_range": True, "steps": args.steps, "launches": int(launches), "note": "run under a profiler: not a bench value"}), flush=True)
        os._exit(0)
    ms, samples, launches, _ = timed(args.steps, from_host=False)
    clocks = sampler.result()
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "value": round(samples * world / (ms * 1e-3), 1), "unit": "samples/s", "n_gpus": world, "ms_per_step": round(ms / args.steps, 3),
                              "nccl_sms": nccl_sms, "dp_mode": args.dp_mode, "grad_wire": args.grad_wire, "nccl_channels": args.nccl_channels,
                              "nccl_high_priority_stream": args.nccl_high_prio, "clocks": clocks}), flush=True)
        if world > 1:
            if trainer is not None:
                trainer.steps.clear()          # captured graphs hold NCCL work: destroy_process_group hangs while they are alive
            torch.cuda.synchronize()
            dist.barrier()
            dist.destroy_process_group()
        return
    ms_e2e, samples_e2e, _, d2h = timed(args.steps, from_host=True)
    if args.diag and rank == 0:
        print(f"[diag] e2e {ms_e2e / args.steps:.2f} ms/step; host ms/step: " + ", ".join(f"{k}={v / args.steps * 1e3:.2f}" for k, v in diag_t.items()), file=sys.stderr)
        m2, _, _, _ = timed(args.steps, from_host=True, read_loss=False)
        print(f"[diag] e2e without per-step loss read: {m2 / args.steps:.2f} ms/step; host: " + ", ".join(f"{k}={v / args.steps * 1e3:.2f}" for k, v in diag_t.items()), file=sys.stderr)
        m3, _, _, _ = timed(args.steps, from_host=True, copy=False)
        print(f"[diag] e2e without H2D copies (staging reused): {m3 / args.steps:.2f} ms/step; host: " + ", ".join(f"{k}={v / args.steps * 1e3:.2f}" for k, v in diag_t.items()), file=sys.stderr)
        t0 = time.perf_counter()
        m4, _, _, _ = timed(args.steps, from_host=False)
        print(f"[diag] device-resident again: {m4 / args.steps:.2f} ms/step, host wall {1e3 * (time.perf_counter() - t0) / args.steps:.2f} ms/step", file=sys.stderr)
    n_graphs1 = len(trainer.steps) if trainer else 0

    # ---- second end-to-end variant: view features resident in HBM (hamt_b200.FeatureStore), the host ships INDICES ----
    # The reference collates ~103 MB of fp32 view features per batch on the CPU and copies them (data/r2r_data.py:264-329, loader.py:78-125);
    # here a batch is (panorama, view) indices + the small per-sample tensors, the features are gathered on the device in bf16.
    e2e_store = None
    if use_graphs and not args.no_store_leg:
        from hamt_b200.feature_store import FeatureStore
        V = 2048
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        feats = torch.randn(V, 36, SHAPE["feat"] + 1000, device=dev, generator=gen)
        store = FeatureStore([f"scan_{i}" for i in range(V)], feats, dev, image_feat_size=SHAPE["feat"])
        del feats
        FEAT_KEYS = ("hist_img_fts", "hist_pano_img_fts", "hist_pano_ang_fts", "ob_img_fts", "ob_ang_fts", "hist_img_probs")
        idx_packed = []
        g_cpu = torch.Generator().manual_seed(4321 + rank)
        for j, t in enumerate(schedule):
            hb = {k: v for k, v in host_batches[j].items() if k not in FEAT_KEYS and not k.startswith("_")}
            Bt, T = batch_size_of(t, B), SHAPE["hist_len"]
            hb["hist_pano"] = torch.randint(0, V, (Bt, T), generator=g_cpu)
            hb["hist_view"] = torch.randint(0, 36, (Bt, T), generator=g_cpu)
            if t in ("sap", "sar", "sprel"):
                hb["ob_pano"] = torch.randint(0, V, (Bt,), generator=g_cpu)
                hb["ob_view"] = torch.randint(0, 36, (Bt,), generator=g_cpu)
            if t == "itm":
                hb["_hist_masks_host"] = hb["hist_masks"]
            np.random.seed(j); torch.manual_seed(j)
            idx_packed.append(loader.PackedBatch(graph.add_sync_free_extras(t, hb), dev))

        def assemble(t, db):
            b = {k: v for k, v in db.items() if k not in ("hist_pano", "hist_view", "ob_pano", "ob_view", "_packed")}
            b.update(store.assemble_history(db["hist_pano"], db["hist_view"], with_pano=True, with_probs=(t == "mrc"),
                                            mrc_mask=db.get("hist_mrc_masks") if t == "mrc" else None))
            if "ob_pano" in db:
                b.update(store.assemble_observation(db["ob_pano"], db["ob_view"]))
            return b

        def store_steps(n_steps):
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n, d2h_b = 0, 0
            reader = loader.LossReader(depth=4)
            q = [(idx_packed[k % len(schedule)].to_device(copy_stream), idx_packed[k % len(schedule)].ready) for k in range(min(2, n_steps))]
            for i in range(n_steps):
                j = i % len(schedule)
                t = schedule[j]
                db, ev = q.pop(0)
                if i + 2 < n_steps:
                    jj = (i + 2) % len(schedule)
                    if t == "itm" or schedule[jj] == "itm":
                        np.random.seed(i + 2); torch.manual_seed(i + 2)
                        if schedule[jj] == "itm":
                            hbj = {k: v for k, v in idx_packed[jj].host_views.items() if not k.startswith("_") and k != "itm_plan"}
                            hbj["_hist_masks_host"] = hbj["hist_masks"]
                            idx_packed[jj].fill(graph.add_sync_free_extras("itm", hbj), only_plan=True)
                    q.append((idx_packed[jj].to_device(copy_stream), idx_packed[jj].ready))
                torch.cuda.current_stream().wait_event(ev)
                np.random.seed(i); torch.manual_seed(i)
                lm = trainer.step(t, assemble(t, db)).float().mean()
                reader.push(lm)
                if reader.pending() > 1:
                    assert np.isfinite(reader.pop()); d2h_b += 4
                n += batch_size_of(t, B)
            while reader.pending():
                assert np.isfinite(reader.pop()); d2h_b += 4
            e1.record()
            sync_all()
            t_ms = e0.elapsed_time(e1)
            if world > 1:
                tt = torch.tensor([t_ms], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t_ms = float(tt.item())
            return t_ms, n, d2h_b

        store_steps(len(schedule))                       # captures the bf16-feature graphs, one per task
        ms_s, n_s, d2h_s = store_steps(args.steps)
        e2e_store = {"value": round(n_s * world / (ms_s * 1e-3), 1), "unit": "samples/s", "ms_per_step": round(ms_s / args.steps, 3),
                     "h2d_bytes_per_step": int(np.mean([pb.layout.payload_bytes for pb in idx_packed])), "d2h_bytes_per_step": int(d2h_s / args.steps),
                     "how": f"view features of {V} panoramas resident in HBM as bf16 (hamt_b200.FeatureStore); per step the host ships (panorama, view) indices + "
                            "the small per-sample tensors in one pinned blob, the device gathers the history / panorama / observation features "
                            "(hamt_gather_rows_pad_bf16) and runs the captured step; loss read back through the pinned ring"}
        del store
        torch.cuda.empty_cache()
    # copy-only leg: how long the per-step host->device transfer takes on this box when nothing else runs
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(copy_stream)
    for i in range(len(schedule)):
        prefetch(i)
    c1.record(copy_stream)
    torch.cuda.synchronize()
    h2d_ms = c0.elapsed_time(c1) / len(schedule)

    # ---- rooflines: per-launch CUDA-event timing on the launching stream ----
    # Every GEMM / attention / LayerNorm call of one eager pass over the schedule is recorded by signature; each distinct signature is
    # then replayed from a small CUDA graph (REP launches, no CPU gaps) between two events on the launching stream, with the operands of
    # its first occurrence.  achieved = sum(count * algorithmic work) / sum(count * launch duration).  These are ISOLATED timings (the
    # kernel alone on the GPU at boost clocks), so the GEMM fraction is taken against the BURST bf16 peak of MEASURED_PEAKS.json.
    calls = {}
    hooked = {n: getattr(ops, n) for n in ("gemm", "attn_fwd", "attn_bwd", "ln_fwd", "ln_bwd")}
    REP = 4

    def _record(kind, sig, work, fn, args_, kw):
        ent = calls.get((kind, sig))
        if ent is None:
            calls[(kind, sig)] = [1, work, fn, args_, kw]
        else:
            ent[0] += 1

    def recording_gemm(a, b, **kw):
        M, K = (a.shape[1], a.shape[0]) if kw.get("a_mn") else a.shape
        N = b.shape[1] if kw.get("b_mn") else b.shape[0]
        out = hooked["gemm"](a, b, **kw)
        sig = (M, N, K, bool(kw.get("a_mn")), bool(kw.get("b_mn")), kw.get("act", 0), kw.get("aux_mode", 0), bool(kw.get("accumulate")),
               str(out.dtype), kw.get("bias") is not None)
        _record("gemm", sig, 2.0 * M * N * K, hooked["gemm"], (a, b), dict(kw, out=out))
        return out

    def recording_attn_fwd(q, k, v, B_, Sq, Sk, heads, mask, drop=ops.NO_DROP, need_lse=True, out=None):
        r = hooked["attn_fwd"](q, k, v, B_, Sq, Sk, heads, mask, drop, need_lse, out)
        # algorithmic bytes: q, k, v read + out written, bf16, 64 per head
        _record("attn_fwd", (B_, Sq, Sk, heads, mask is not None, drop.p > 0), B_ * heads * (2 * Sq + 2 * Sk) * 128.0, hooked["attn_fwd"],
                (q, k, v, B_, Sq, Sk, heads, mask, drop, need_lse, r[0]), {})
        return r

    def recording_attn_bwd(q, k, v, out, lse, dout, dq, dk, dv, B_, Sq, Sk, heads, mask, drop=ops.NO_DROP, dbias=None):
        hooked["attn_bwd"](q, k, v, out, lse, dout, dq, dk, dv, B_, Sq, Sk, heads, mask, drop, dbias)
        # algorithmic bytes: q, k, v, dout read + dq, dk, dv written (the saved output is an implementation choice, not counted)
        scratch = torch.zeros_like(dbias) if dbias is not None else None
        _record("attn_bwd", (B_, Sq, Sk, heads, mask is not None, drop.p > 0, dbias is not None), B_ * heads * (3 * Sq + 4 * Sk) * 128.0, hooked["attn_bwd"],
                (q, k, v, out, lse, dout, dq, dk, dv, B_, Sq, Sk, heads, mask, drop, scratch), {})

    def recording_ln_fwd(x, res, gamma, beta, eps, drop=ops.NO_DROP, save_z=True, inplace_z=True, out=None, out32=None):
        r = hooked["ln_fwd"](x, res, gamma, beta, eps, drop, save_z, inplace_z, out, out32)
        M, H = x.shape
        per = 2 + (0 if res is None else res.element_size()) + 2 + (4 if out32 is not None else 0) + (2 if save_z else 0)
        _record("ln_fwd", (M, H, None if res is None else str(res.dtype), out32 is not None, save_z, drop.p > 0), float(M) * H * per, hooked["ln_fwd"],
                (x, res, gamma, beta, eps, drop, save_z, inplace_z, r[0], out32), {})
        return r

    def recording_ln_bwd(dy, z, mean, rstd, gamma, dgamma, dbeta, dbias=None, dres_in=None, want_dx=True, want_dres=True, drop=ops.NO_DROP, dres_out=None):
        r = hooked["ln_bwd"](dy, z, mean, rstd, gamma, dgamma, dbeta, dbias, dres_in, want_dx, want_dres, drop, dres_out)
        M, H = dy.shape
        per = 2 + 2 + (2 if want_dx else 0) + (2 if want_dres else 0) + (2 if dres_in is not None else 0)
        sc = [None if t is None else torch.zeros_like(t) for t in (dgamma, dbeta, dbias)]
        _record("ln_bwd", (M, H, want_dx, want_dres, dres_in is not None, drop.p > 0), float(M) * H * per, hooked["ln_bwd"],
                (dy, z, mean, rstd, gamma, sc[0], sc[1], sc[2], dres_in, want_dx, want_dres, drop, r[1]), {})
        return r

    for n, f in (("gemm", recording_gemm), ("attn_fwd", recording_attn_fwd), ("attn_bwd", recording_attn_bwd), ("ln_fwd", recording_ln_fwd), ("ln_bwd", recording_ln_bwd)):
        setattr(ops, n, f)
    n_prof = len(schedule)
    for i in range(n_prof):
        step(i, dev_batches[i % len(schedule)], eager=True)
    torch.cuda.synchronize()
    for n, f in hooked.items():
        setattr(ops, n, f)
    agg = {}           # kind -> [work, us, launches]
    for (kind, sig), (cnt, work, fn, fargs, fkw) in calls.items():
        fn(*fargs, **fkw)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(REP):
                fn(*fargs, **fkw)
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        t_us = e0.elapsed_time(e1) * 1e3 / REP
        a = agg.setdefault(kind, [0.0, 0.0, 0])
        a[0] += cnt * work; a[1] += cnt * t_us; a[2] += cnt
        if args.diag and rank == 0:
            unit = f"TF={work / t_us / 1e6:7.1f}" if kind == "gemm" else f"GB/s={work / t_us / 1e3:7.1f}"
            print(f"[{kind}] {sig} n={cnt:4d} us={t_us:8.1f} {unit} tot_ms={cnt * t_us / 1e3:7.3f}", file=sys.stderr)
        del g
    calls.clear()
    gemm_flops, gemm_us, n_gemm = agg.get("gemm", [0.0, 1.0, 0])
    gemm_ms = gemm_us * 1e-3
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12
    peaks = load_peaks()
    step_ms = ms / args.steps
    hbm_rooflines = []
    for kind, name in (("attn_fwd", "attn_fwd_tc_kernel / attn_fwd_kernel"), ("attn_bwd", "attn_bwd_kernel / attn_bwd_tc_kernel"), ("ln_fwd", "ln_fwd_kernel"), ("ln_bwd", "ln_bwd_kernel")):
        if kind in agg:
            w, us, n = agg[kind]
            gbs = w / us / 1e3
            hbm_rooflines.append({"kernel": name, "bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm"], "unit": "GB/s", "frac": round(gbs / peaks["hbm"], 3),
                                  "launches": n, "share_of_step": round(us * 1e-3 / n_prof / step_ms, 3), "traffic": None})
    # DRAM bytes per launch of the most expensive GEMM signature, from the committed ncu --set full capture (profiles/r02_traffic.json)
    traffic, traffic_of = None, None
    tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.isfile(tp):
        tj = json.load(open(tp))
        traffic, traffic_of = tj.get("dram_bytes_per_launch"), {"signature": tj.get("signature"), "algorithmic_bytes_per_launch": tj.get("algorithmic_bytes_per_launch"), "source": tj.get("source")}

    if rank == 0:
        value = samples * world / (ms * 1e-3)
        e2e_value = samples_e2e * world / (ms_e2e * 1e-3)
        if args.config == "r2r":
            alg_flops_per_step = 3e9 * np.mean([FWD_GFLOP[t] * batch_size_of(t, B) for t in schedule])
        else:
            alg_flops_per_step = 3e9 * np.mean([fwd_gflop_per_sample(t, SHAPE["txt_len"], SHAPE["hist_len"]) * batch_size_of(t, B) for t in schedule])
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": conf["workload"] if not args.tasks else f"tasks={args.tasks}; " + conf["workload"],
                       "global_batch": B * world, "per_gpu_batch": B, "itm_batch": B // 2, "parallelism": f"dp{world}", "mode": "train (dropout 0.1)", "launch": "cuda-graph per (task, batch signature)" if use_graphs else "eager",
                       "l2": "no explicit flush: each step streams ~10 GB of activations + 1.4 GB of weights/grads, >> 126 MB L2",
                       "grad_exchange": ("layer-overlapped all_reduce(AVG) on flat fp32 grads" if overlap else
                                         ("all_reduce(AVG) on the flat fp32 grad slices, " + ("captured at the end of the step graph" if in_graph and use_graphs else "after backward")))
                       if world > 1 else "none", "nccl_sms_reserved_in_backward": nccl_sms, "grad_wire_dtype": args.grad_wire,
                       "nccl_high_priority_stream": args.nccl_high_prio if world > 1 else None},
            "gpu_launches": int(launches),
            "model_tflops": round(alg_flops_per_step * args.steps / (ms * 1e-3) / 1e12 * 1.0, 1),
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": "samples/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h / args.steps),
                    "ms_per_step": round(ms_e2e / args.steps, 3), "h2d_only_ms_per_step": round(h2d_ms, 3),
                    "graphs_captured_in_timed_regions": n_graphs1 - n_graphs0,
                    "how": "packed pinned host batch -> one cudaMemcpyAsync on a copy stream, 2 batches ahead (PrefetchLoader style, hamt_b200.loader) -> captured step -> per-step loss through a pinned slot + event (read one step behind)"},
            "roofline": {"kernel": "gemm_tcgen05_kernel", "bound": "tensor", "achieved": round(achieved_tf, 1), "peak": peaks["tf_burst"],
                         "unit": "TFLOP/s", "frac": round(achieved_tf / peaks["tf_burst"], 3), "traffic": traffic, "traffic_of": traffic_of,
                         "peak_source": peaks["src"] + " (burst: the signatures are timed in isolation)", "frac_of_sustained_peak": round(achieved_tf / peaks["tf_sustained"], 3),
                         "launches": n_gemm, "share_of_step": round(gemm_ms / n_prof / step_ms, 3),
                         "how": "sum of 2*M*N*K over every GEMM launch of one pass over the task schedule / sum of their durations; each distinct "
                                "GEMM signature of the pass is timed with CUDA events around a graph of 4 back-to-back launches on its real operands"},
            "roofline_hbm": hbm_rooflines,
        }
        if e2e_store is not None:
            line["e2e_feature_store"] = e2e_store
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample of the SAME workload at the SAME batch size: the first three steps of the schedule after one untimed step
            torch.cuda.empty_cache()
            sps, dt, threads, kind = reference_samples_per_sec("cpu", 3, 1, conf)
            line["cpu_baseline"] = {"value": round(sps, 3), "unit": "samples/s", "cores": threads, "kind": kind,
                                    "sample": ("unmodified reference MultiStepNavCMTPreTraining (baseline/_ref)" if kind == "reference" else "oracle port (oracle/hamt_oracle.py)")
                                              + f", train mode (dropout on), fp32 torch CPU, fwd+bwd, 3 steps of the task schedule at batch {B} (ITM {B // 2}) after 1 warm-up step ({dt:.1f} s)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        try:
            if trainer is not None:
                trainer.steps.clear()          # drop the captured graphs (they hold NCCL work) before the group goes away
            torch.cuda.synchronize()
            dist.barrier()
            dist.destroy_process_group()
        except Exception as e:  # pragma: no cover - teardown only
            print(f"[bench] teardown: {type(e).__name__}: {e}", file=sys.stderr)


if __name__ == "__main__":
    main()
