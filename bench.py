#!/usr/bin/env python
"""bench.py — LM-Net training throughput on B200 (BASELINE.json: "LM-Net train images/sec @352²").

    python bench.py --gpus 1 --steps 20 --warmup 5            # our arm (sm_100a kernels)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1   # the reference's CPU path

Workload (N=1): BASELINE.json configs[1] — LM-Net training, bf16 autocast, batch 16, 352x352 synthetic
Kvasir-SEG-shaped data, random-init weights, optimiser/loss of the reference's train.py:156-160.
A step = forward + CE+Dice loss + backward + AdamW step over one batch.  N>1: one process per GPU,
16 images per GPU (weak scaling), DistributedDataParallel gradient all-reduce over NCCL.

One JSON line on rank 0:
  value       whole-job images/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e         the same metric through the reference-shaped public API lmnet_b200.train.train_one_epoch
              with pinned HOST batches: H2D of images + masks every step, D2H of the loss scalar every step
              (4 bytes; the confusion matrix of the argmax mask is accumulated on the GPU).  `e2e_as_is`
              repeats it with the reference loop's exact host traffic (utils/train_eval_utils.py:128-156:
              no prefetch, loss.item() every step, argmax mask and labels shipped to the host every step)
  roofline    dominant lmnet_b200 kernel of the step: algorithmic bytes / CUDA-event time (measured
              live in a separate profiled leg, never inside the timed region) vs MEASURED_PEAKS.json
  cpu_baseline the CPU oracle path (reference modules in torch CPU + natten's ops restated in C/OpenMP)
              on a bounded sample, rank 0, N=1 only
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "lm-net_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "LM-Net train images/sec @352x352 (fwd+bwd+AdamW, bf16 autocast, batch 16/GPU)"
FALLBACK_HBM_GBS = 6650.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU")
    ap.add_argument("--res", type=int, default=352)
    ap.add_argument("--cpu-batch", type=int, default=2, help="images per step of the CPU arms (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--graph-collective", action="store_true",
                    help="N > 1: capture the gradient all-reduce + optimiser step inside the CUDA graph as well.  Off by "
                         "default: measured 31.0 vs 30.7 ms/step at 2 GPUs, and NCCL teardown hangs after captured collectives")
    ap.add_argument("--workload", default="train", choices=["train", "cfg1", "na2d", "infer1024"],
                    help="train = BASELINE configs[1]/[2] (default); cfg1 = configs[0] (CPU fwd+bwd, batch 2, 256x256, fp32); "
                         "na2d = configs[3] (na2d micro-benchmark, kernel 3/7, dilation 1/2, every stage shape); "
                         "infer1024 = configs[4] (bf16 inference, batch 8, 1024x1024)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU path (oracle port) — also `--impl reference`
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(res, batch, steps, warmup):
    """fwd+bwd+AdamW of the reference op sequence on the host cores (fp32), images/s."""
    from lmnet_b200.train import build_training, synthetic_batches, train_step
    from oracle.lmnet_ref import build_cpu_reference

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    net = build_cpu_reference(3, 2, seed=42).train()
    opt, crit, dice = build_training(net, "cpu", fused=False)
    images, labels = synthetic_batches(1, batch, res, seed=0, pin=False)[0]
    for _ in range(warmup):
        train_step(net, opt, images, labels, crit, dice, amp_dtype=None)
    t0 = time.perf_counter()
    for _ in range(steps):
        loss, _ = train_step(net, opt, images, labels, crit, dice, amp_dtype=None)
    dt = time.perf_counter() - t0
    return {"value": batch * steps / dt, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"{steps} steps of fwd+bwd+AdamW, batch {batch}, {res}x{res}, fp32, torch CPU ops + "
                      f"C/OpenMP restatement of natten's CPU ops (oracle/), {threads} threads",
            "ms_per_step": 1e3 * dt / steps, "loss": float(loss)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    r = cpu_reference_run(args.res, args.cpu_batch, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"LM-Net training step on the host CPU, bounded sample: batch {args.cpu_batch}, "
                                   f"{args.res}x{args.res}, fp32 (reference op sequence; natten CPU ops restated in C)",
                       "global_batch": args.cpu_batch, "resolution": args.res},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def measured_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def csrc_fingerprint():
    """sha256 over the CUDA sources the library was built from: stamps profiles/*_traffic.json (written by
    tools/launch_list_summary.py) so that a stale ncu traffic file is refused instead of silently reported."""
    import glob
    import hashlib

    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "lm-net_b200", "csrc", "*.cu")) + glob.glob(os.path.join(ROOT, "lm-net_b200", "csrc", "*.cuh"))):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the newest committed ncu launch list whose source stamp matches this tree."""
    import glob

    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            tj = json.load(open(f))
        except Exception:
            continue
        stamp = tj.get("_csrc_sha256")
        rel = os.path.relpath(f, ROOT)
        if stamp is None:
            continue                                     # unstamped files (round 1) are never trusted
        if stamp != csrc_fingerprint():
            return None, f"{rel} is stale (kernels changed since the ncu capture: {stamp} != {csrc_fingerprint()})"
        if kernel in tj:
            return round(tj[kernel]["dram_bytes_per_launch"]), f"{rel} (ncu dram__bytes_read+write per launch, same sources: {stamp})"
        return None, f"{rel} has no entry for {kernel}"
    return None, "no stamped ncu launch list committed"


def step_roofline(batch, res, ms_per_step):
    """t_roof = sum_ops max(B/BW, F/P) over one training step, from the committed per-op byte / FLOP
    table (profiles/step_roofline.json, generated on the CPU by tests/golden/make_step_roofline.py) and the measured
    peaks of this box; `frac` = t_roof / measured step time of one GPU."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "step_roofline.json")))
        if tab["resolution"] != res:
            return None
        try:
            pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            bw, tf, src = float(pk["hbm_gbs"]) * 1e9, float(pk["bf16_tflops_sustained"]) * 1e12, "measured (MEASURED_PEAKS.json)"
        except Exception:
            bw, tf, src = FALLBACK_HBM_GBS * 1e9, 1400e12, "fallback (B200_PROFILING.md)"
        t = sum(max((batch * bf + p) / bw, batch * ff / tf) + max((batch * bb + p) / bw, batch * fb / tf)
                for bf, ff, bb, fb, p in tab["ops"])
        tot = tab["totals_per_image"]
        return {"t_roof_ms": round(1e3 * t, 3), "frac": round(1e3 * t / ms_per_step, 4),
                "algorithmic_GB_per_step": round(batch * (tot["bytes_fwd"] + tot["bytes_bwd"]) / 1e9, 2),
                "algorithmic_GFLOP_per_step": round(batch * (tot["flops_fwd"] + tot["flops_bwd"]) / 1e9, 1),
                "peaks": src, "formula": tab["formula"], "source": "profiles/step_roofline.json"}
    except Exception:
        return None


def run_b200_arm(args):
    import torch.distributed as dist

    from lmnet_b200 import _lib
    from lmnet_b200.distributed import get_rank, get_world_size, init_distributed_mode, wrap_ddp
    from lmnet_b200.model import LM_Net
    from lmnet_b200.train import (ConfusionMetrics, GraphedTrainStep, build_training, synthetic_batches,
                                  train_one_epoch, train_step)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: the b200 arm needs a CUDA device (no CPU fallback exists)")
    _lib.lib()  # fail loudly if the extension is missing
    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")   # required for NCCL inside CUDA-graph capture
    dargs = init_distributed_mode()
    rank, world = get_rank(), get_world_size()
    local = getattr(dargs, "gpu", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cuda.matmul.allow_tf32 = True      # train.py:39
    torch.backends.cudnn.benchmark = True             # train.py:40
    torch.manual_seed(42 + rank)                      # train.py:42-43

    B, R = args.batch, args.res
    use_graph = not args.no_graph
    net = LM_Net(3, 2).to(dev).train()
    # eager: torch DDP (bucketed all-reduce overlapped with backward).  graph: the local step is replayed from
    # a CUDA graph and GraphedTrainStep averages one flat gradient buffer with a single NCCL all-reduce.
    model = net if use_graph else wrap_ddp(net, dev)
    opt, crit, dice = build_training(model, dev, capturable=use_graph)
    host = synthetic_batches(2, B, R, seed=rank)       # pinned host batches
    resident = [(i.to(dev), m.to(dev)) for i, m in host]
    graphed = None
    if use_graph:
        graphed = GraphedTrainStep(model, opt, crit, dice, *resident[0], warmup=3, capture_collective=args.graph_collective)
        if graphed.graph is None and rank == 0:
            print(f"[bench] CUDA-graph capture unavailable, running eagerly: {graphed.fallback_reason}", file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        img, msk = resident[i % len(resident)]
        if graphed is not None:
            return graphed(img, msk)[0]               # device-to-device copy into the static buffers + replay
        return train_step(model, opt, img, msk, crit, dice)[0]

    # ---- device-resident timing (value) ----
    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0 and os.environ.get("LMNET_NO_CLOCK_SAMPLER") != "1":
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ncu_range = os.environ.get("LMNET_NCU_RANGE") == "1"     # `ncu --profile-from-start off`: capture the timed steps only
    if ncu_range:
        torch.cuda.profiler.start()
    e0.record()
    for i in range(args.steps):
        loss = step_resident(i)
    e1.record()
    barrier()
    if ncu_range:
        torch.cuda.profiler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    n_launch = _lib.launch_count() - launches0
    if graphed is not None and graphed.graph is not None:      # replayed launches are not seen by the counter
        n_launch = graphed.library_launches_per_step * args.steps
    launches = torch.tensor([n_launch], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = float(ms)
    value = B * world * args.steps / (ms_total / 1e3)
    final_loss = float(loss)

    # ---- end-to-end timing through the reference-shaped epoch loop, host batches (e2e) ----
    class Loader:
        def __init__(self, n):
            self.n = n

        def __iter__(self):
            for i in range(self.n):
                if trace is not None:
                    trace.append(time.perf_counter())
                yield host[i % len(host)]

    trace = [] if os.environ.get("LMNET_E2E_TRACE") == "1" else None      # diagnosis: host timestamps of the loader calls
    metrics = ConfusionMetrics(2)
    scaler_flag = object()   # non-None => autocast branch, as in the reference loop
    train_one_epoch(model, opt, metrics, 2, Loader(max(1, min(3, args.warmup))), dev, crit, scaler_flag, dice,
                    step_fn=graphed)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    train_one_epoch(model, opt, metrics, 2, Loader(args.steps), dev, crit, scaler_flag, dice, step_fn=graphed)
    e1.record()
    barrier()
    if trace is not None and rank == 0:
        print(f"[bench] e2e trace: events {e0.elapsed_time(e1):.1f} ms, wall {1e3 * (time.perf_counter() - t0):.1f} ms; loader calls at "
              + " ".join(f"{1e3 * (s - t0):.1f}" for s in trace[-args.steps:]), file=sys.stderr)
    e2e_ms = torch.tensor([max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / (float(e2e_ms) / 1e3)
    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 8
    d2h = 4                     # loss scalar every step (loss.item()); the confusion matrix stays on the GPU

    # ---- the reference loop AS IS (utils/train_eval_utils.py:128-156): no prefetch, loss.item() every step,
    #      argmax mask + labels copied to the host for the metric update every step
    asis = dict(metrics_on_device=False, prefetch=False, defer_loss_read=False)
    train_one_epoch(model, opt, metrics, 2, Loader(2), dev, crit, scaler_flag, dice, step_fn=graphed, **asis)
    barrier()
    t0 = time.perf_counter()
    train_one_epoch(model, opt, metrics, 2, Loader(args.steps), dev, crit, scaler_flag, dice, step_fn=graphed, **asis)
    barrier()
    asis_ms = torch.tensor([1e3 * (time.perf_counter() - t0)], device=dev)
    if world > 1:
        dist.all_reduce(asis_ms, op=dist.ReduceOp.MAX)
    asis_value = B * world * args.steps / (float(asis_ms) / 1e3)
    asis_d2h = 4 + 2 * host[0][1].numel() * 8          # loss + int64 argmax mask + int64 labels

    # ---- per-kernel profile leg (separate from both timed regions) ----
    roofline, kernels = None, None
    if rank == 0 and not args.no_profile:
        peak, peak_src = measured_peaks()
        prof_steps = 3
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for i in range(prof_steps):
            train_step(net, opt, *resident[i % len(resident)], crit, dice)
        pe1.record()
        torch.cuda.synchronize()
        prof = _lib.profile_collect()
        _lib.profile_enable(False)
        step_ms = pe0.elapsed_time(pe1) / prof_steps
        kernels = {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] // prof_steps,
                       "alg_GB_per_step": v["alg_bytes"] / prof_steps / 1e9,
                       "achieved_GBs": (v["alg_bytes"] / 1e9) / (v["ms"] / 1e3) if v["alg_bytes"] > 0 else None}
                   for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        own_ms = sum(v["ms_per_step"] for v in kernels.values())
        top = next(k for k, v in kernels.items() if v["achieved_GBs"])
        kt = kernels[top]
        traffic, traffic_src = measured_traffic(top)
        roofline = {"kernel": top, "bound": "hbm", "achieved": round(kt["achieved_GBs"], 1), "peak": peak,
                    "unit": "GB/s", "frac": round(kt["achieved_GBs"] / peak, 4), "traffic": traffic,
                    "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": round(kt["alg_GB_per_step"] * 1e9 / max(kt["launches_per_step"], 1)),
                    "peak_source": peak_src, "launches_per_step": kt["launches_per_step"],
                    "kernel_ms_per_step": round(kt["ms_per_step"], 3),
                    "own_kernels_ms_per_step": round(own_ms, 3), "profiled_step_ms": round(step_ms, 3),
                    "note": "achieved = algorithmic bytes (DESIGN.md §5/§6) / CUDA-event time of that kernel, summed over its "
                            "launches in the profiled steps; the dw kernels are instruction-bound, not HBM-bound (DESIGN.md §5)"}

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(R, args.cpu_batch, 3, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"LM-Net training (fwd+bwd+AdamW), bf16 autocast, batch {B}/GPU, {R}x{R}, "
                                       "synthetic Kvasir-SEG-shaped RGB + binary masks, random-init weights",
                           "global_batch": B * world, "resolution": R, "parallelism": f"dp{world}" + ((" (one CUDA graph per step incl. the flat NCCL gradient all-reduce + AdamW)" if graphed.collective_in_graph else " (CUDA-graph local step, then one flat NCCL gradient all-reduce + AdamW)") if (graphed is not None and world > 1) else ""),
                           "l2": "no flush needed: per-step working set (activations, GBs) >> 126 MB L2; "
                                 "two alternating input batches"},
                "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": round(float(e2e_ms) / args.steps, 3),
                        "api": "lmnet_b200.train.train_one_epoch (reference-shaped loop, pinned host batches)"},
                "e2e_as_is": {"value": round(asis_value, 2), "unit": "images/s", "h2d_bytes_per_step": h2d,
                              "d2h_bytes_per_step": asis_d2h, "ms_per_step": round(float(asis_ms) / args.steps, 3),
                              "api": "train_one_epoch(metrics_on_device=False, prefetch=False, defer_loss_read=False): the "
                                     "reference loop's exact host traffic and synchronisation"},
                "cuda_graph": bool(graphed is not None and graphed.graph is not None),
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "step_roofline": step_roofline(B, R, ms_total / args.steps),
                "kernels": kernels, "loss": final_loss}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_cfg1(args):
    """BASELINE configs[0]: LM-Net fwd+bwd, batch 2, 256x256, fp32 on the host CPU (reference op sequence + the C/OpenMP
    restatement of natten's CPU ops).  Pure CPU line; printed by both arms."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    r = cpu_reference_run(256, 2, max(1, min(args.steps, 5)), max(1, min(args.warmup, 2)))
    line = {"impl": "reference" if args.impl == "reference" else "b200", "metric": "LM-Net fwd+bwd+AdamW images/sec @256x256 on the host CPU (BASELINE configs[0])",
            "value": r["value"], "unit": "images/s", "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "configs[0]: LM-Net fwd+bwd, batch 2, 256x256 synthetic RGB + binary mask, fp32, CPU",
                                            "global_batch": 2, "resolution": 256},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_na2d(args):
    """BASELINE configs[3]: fused na2d forward and backward, kernel 3 and 7, dilation 1/2, at every LM-Net stage shape
    (bf16, batch `--batch`, 12 heads).  value = algorithmic GB/s over all cases (fwd 4N, bwd 7N bytes); L2 flushed."""
    from lmnet_b200 import _lib
    from natten.functional import na2d

    dev = torch.device("cuda")
    _lib.lib()
    peak, peak_src = measured_peaks()
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    gen = torch.Generator(device=dev).manual_seed(0)

    def timed(fn, iters):
        for _ in range(args.warmup):
            fn()
        ts = []
        for _ in range(iters):
            flush.fill_(1.0)
            torch.cuda._sleep(2_000_000)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    B, R0 = args.batch, args.res
    cases, tot_bytes, tot_ms, launches0 = [], 0.0, 0.0, _lib.launch_count()
    sampler = ClockSampler(0)
    sampler.start()
    for lvl, hd in enumerate((1, 2, 4, 8)):
        R = R0 >> lvl
        for K, d in ((3, 1), (7, 1), (7, 2)):
            q, k, v = (torch.randn(B, R, R, 12, hd, device=dev, dtype=torch.bfloat16, generator=gen).requires_grad_() for _ in range(3))
            rpb = (0.02 * torch.randn(12, 2 * K - 1, 2 * K - 1, device=dev, generator=gen)).requires_grad_()
            go = torch.randn(B, R, R, 12, hd, device=dev, dtype=torch.bfloat16, generator=gen)
            nb = q.numel() * 2
            with torch.no_grad():
                tf = timed(lambda: na2d(q, k, v, K, d, rel_pos_bias=rpb), args.steps)
            out = na2d(q, k, v, K, d, rel_pos_bias=rpb)
            tb = timed(lambda: torch.autograd.grad(out, (q, k, v, rpb), go, retain_graph=True), args.steps)
            cases.append({"shape": [B, R, R, 12, hd], "kernel": K, "dilation": d, "fwd_ms": round(tf, 4), "bwd_ms": round(tb, 4),
                          "fwd_GBs": round(4 * nb / tf / 1e6, 1), "bwd_GBs": round(7 * nb / tb / 1e6, 1),
                          "fwd_frac": round(4 * nb / tf / 1e6 / peak, 4), "bwd_frac": round(7 * nb / tb / 1e6 / peak, 4)})
            tot_bytes += 11 * nb
            tot_ms += tf + tb
            del q, k, v, out
    clocks = sampler.stop()
    value = tot_bytes / tot_ms / 1e6
    line = {"metric": "na2d fwd+bwd algorithmic GB/s vs roofline (BASELINE configs[3])", "value": round(value, 1), "unit": "GB/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(tot_ms, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"configs[3]: fused na2d fwd+bwd, kernel 3/7, dilation 1/2, [{B},R,R,12,hd] for R,hd = "
                                   f"{R0},1 / {R0 // 2},2 / {R0 // 4},4 / {R0 // 8},8; a step = all 12 cases once", "l2": "flushed (512 MB write) between iterations"},
            "roofline": {"bound": "hbm", "achieved": round(value, 1), "peak": peak, "unit": "GB/s", "frac": round(value / peak, 4),
                         "traffic": None, "peak_source": peak_src},
            "gpu_launches": _lib.launch_count() - launches0, "clocks": clocks, "cases": cases}
    print(json.dumps(line), flush=True)


def run_infer1024(args):
    """BASELINE configs[4]: LM-Net inference, bf16 autocast, batch 8, 1024x1024, eval mode (running statistics)."""
    from lmnet_b200 import _lib
    from lmnet_b200.model import LM_Net

    dev = torch.device("cuda")
    _lib.lib()
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    net = LM_Net(3, 2).to(dev).eval()
    B, R = 8, 1024
    host = torch.randn(B, 3, R, R).pin_memory()
    x = host.to(dev)
    sampler = ClockSampler(0)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for _ in range(args.warmup):
            net(x)
        torch.cuda.synchronize()
        sampler.start()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y = net(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        launches = _lib.launch_count() - l0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pred = net(host.to(dev, non_blocking=True)).argmax(1).to(torch.uint8).cpu()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    line = {"metric": "LM-Net inference images/sec @1024x1024 (bf16, batch 8; BASELINE configs[4])", "value": round(B * 1e3 / ms, 2),
            "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[4]: LM-Net inference, bf16 autocast, batch 8, 1024x1024, eval mode", "global_batch": B,
                       "resolution": R, "l2": "activations >> 126 MB L2"},
            "e2e": {"value": round(B * 1e3 / e2e_ms, 2), "unit": "images/s", "h2d_bytes_per_step": host.numel() * 4,
                    "d2h_bytes_per_step": B * R * R, "ms_per_step": round(e2e_ms, 3),
                    "api": "LM_Net(x.to(device)).argmax(1) -> uint8 mask on the host"},
            "gpu_launches": launches, "clocks": clocks, "finite": bool(torch.isfinite(y.float()).all()),
            "peak_memory_GiB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2), "pred_shape": list(pred.shape)}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.workload == "cfg1":
        run_cfg1(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "na2d":
        run_na2d(args)
    elif args.workload == "infer1024":
        run_infer1024(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
