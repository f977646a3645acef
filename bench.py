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
              with pinned HOST batches: H2D of images+masks and D2H of loss + argmax mask every step
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
        graphed = GraphedTrainStep(model, opt, crit, dice, *resident[0], warmup=3)
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
    if rank == 0:
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
                yield host[i % len(host)]

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
    e2e_ms = torch.tensor([max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / (float(e2e_ms) / 1e3)
    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 8
    d2h = 4                     # loss scalar every step (loss.item()); the confusion matrix stays on the GPU

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
        traffic, traffic_src = None, None
        try:   # DRAM bytes per launch of that kernel from the committed ncu launch list of this very command
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            traffic = round(tj[top]["dram_bytes_per_launch"])
            traffic_src = "profiles/r01_traffic.json (ncu dram__bytes_read+write per launch, average over the step's launches)"
        except Exception:
            pass
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
                           "global_batch": B * world, "resolution": R, "parallelism": f"dp{world}" + (" (CUDA-graph local step + one flat NCCL gradient all-reduce)" if (graphed is not None and world > 1) else ""),
                           "l2": "no flush needed: per-step working set (activations, GBs) >> 126 MB L2; "
                                 "two alternating input batches"},
                "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": round(float(e2e_ms) / args.steps, 3),
                        "api": "lmnet_b200.train.train_one_epoch (reference-shaped loop, pinned host batches)"},
                "cuda_graph": bool(graphed is not None and graphed.graph is not None),
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "step_roofline": step_roofline(B, R, ms_total / args.steps),
                "kernels": kernels, "loss": final_loss}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
