"""Kernel-level breakdown of one LM-Net training step (torch.profiler / CUPTI; no nsys in the image).

    python tools/profile_step.py [--batch 16] [--res 352] [--out gpurun_out/step_profile.txt]

Prints the top CUDA kernels by total device time over 3 profiled steps and the share taken by the
lmnet_b200 kernels vs library (cuDNN / cuBLAS / ATen) kernels.  Used to decide what to widen next
(SURVEY.md §8 f1-f4)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200")]

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from lmnet_b200.model import LM_Net  # noqa: E402
from lmnet_b200.train import build_training, synthetic_batches, train_step  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--res", type=int, default=352)
    ap.add_argument("--out", default=None)
    ap.add_argument("--rows", type=int, default=45)
    a = ap.parse_args()
    dev = torch.device("cuda")
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(42)
    net = LM_Net(3, 2).to(dev).train()
    opt, crit, dice = build_training(net, dev)
    img, msk = (t.to(dev) for t in synthetic_batches(1, a.batch, a.res, pin=False)[0])
    for _ in range(4):
        train_step(net, opt, img, msk, crit, dice)
    torch.cuda.synchronize()
    steps = 3
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            train_step(net, opt, img, msk, crit, dice)
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "self_device_time_total", None)
        if t is None:
            t = getattr(e, "self_cuda_time_total", 0)
        if t > 0 and e.device_type is not None and "cuda" in str(e.device_type).lower():
            rows.append((t / steps / 1e3, e.count // steps, e.key))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    own = sum(r[0] for r in rows if "lmnet" in r[2])
    lines = [f"LM-Net training step, batch {a.batch}, {a.res}x{a.res}, bf16 autocast: device time per step "
             f"{total:.2f} ms over {sum(r[1] for r in rows)} kernel launches; lmnet_b200 kernels {own:.2f} ms "
             f"({100 * own / total:.1f} %), library kernels {total - own:.2f} ms",
             f"{'ms/step':>9} {'%':>6} {'calls':>6}  kernel"]
    for ms, n, k in rows[:a.rows]:
        lines.append(f"{ms:9.3f} {100 * ms / total:6.2f} {n:6d}  {k[:150]}")
    text = "\n".join(lines)
    print(text)
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        open(a.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
