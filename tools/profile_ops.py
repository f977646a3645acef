"""Where do the small ATen launches of a training step come from?  One profiled step with shapes and Python stacks:
for the glue operators (copy_ / to / contiguous / sum / add / mul / cat / fill) prints count, device time, input shapes
and the innermost source line of THIS repository that issued them.

    python tools/profile_ops.py [--out gpurun_out/ops.txt]
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200")]

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from lmnet_b200.model import LM_Net  # noqa: E402
from lmnet_b200.train import build_training, synthetic_batches, train_step  # noqa: E402

GLUE = ("aten::copy_", "aten::sum", "aten::add", "aten::add_", "aten::mul", "aten::cat", "aten::fill_", "aten::zero_",
        "aten::clone", "aten::_to_copy", "aten::div", "aten::mean", "aten::flip", "aten::constant_pad_nd")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--res", type=int, default=352)
    a = ap.parse_args()
    dev = torch.device("cuda")
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(42)
    net = LM_Net(3, 2).to(dev).train()
    opt, crit, dice = build_training(net, dev)
    img, msk = (t.to(dev) for t in synthetic_batches(1, a.batch, a.res, pin=False)[0])
    for _ in range(3):
        train_step(net, opt, img, msk, crit, dice)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True,
                 experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
        train_step(net, opt, img, msk, crit, dice)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.name not in GLUE:
            continue
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if t <= 0:
            continue
        where = "?"
        for fr in (e.stack or []):
            if ("lmnet_b200/" in fr or "natten/" in fr) and "train.py" not in fr and "_lib.py" not in fr:
                where = fr.split("/repo/")[-1][-70:]
                break
        if where == "?" and e.stack:
            where = "bwd/" + e.stack[0][-60:]
        key = (e.name, str(e.input_shapes)[:70], where)
        agg[key][0] += 1
        agg[key][1] += t
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    lines = [f"{'us':>9} {'calls':>5}  op / input shapes / issuing line"]
    for (name, shapes, where), (n, t) in rows[:90]:
        lines.append(f"{t:9.1f} {n:5d}  {name:16s} {shapes:70s} {where}")
    lines.append(f"total glue device time {sum(v[1] for v in agg.values()) / 1e3:.2f} ms in {sum(v[0] for v in agg.values())} ops")
    text = "\n".join(lines)
    print(text)
    if a.out:
        open(a.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
