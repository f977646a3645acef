"""Per-op algorithmic byte / FLOP count of one LM-Net training step -> profiles/step_roofline.json

    python tests/golden/make_step_roofline.py            # CPU only; runs the oracle's LM-Net once for the shapes

Definition (the formula bench.py evaluates with the measured peaks of the box it runs on):

    t_roof(step) = sum over ops of  max(B_fwd / BW, F_fwd / P)  +  max(B_bwd / BW, F_bwd / P)

with BW = measured HBM bandwidth, P = measured dense bf16 throughput, es = 2 bytes (bf16 activations) and, per op,
the ALGORITHMIC traffic of a perfectly fused implementation of that op alone:

* hot-path units use SURVEY §8's figures: fused neighbourhood attention 4N / 7N elements (d4) and 4K²N / 10K²N FLOPs;
  the ReparamConv branch section (4x depthwise conv + BN, sum, GELU) 3T / 5T elements (d5), FLOPs 2*40 per element
  forward, twice that backward;
* Conv2d / Linear: forward reads the input, writes the output (+ parameters); backward reads dout and the saved
  input, writes dx (+ parameter gradients); FLOPs 2*MACs forward, 4*MACs backward;
* BatchNorm2d in training mode: 3T forward (statistics pass + apply pass), 5T backward; LayerNorm 2T / 3T;
* every other leaf module (activations, pooling, up-sampling, dropout): in + out forward, dout + saved + dx backward.
Ops that are not modules (residual adds, concatenations, the SE multiply, permutes) are NOT counted, so t_roof is a
lower bound and the reported fraction errs against us.  Everything is per image; a batch multiplies activations
(not parameters) by B.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200"), os.path.join(ROOT, "tests")]

from lmnet_b200 import model as M  # noqa: E402
from oracle.lmnet_ref import build_cpu_reference  # noqa: E402
from oracle.na2d_ref import OracleNeighborhoodAttention2D  # noqa: E402

ES = 2  # bytes per activation element (bf16)


def numel(x):
    if isinstance(x, torch.Tensor):
        return x.numel()
    if isinstance(x, (tuple, list)):
        return sum(numel(t) for t in x)
    return 0


def main(res=352):
    net = build_cpu_reference(3, 2, seed=0).train()
    ops = []          # (class, act_bytes_fwd, flops_fwd, act_bytes_bwd, flops_bwd, param_bytes)
    skip = set()      # leaves inside the fused ReparamConv branch section (counted as one unit)
    for mod in net.modules():
        if isinstance(mod, M.ReparamConv):
            for name in ("large_conv", "square_conv", "ver_conv", "hor_conv"):
                skip.update(id(m) for m in getattr(mod, name).modules())
            skip.add(id(mod.active))

    def leaf_hook(mod, inp, out):
        if id(mod) in skip:
            return
        i, o = numel(inp), numel(out)
        p = sum(t.numel() for t in mod.parameters(recurse=False)) * 4          # fp32 master parameters
        cls = type(mod).__name__
        if isinstance(mod, torch.nn.Conv2d):
            macs = o * (mod.in_channels // mod.groups) * mod.kernel_size[0] * mod.kernel_size[1]
            ops.append((cls, (i + o) * ES, 2 * macs, (o + 2 * i) * ES, 4 * macs, p))
        elif isinstance(mod, torch.nn.Linear):
            macs = o * mod.in_features
            ops.append((cls, (i + o) * ES, 2 * macs, (o + 2 * i) * ES, 4 * macs, p))
        elif isinstance(mod, torch.nn.BatchNorm2d):
            ops.append((cls, 3 * i * ES, 0, 5 * i * ES, 0, p))
        elif isinstance(mod, torch.nn.LayerNorm):
            ops.append((cls, 2 * i * ES, 0, 3 * i * ES, 0, p))
        else:
            ops.append((cls, (i + o) * ES, 0, (o + 2 * i) * ES, 0, p))

    def dw_hook(mod, inp, out):          # the fused branch section of a ReparamConv: T = E*H*W per image
        x = inp[0]
        T = mod.expand_channels * x.shape[0] * x.shape[2] * x.shape[3]
        ops.append(("ReparamConv.dw_section", 3 * T * ES, 80 * T, 5 * T * ES, 160 * T, 0))

    def na_hook(mod, inp, out):          # fused neighbourhood attention core: N = tokens * C
        N, KK = inp[0].numel(), mod.kernel_size ** 2
        ops.append(("NeighborhoodAttention2D.core", 4 * N * ES, 4 * KK * N, 7 * N * ES, 10 * KK * N, mod.rpb.numel() * 4))

    for mod in net.modules():
        if isinstance(mod, M.ReparamConv):
            mod.register_forward_hook(dw_hook)
        elif isinstance(mod, OracleNeighborhoodAttention2D):
            mod.register_forward_hook(na_hook)
        elif not list(mod.children()):
            mod.register_forward_hook(leaf_hook)
    with torch.no_grad():
        net(torch.randn(1, 3, res, res))

    by_class = {}
    for cls, bf, ff, bb, fb, p in ops:
        c = by_class.setdefault(cls, {"ops": 0, "bytes_fwd": 0, "flops_fwd": 0, "bytes_bwd": 0, "flops_bwd": 0})
        c["ops"] += 1
        c["bytes_fwd"] += bf; c["flops_fwd"] += ff; c["bytes_bwd"] += bb; c["flops_bwd"] += fb
    out = {
        "what": "algorithmic bytes / FLOPs per op of one LM-Net training step, PER IMAGE at %dx%d, bf16 activations "
                "(SURVEY.md §8 d6); generated by tests/golden/make_step_roofline.py" % (res, res),
        "formula": "t_roof = sum_ops max(B_fwd/BW, F_fwd/P) + max(B_bwd/BW, F_bwd/P); B = batch*act_bytes + param_bytes",
        "resolution": res,
        "ops": [[bf, ff, bb, fb, p] for _, bf, ff, bb, fb, p in ops],
        "by_class": by_class,
        "totals_per_image": {"bytes_fwd": sum(o[1] for o in ops), "flops_fwd": sum(o[2] for o in ops),
                             "bytes_bwd": sum(o[3] for o in ops), "flops_bwd": sum(o[4] for o in ops)},
    }
    path = os.path.join(ROOT, "profiles", "step_roofline.json")
    json.dump(out, open(path, "w"))
    bw, peak = 6550.1e9, 1406.4e12
    B = 16
    t = sum(max((B * bf + p) / bw, B * ff / peak) + max((B * bb + p) / bw, B * fb / peak) for _, bf, ff, bb, fb, p in ops)
    tot = out["totals_per_image"]
    print(f"{len(ops)} ops; per image: {tot['bytes_fwd'] / 1e6:.1f} MB fwd, {tot['bytes_bwd'] / 1e6:.1f} MB bwd, "
          f"{tot['flops_fwd'] / 1e9:.2f} GFLOP fwd, {tot['flops_bwd'] / 1e9:.2f} GFLOP bwd")
    print(f"t_roof(batch 16) = {1e3 * t:.2f} ms at 6550.1 GB/s, 1406.4 TFLOP/s  ->  {B / t:.0f} images/s at the roofline")
    for cls, c in sorted(by_class.items(), key=lambda kv: -(kv[1]["bytes_fwd"] + kv[1]["bytes_bwd"])):
        print(f"  {cls:32s} {c['ops']:4d} ops  {(c['bytes_fwd'] + c['bytes_bwd']) / 1e6:8.1f} MB/image  "
              f"{(c['flops_fwd'] + c['flops_bwd']) / 1e9:7.2f} GFLOP/image")


if __name__ == "__main__":
    main()
