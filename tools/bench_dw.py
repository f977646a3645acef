"""Per-kernel timing of the fused depthwise section (raw op, training forward + backward) at LM-Net's four stage
shapes of the 352x352 / batch-16 workload, using the library's own per-launch CUDA-event profile.

    python tools/bench_dw.py [--levels 1 2 3 4] [--iters 5] [--no-tma]

Prints microseconds and achieved algorithmic GB/s per kernel (SURVEY §8 d5 byte counts) and per level; the L2 is
flushed between iterations by the op's own working set (>= 190 MB per level-1 tensor triple)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200")]

import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--levels", type=int, nargs="*", default=[1, 2, 3, 4])
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--res", type=int, default=352)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--no-tma", action="store_true")
a = ap.parse_args()
if a.no_tma:
    os.environ["LMNET_DW_NO_TMA"] = "1"

from lmnet_b200 import _lib  # noqa: E402
from lmnet_b200.model import ReparamConv  # noqa: E402
from lmnet_b200.reparam import fused_dw_bn_gelu  # noqa: E402

PEAK = 6550.1
dev = torch.device("cuda")
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
grand = 0.0
for level in a.levels:
    width = 12 * 2 ** (level - 1)
    E, R = 2 * width, a.res // 2 ** (level - 1)
    m = ReparamConv(width, E, width).to(dev).train()
    x1 = torch.randn(a.batch, E, R, R, device=dev).to(torch.bfloat16).requires_grad_()
    gz = torch.randn(a.batch, E, R, R, device=dev).to(torch.bfloat16)
    gp = torch.randn(a.batch, E, device=dev)

    def step():
        z, p = fused_dw_bn_gelu(m, x1)
        torch.autograd.backward([z, p], [gz, gp])

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(a.iters):
        flush.zero_()
        step()
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    T = a.batch * E * R * R * 2
    tot = 0.0
    print(f"level {level}: [B={a.batch}, E={E}, {R}x{R}] bf16, one tensor = {T / 1e6:.1f} MB")
    for k, v in prof.items():
        if not k.startswith("dw_"):
            continue
        us = 1e3 * v["ms"] / a.iters
        per = v["launches"] // a.iters
        gbs = v["alg_bytes"] / a.iters / (us * 1e-6) / 1e9 if us > 0 else 0.0
        tot += us
        print(f"  {k:18s} {us:9.1f} us  x{per}  {gbs:8.1f} GB/s alg  ({100 * gbs / PEAK:5.1f} % of measured HBM peak)")
    gbs = 8 * T / (tot * 1e-6) / 1e9
    grand += tot * 4       # four ReparamConv blocks per level in LM-Net
    print(f"  section fwd+bwd    {tot:9.1f} us        {gbs:8.1f} GB/s alg  ({100 * gbs / PEAK:5.1f} %)  [8T bytes]")
print(f"whole model (16 blocks): {grand / 1e3:.2f} ms per training step for the depthwise section")
