"""Runs one hot-path unit in isolation (for ncu): a ReparamConv block or a NeighborhoodAttention2D module,
training forward + backward, at one of LM-Net's four stage shapes of the 352x352 / batch-16 workload.

    python tools/run_block.py --unit reparam --level 1 --iters 3
    python tools/run_block.py --unit na --level 1 --iters 3 [--kernel 7 --dilation 2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200")]

import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--unit", default="reparam", choices=["reparam", "na", "conv", "natt"])
ap.add_argument("--level", type=int, default=1)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--res", type=int, default=352)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--kernel", type=int, default=3)
ap.add_argument("--dilation", type=int, default=1)
ap.add_argument("--dtype", default="bf16")
a = ap.parse_args()

dev = torch.device("cuda")
torch.manual_seed(0)
width = 12 * 2 ** (a.level - 1)
R = a.res // 2 ** (a.level - 1)
amp = a.dtype == "bf16"
if a.unit == "reparam":
    from lmnet_b200.model import ReparamConv

    m = ReparamConv(width, 2 * width, width).to(dev).train()
    x = torch.randn(a.batch, width, R, R, device=dev, requires_grad=True)
elif a.unit == "conv":
    from lmnet_b200.conv3x3 import Conv3x3

    m = Conv3x3(2 * width, width, 3, 1, 1).to(dev).train()            # the decoder / skip shape: 2C -> C at this level
    x = torch.randn(a.batch, 2 * width, R, R, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_()
elif a.unit == "natt":
    from lmnet_b200.model import NeighborhoodTransformer

    m = NeighborhoodTransformer(width).to(dev).train()
    x = torch.randn(a.batch, width, R, R, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_()
else:
    import natten

    m = natten.NeighborhoodAttention2D(dim=width, num_heads=12, kernel_size=a.kernel, dilation=a.dilation).to(dev).train()
    x = torch.randn(a.batch, R, R, width, device=dev, requires_grad=True)
for _ in range(a.iters):
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        y = m(x)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("done", a.unit, "level", a.level, tuple(x.shape))
