"""debug: runs the fused depthwise op forward then backward on a tiny TMA-eligible shape with explicit syncs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200")]
import torch
from lmnet_b200.model import ReparamConv
from lmnet_b200.reparam import fused_dw_bn_gelu
torch.manual_seed(0)
B, E, H, W = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (2, 8, 40, 72))]
m = ReparamConv(4, E, 4).cuda().train()
x1 = torch.randn(B, E, H, W, device="cuda").to(torch.bfloat16).requires_grad_()
print("forward...", flush=True)
z, p = fused_dw_bn_gelu(m, x1)
torch.cuda.synchronize()
print("forward ok", float(z.float().abs().mean()), flush=True)
print("backward...", flush=True)
torch.autograd.backward([z, p], [torch.randn_like(z), torch.randn_like(p)])
torch.cuda.synchronize()
print("backward ok", float(x1.grad.float().abs().mean()), flush=True)
