"""How far is the training step from being CPU-launch-bound?  Runs the step at batch 1 / 64x64 (GPU work tiny,
same number of launches): its wall time per step approximates the host cost of enqueueing a step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200")]
import torch
from lmnet_b200.model import LM_Net
from lmnet_b200.train import build_training, synthetic_batches, train_step
dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True
net = LM_Net(3, 2).to(dev).train()
opt, crit, dice = build_training(net, dev)
for B, R in ((1, 64), (16, 352)):
    img, msk = (t.to(dev) for t in synthetic_batches(1, B, R, pin=False)[0])
    for _ in range(5):
        train_step(net, opt, img, msk, crit, dice)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        train_step(net, opt, img, msk, crit, dice)
    t_enq = (time.perf_counter() - t0) / 10
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 10
    print(f"batch {B} res {R}: host enqueue {1e3 * t_enq:.1f} ms/step, wall {1e3 * t_all:.1f} ms/step")
x = torch.randn(2, 372, 22, 22, device=dev)
with torch.autocast("cuda", dtype=torch.bfloat16):
    y = net.gft.conv(x)
    print("gft.conv out dtype", y.dtype, "upsample out dtype", net.up1[0](y).dtype)
    y2 = net.gft(x)
    print("gft out dtype", y2.dtype)
