"""Finds which part of the training step (if any) cannot be captured into a CUDA graph."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200")]
import torch
from lmnet_b200.model import LM_Net, ReparamConv, NeighborhoodTransformer
from lmnet_b200.train import build_training, loss_fn, synthetic_batches

dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True


def try_capture(name, fn, warm=3):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        print(f"[ok]   {name}")
    except Exception as e:
        print(f"[FAIL] {name}: {type(e).__name__}: {str(e).splitlines()[0]}")
        torch.cuda.synchronize()


x = torch.randn(2, 12, 64, 64, device=dev, requires_grad=True)
blk = ReparamConv(12, 24, 12).to(dev).train()
nt = NeighborhoodTransformer(12).to(dev).train()


def fwd_bwd(m, inp):
    def f():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(inp)
        m.zero_grad(set_to_none=True)
        y.float().sum().backward()
    return f


def fwd_only(m, inp):
    def f():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            m(inp)
    return f

which = sys.argv[1] if len(sys.argv) > 1 else "all"
tests = {
    "reparam_fwd": lambda: try_capture("ReparamConv forward (no grad)", fwd_only(blk, x)),
    "reparam": lambda: try_capture("ReparamConv fwd+bwd", fwd_bwd(blk, x)),
    "natt_fwd": lambda: try_capture("NeighborhoodTransformer forward (no grad)", fwd_only(nt, x)),
    "natt": lambda: try_capture("NeighborhoodTransformer fwd+bwd", fwd_bwd(nt, x)),
}
net = LM_Net(3, 2).to(dev).train()
opt, crit, dice = build_training(net, dev, capturable=True)
img, msk = (t.to(dev) for t in synthetic_batches(1, 2, 64, pin=False)[0])


def model_fwd_bwd():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = loss_fn(net(img), msk, crit, dice)
    opt.zero_grad(set_to_none=True)
    loss.backward()


def full_step():
    model_fwd_bwd()
    opt.step()

def part(name, f):
    tests[name] = lambda: try_capture(name, f)

f12 = torch.randn(2, 12, 64, 64, device=dev, requires_grad=True)
f24 = torch.randn(2, 24, 32, 32, device=dev, requires_grad=True)
f48 = torch.randn(2, 48, 16, 16, device=dev, requires_grad=True)
f96 = torch.randn(2, 96, 8, 8, device=dev, requires_grad=True)
f192 = torch.randn(2, 192, 4, 4, device=dev, requires_grad=True)
logits = torch.randn(2, 2, 64, 64, device=dev, requires_grad=True)

def run(mod_fn):
    def f():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = mod_fn()
        net.zero_grad(set_to_none=True)
        y.float().sum().backward()
    return f

part("gft", run(lambda: net.gft(net.pyramidpool(f12, f24, f48, f96, f192))))
part("skip2", run(lambda: net.skip2(f24, f48, f96)))
part("skip1", run(lambda: net.skip1(f48, f96)))
part("up1", run(lambda: net.up1(f192)))
part("down", run(lambda: net.down1(f12)))
part("loss", run(lambda: loss_fn(logits, msk, crit, dice)))
part("ce", run(lambda: crit(logits, msk)))
part("dice", run(lambda: dice(logits, msk.unsqueeze(1).float(), weight=[1.0, 4.0])))
tests["model"] = lambda: try_capture("LM_Net fwd+loss+bwd", model_fwd_bwd)
tests["step"] = lambda: try_capture("full step incl. AdamW", full_step)
for k, t in tests.items():
    if which in ("all", k):
        t()
