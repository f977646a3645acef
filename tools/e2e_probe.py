"""Where does the end-to-end loop lose time against the device-resident step?  Times train_one_epoch variants."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200")]
import torch
from lmnet_b200.model import LM_Net
from lmnet_b200.train import ConfusionMetrics, GraphedTrainStep, build_training, synthetic_batches, train_one_epoch

dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = LM_Net(3, 2).to(dev).train()
opt, crit, dice = build_training(net, dev, capturable=True)
host = synthetic_batches(2, 16, 352, seed=0)
res = [(i.to(dev), m.to(dev)) for i, m in host]
g = GraphedTrainStep(net, opt, crit, dice, *res[0], warmup=3)
N = 30
class Loader:
    def __init__(self, src, n): self.src, self.n = src, n
    def __iter__(self):
        for i in range(self.n): yield self.src[i % 2]
def run(name, src, **kw):
    m = ConfusionMetrics(2)
    train_one_epoch(net, opt, m, 2, Loader(src, 3), dev, crit, object(), dice, step_fn=g, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    train_one_epoch(net, opt, m, 2, Loader(src, N), dev, crit, object(), dice, step_fn=g, **kw)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / N * 1e3
    print(f"{name:55s} {dt:7.2f} ms/step")
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(N): g(*res[i % 2])
torch.cuda.synchronize(); print(f"{'graph replay, resident batches':55s} {(time.perf_counter()-t0)/N*1e3:7.2f} ms/step")
t0 = time.perf_counter()
for i in range(N): g.graph.replay()
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"{'host time of graph.replay() alone':55s} {(t1-t0)/N*1e3:7.2f} ms/call")
if os.environ.get("E2E_TRACE"):
    # per-step host timestamps of the first end-to-end epoch after the resident loop (bench.py's order of events)
    class TLoader(Loader):
        def __iter__(self):
            for i in range(self.n):
                stamps.append(time.perf_counter())
                yield self.src[i % 2]
    for rep in range(2):
        stamps = []
        m = ConfusionMetrics(2)
        train_one_epoch(net, opt, m, 2, Loader(host, 3), dev, crit, object(), dice, step_fn=g)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        train_one_epoch(net, opt, m, 2, TLoader(host, 10), dev, crit, object(), dice, step_fn=g)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"trace rep {rep}: 10 steps {1e3*(t2-t0):.1f} ms (host loop {1e3*(t1-t0):.1f}); loader calls at",
              " ".join(f"{1e3*(s-t0):.1f}" for s in stamps))
run("e2e default (pinned host, prefetch, device metrics)", host)
run("e2e, device-resident loader (no H2D)", res)
run("e2e, no metrics", host, ) if False else None
m = None
def run_nometric(name, src):
    train_one_epoch(net, opt, None, 2, Loader(src, 3), dev, crit, object(), dice, step_fn=g)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    train_one_epoch(net, opt, None, 2, Loader(src, N), dev, crit, object(), dice, step_fn=g)
    torch.cuda.synchronize(); print(f"{name:55s} {(time.perf_counter()-t0)/N*1e3:7.2f} ms/step")
run_nometric("e2e, pinned host, no metric update", host)
run("e2e, no prefetch", host, prefetch=False)
run("e2e, loss read immediately", host, defer_loss_read=False)
# raw H2D bandwidth
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(10):
    a = host[0][0].to(dev, non_blocking=True); b = host[0][1].to(dev, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
nb = host[0][0].numel() * 4 + host[0][1].numel() * 8
print(f"H2D of one batch ({nb/1e6:.1f} MB): {dt*1e3:.2f} ms = {nb/dt/1e9:.1f} GB/s; cpu count {os.cpu_count()}")
