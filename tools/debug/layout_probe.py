import os, sys, traceback, collections
ROOT = os.getcwd()
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200")]
import torch
from lmnet_b200 import conv3x3 as c3
from lmnet_b200.model import LM_Net
from lmnet_b200.train import build_training, synthetic_batches, train_step
seen = collections.Counter()
orig = c3._cl
def probe(t):
    if t.dim() == 4 and not t.is_contiguous(memory_format=torch.channels_last):
        st = [f"{os.path.basename(f.filename)}:{f.lineno}:{f.name}" for f in traceback.extract_stack()[-9:-1] if "lmnet_b200" in f.filename or "natten" in f.filename]
        seen[(tuple(t.shape), tuple(t.stride()), " < ".join(reversed(st)))] += 1
    return orig(t)
c3._cl = probe
dev = torch.device("cuda")
net = LM_Net(3, 2).to(dev).train()
opt, crit, dice = build_training(net, dev)
img, msk = (t.to(dev) for t in synthetic_batches(1, 16, 352, pin=False)[0])
train_step(net, opt, img, msk, crit, dice)
torch.cuda.synchronize()
for k, n in seen.most_common(30):
    print(n, k)
