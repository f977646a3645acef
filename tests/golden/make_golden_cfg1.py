"""Golden vectors at BASELINE.json configs[0] size — LM-Net, batch 2, 256x256, fp32 inputs — FROM THE UNMODIFIED
REFERENCE model (/root/reference/core/LM_Net.py + core/modules.py imported as they are, natten's arithmetic supplied
by the CPU oracle), computed in fp64.  Run in the build container (where /root/reference exists):

    python tests/golden/make_golden_cfg1.py        ->  tests/golden/lmnet_cfg1_golden.npz   (~2 MB)

Contents: eval-mode logits, train-mode logits (batch statistics, dropout off), the training loss of
train.py:157-158 (CE weight [1,4] label-smoothing 1e-3 + Dice weight [1,4]), 514 gradient norms, a checksum of the
seeded input.  The input itself is regenerated from its seed (lmnet_b200.train.synthetic_batches(1, 2, 256, seed=0)).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200"), os.path.join(ROOT, "tests")]

from _helpers import fill_deterministic, import_reference  # noqa: E402
from lmnet_b200.train import DiceLoss, synthetic_batches  # noqa: E402
from oracle.lmnet_ref import to_oracle  # noqa: E402


def main():
    lm, _ = import_reference()
    torch.manual_seed(0)
    net = to_oracle(lm.LM_Net(3, 2), swap_reparam=False).double()      # reference ReparamConv.forward stays as is
    fill_deterministic(net, seed=3)
    for m in net.modules():                                            # dropout off: golden independent of RNG streams
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    images, masks = synthetic_batches(1, 2, 256, seed=0, pin=False)[0]
    img = images.double()
    net.eval()
    with torch.no_grad():
        eval_logits = net(img)
    net.train()
    logits = net(img)
    ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([1.0, 4.0], dtype=torch.float64), label_smoothing=1e-3)
    loss = ce(logits, masks) + DiceLoss(2)(logits, masks.unsqueeze(1).double(), weight=[1.0, 4.0])
    loss.backward()
    names = [k for k, _ in net.named_parameters()]
    norms = np.array([float(p.grad.norm()) for _, p in net.named_parameters()])
    np.savez_compressed(os.path.join(HERE, "lmnet_cfg1_golden.npz"),
                        eval_logits=eval_logits.float().numpy(), train_logits=logits.detach().float().numpy(),
                        loss=np.float64(loss.detach()), grad_names=np.array(names), grad_norms=norms,
                        image_checksum=np.float64(images.double().abs().sum()), mask_sum=np.int64(masks.sum()))
    marg = (eval_logits[:, 0] - eval_logits[:, 1]).abs()
    print("cfg1 golden written; loss", float(loss), "min |margin| eval", float(marg.min()),
          "pixels with margin < 1e-3:", int((marg < 1e-3).sum()))


if __name__ == "__main__":
    main()
