"""Generates the committed golden vectors FROM THE UNMODIFIED REFERENCE (run in the build container,
where /root/reference exists):

    python tests/golden/make_golden.py

* lmnet_keys.json     — state_dict keys and shapes of reference core.LM_Net.LM_Net(3, 2)
* reparam_golden.pt   — reference ReparamConv(6, 8, 4, 5, 3) in fp64: train forward/backward (incl. the
                        BatchNorm running-stat updates), eval forward, deploy forward
* lmnet_golden.pt     — reference LM_Net forward logits + per-parameter gradient norms, fp64, B=1, 32x32,
                        with natten's arithmetic supplied by the CPU oracle (natten is absent offline)
Weights come from tests/_helpers.fill_deterministic, so no checkpoint has to be stored.
The reference has no golden vectors of its own; these are self-generated.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200"), os.path.join(ROOT, "tests")]

from _helpers import fill_deterministic, import_reference  # noqa: E402
from oracle.lmnet_ref import to_oracle  # noqa: E402


def main():
    lm, mods = import_reference()
    torch.manual_seed(0)

    net = lm.LM_Net(3, 2)
    keys = {k: list(v.shape) for k, v in net.state_dict().items()}
    json.dump(keys, open(os.path.join(HERE, "lmnet_keys.json"), "w"), indent=0)

    # ---- ReparamConv ----
    g = torch.Generator().manual_seed(1)
    blk = mods.ReparamConv(6, 8, 4, 5, 3).double()
    fill_deterministic(blk, seed=7)
    x = torch.randn(2, 6, 9, 11, generator=g, dtype=torch.float64, requires_grad=True)
    go = torch.randn(2, 4, 9, 11, generator=g, dtype=torch.float64)
    blk.train()
    y = blk(x)
    y.backward(go)
    rec = {"x": x.detach(), "go": go, "train_out": y.detach(), "dx": x.grad.clone(),
           "grads": {k: p.grad.clone() for k, p in blk.named_parameters()},
           "buffers_after": {k: b.clone() for k, b in blk.named_buffers()}}
    blk.eval()
    with torch.no_grad():
        rec["eval_out"] = blk(x.detach())
        blk.switch_to_deploy()
        rec["deploy_out"] = blk(x.detach())
        rec["deploy_weight"] = blk.fuse_conv.weight.clone()
        rec["deploy_bias"] = blk.fuse_conv.bias.clone()
    torch.save(rec, os.path.join(HERE, "reparam_golden.pt"))

    # ---- whole model ----
    net = to_oracle(net, swap_reparam=False).double()   # reference ReparamConv.forward stays as is
    fill_deterministic(net, seed=3)
    g = torch.Generator().manual_seed(2)
    img = torch.randn(1, 3, 32, 32, generator=g, dtype=torch.float64)
    net.eval()
    with torch.no_grad():
        eval_logits = net(img)
    net.train()
    for m in net.modules():   # dropout off: keeps the golden independent of RNG streams
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    logits = net(img)
    target = (torch.rand(1, 32, 32, generator=g) > 0.7).long()
    loss = torch.nn.functional.cross_entropy(logits, target)
    loss.backward()
    rec = {"img": img, "target": target, "eval_logits": eval_logits, "train_logits": logits.detach(),
           "loss": loss.detach(), "grad_norms": {k: p.grad.norm() for k, p in net.named_parameters()},
           "grad_rpb": {k: p.grad.clone() for k, p in net.named_parameters() if k.endswith("rpb")}}
    torch.save(rec, os.path.join(HERE, "lmnet_golden.pt"))
    print("golden written:", os.listdir(HERE))


if __name__ == "__main__":
    main()
