"""GPU parity of the fused training loss (csrc/seg_loss.cu through the C ABI) against the reference's op sequence
(nn.CrossEntropyLoss(weight, label_smoothing) + DiceLoss, utils/train_eval_utils.py:141-142) evaluated in fp64 on the CPU."""
import pytest
import torch
from torch import nn

from _helpers import rel_err

pytestmark = pytest.mark.gpu


def _reference(logits64, labels, ce_w, eps, dice_w):
    from lmnet_b200.train import DiceLoss

    C = logits64.shape[1]
    ce = nn.CrossEntropyLoss(weight=torch.tensor(ce_w, dtype=torch.float64), label_smoothing=eps)
    return ce(logits64, labels) + DiceLoss(C)(logits64, labels.unsqueeze(1).double(), weight=list(dice_w))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape,ce_w,dice_w", [
    ((2, 2, 33, 35), (1.0, 4.0), (1.0, 4.0)),          # the LM-Net configuration, ragged size
    ((16, 2, 352, 352), (1.0, 4.0), (1.0, 4.0)),       # the BASELINE workload
    ((3, 3, 17, 9), (0.5, 2.0, 1.0), (1.0, 1.0, 3.0)),
    ((1, 4, 40, 24), (1.0, 1.0, 1.0, 1.0), (2.0, 1.0, 0.5, 1.0)),
    ((2, 8, 8, 8), tuple(0.5 + 0.25 * i for i in range(8)), tuple(1.0 + 0.5 * (i % 3) for i in range(8))),
])
def test_seg_loss_vs_reference_ops(shape, ce_w, dice_w, dtype):
    from lmnet_b200.segloss import seg_loss

    eps = 1e-3
    g = torch.Generator().manual_seed(3)
    B, C = shape[:2]
    logits = (2.0 * torch.randn(*shape, generator=g)).to(dtype)
    labels = torch.randint(0, C, (B,) + shape[2:], generator=g)
    ref_in = logits.double().requires_grad_()
    ref = _reference(ref_in, labels, ce_w, eps, dice_w)
    (3.0 * ref).backward()
    x = logits.cuda().requires_grad_()
    out = seg_loss(x, labels.cuda(), torch.tensor(ce_w), eps, dice_w)
    (3.0 * out).backward()
    assert out.dtype == torch.float32 and out.dim() == 0
    assert abs(float(out) - float(ref)) < 2e-5 * max(1.0, abs(float(ref)))
    tol = 1e-4 if dtype == torch.float32 else 1e-2           # 16-bit: the gradient is rounded to the storage type
    assert rel_err(x.grad.float().cpu(), ref_in.grad) < tol


def test_loss_fn_routes_the_reference_criteria_through_the_fused_op():
    """train.loss_fn with the reference's pair of criteria == the stock op sequence, value and gradient; and it is
    deterministic (fixed-order reductions)."""
    from lmnet_b200.train import DiceLoss, WeightedSmoothedCE, loss_fn

    g = torch.Generator().manual_seed(9)
    logits = torch.randn(4, 2, 64, 48, generator=g).cuda()
    labels = torch.randint(0, 2, (4, 64, 48), generator=g).cuda()
    w = torch.tensor([1.0, 4.0], device="cuda")
    dice = DiceLoss(2).cuda()
    for crit in (WeightedSmoothedCE(w, 1e-3).cuda(), nn.CrossEntropyLoss(weight=w, label_smoothing=1e-3)):
        a = logits.clone().requires_grad_()
        b = logits.clone().requires_grad_()
        fused = loss_fn(a, labels, crit, dice)
        stock = crit(b, labels) + dice(b, labels.unsqueeze(1).float(), weight=[1.0, 4.0])
        fused.backward()
        stock.backward()
        assert abs(float(fused) - float(stock)) < 1e-5
        assert rel_err(a.grad, b.grad) < 1e-4
        again = loss_fn(logits, labels, crit, dice)
        assert float(again) == float(fused)
