"""GPU parity for the fused BatchNorm2d + activation kernels (csrc/bn_act.cu) against stock
nn.BatchNorm2d + activation evaluated in fp64 on the CPU (the reference's op pair,
/root/reference/core/modules.py:537-539 and :97-100)."""
import copy

import pytest
import torch
import torch.nn.functional as F

from _helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}
ACTS = {"none": lambda t: t, "hardswish": F.hardswish, "gelu": F.gelu, "relu": F.relu}


@pytest.mark.parametrize("channels_last", [False, True], ids=["nchw", "nhwc"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", ["hardswish", "gelu", "none", "relu"])
@pytest.mark.parametrize("shape", [(4, 24, 40, 36), (2, 7, 9, 11), (3, 192, 11, 11), (1, 5, 3, 1000), (2, 12, 16, 20),
                                   (2, 96, 12, 12), (3, 48, 10, 8)])
def test_bn_act_train_eval(shape, act, dtype, channels_last):
    """channels_last=True feeds channels-last tensors: shapes whose element count is a multiple of the 16-byte vector run
    the channels-last kernels (output / gradient stay channels-last), the others fall back to the NCHW kernels."""
    from lmnet_b200.bnact import _cl_ok, bn_act

    B, C, H, W = shape
    g = torch.Generator().manual_seed(3)
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(C, generator=g))
        bn.bias.copy_(0.3 * torch.randn(C, generator=g))
        bn.running_mean.copy_(0.1 * torch.randn(C, generator=g))
        bn.running_var.copy_(0.5 + torch.rand(C, generator=g))
    ref = copy.deepcopy(bn).double().train()
    y = (1.5 * torch.randn(B, C, H, W, generator=g) + 0.3).to(dtype)
    go = torch.randn(B, C, H, W, generator=g).to(dtype)
    yr = y.double().requires_grad_()
    pre = ref(yr)
    # the activation sees the BatchNorm output rounded to the storage type (as in the reference's separate
    # kernels); straight-through rounding keeps the oracle differentiable
    pre_q = pre + (pre.detach().to(dtype).double() - pre.detach())
    outr = ACTS[act](pre_q)
    outr.backward(go.double())
    # hardswish' jumps at +-3: elements whose pre-activation sits within rounding distance of a kink may
    # legitimately take the other branch, so they are left out of the gradient comparison
    near_kink = ((pre.detach().abs() - 3).abs() < 0.05) if act == "hardswish" else torch.zeros_like(pre, dtype=torch.bool)

    bn = bn.cuda().train()
    yc = y.cuda()
    if channels_last:
        yc = yc.contiguous(memory_format=torch.channels_last)
    yc = yc.requires_grad_()
    out = bn_act(bn, yc, act)
    assert out.dtype == dtype
    if channels_last and _cl_ok(yc):
        assert out.is_contiguous(memory_format=torch.channels_last)
    out.backward(go.cuda())
    tol = TOL[dtype]
    assert rel_err(out.float().cpu(), outr) < tol
    keep = ~near_kink
    assert rel_err(yc.grad.float().cpu()[keep], yr.grad[keep]) < 2 * tol
    assert near_kink.float().mean() < 0.05
    # the per-channel sums include the near-kink elements: allow for a few branch flips in bf16 hardswish
    ptol = 0.2 if (act == "hardswish" and dtype == torch.bfloat16) else 2 * tol
    assert rel_err(bn.weight.grad.cpu(), ref.weight.grad) < ptol
    assert rel_err(bn.bias.grad.cpu(), ref.bias.grad) < ptol
    assert torch.allclose(bn.running_mean.cpu().double(), ref.running_mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(bn.running_var.cpu().double(), ref.running_var, rtol=1e-4, atol=1e-5)
    assert int(bn.num_batches_tracked) == 1
    bn.eval(), ref.eval()
    with torch.no_grad():
        assert rel_err(bn_act(bn, y.cuda(), act).float().cpu(), ACTS[act](ref(y.double()))) < tol
    with pytest.raises(NotImplementedError):
        bn_act(bn, yc, act)


def test_full_size_property_zero_mean_unit_variance():
    """At the BASELINE size ([16,24,352,352] bf16) the normalised output has per-channel mean beta and
    standard deviation |gamma| (act = none) — a size-independent check of the statistics pass."""
    from lmnet_b200.bnact import bn_act

    bn = torch.nn.BatchNorm2d(24).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-1, 1)
    y = (torch.randn(16, 24, 352, 352, device="cuda") * 3 + 1).to(torch.bfloat16)
    with torch.no_grad():
        out = bn_act(bn, y, "none").float()
    assert torch.allclose(out.mean(dim=(0, 2, 3)), bn.bias, atol=2e-2)
    assert torch.allclose(out.std(dim=(0, 2, 3)), bn.weight.abs(), rtol=2e-2)
