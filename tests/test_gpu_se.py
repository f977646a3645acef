"""GPU parity of the fused squeeze-excite gate (lmnet_b200.se on csrc/se_gate.cu) against the stock SE sub-modules in
fp64 (reference: SE.forward, /root/reference/core/modules.py:1030-1036)."""
import copy

import pytest
import torch

from _helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,E", [(16, 24), (16, 48), (2, 96), (3, 192), (1, 8), (5, 20)])
def test_se_gate_forward_backward_vs_modules(B, E):
    from lmnet_b200 import _lib
    from lmnet_b200.model import SE
    from lmnet_b200.se import se_gate

    torch.manual_seed(E)
    se = SE(E)
    with torch.no_grad():
        se.fc2.bias.add_(torch.linspace(-4, 4, E))          # put some channels on both flat parts of the Hardsigmoid
    ref = copy.deepcopy(se).double()
    pool = torch.randn(B, E)
    go = torch.randn(B, E)
    pr = pool.double().requires_grad_()
    gr = ref.scale_activation(ref.fc2(ref.activation(ref.fc1(pr.view(B, E, 1, 1))))).reshape(B, E)
    (gr * go.double()).sum().backward()

    se = se.cuda()
    pc = pool.cuda().requires_grad_()
    before = _lib.launch_count()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        g = se_gate(se, pc)
    assert _lib.launch_count() == before + 1 and g.dtype == torch.float32
    (g * go.cuda()).sum().backward()
    assert _lib.launch_count() == before + 2
    assert rel_err(g.cpu(), gr) < 1e-5
    assert rel_err(pc.grad.cpu(), pr.grad) < 1e-5
    for m, r in ((se.fc1, ref.fc1), (se.fc2, ref.fc2)):
        assert rel_err(m.weight.grad.cpu(), r.weight.grad) < 1e-5
        assert rel_err(m.bias.grad.cpu(), r.bias.grad) < 1e-5


@pytest.mark.parametrize("shape,f", [((2, 12, 64, 96), 16), ((2, 24, 32, 48), 8), ((3, 48, 12, 20), 4), ((2, 96, 6, 10), 2), ((1, 12, 352, 352), 16)])
def test_integer_factor_avgpool_channels_last_vs_adaptive_avg_pool(shape, f):
    """lmnet_b200.pool (csrc/pool_nhwc.cu) == F.adaptive_avg_pool2d in fp64 on the same bf16 input, forward and backward
    (reference: PyramidPool.forward, /root/reference/core/modules.py:481-498)."""
    import torch.nn.functional as F

    from lmnet_b200 import _lib
    from lmnet_b200.pool import adaptive_avg_pool

    B, C, H, W = shape
    torch.manual_seed(f)
    x = torch.randn(shape).to(torch.bfloat16)
    go = torch.randn(B, C, H // f, W // f).to(torch.bfloat16)
    xr = x.double().requires_grad_()
    yr = F.adaptive_avg_pool2d(xr, (H // f, W // f))
    yr.backward(go.double())
    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    before = _lib.launch_count()
    y = adaptive_avg_pool(xc, (H // f, W // f))
    y.backward(go.cuda())
    assert _lib.launch_count() == before + 2
    assert y.is_contiguous(memory_format=torch.channels_last) and xc.grad.is_contiguous(memory_format=torch.channels_last)
    assert rel_err(y.float().cpu(), yr) < 4e-3
    assert rel_err(xc.grad.float().cpu(), xr.grad) < 4e-3
    # shapes the kernel does not cover take the stock path
    odd = torch.randn(2, 12, 30, 50, device="cuda").to(torch.bfloat16)
    assert torch.equal(adaptive_avg_pool(odd, (7, 7)), F.adaptive_avg_pool2d(odd, (7, 7)))
