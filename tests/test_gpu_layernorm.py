"""GPU parity for the short-row LayerNorm kernels (csrc/layer_norm.cu) vs nn.LayerNorm in fp64."""
import copy

import pytest
import torch

from _helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C", [12, 24, 48, 96])
@pytest.mark.parametrize("lead", [(2, 13, 17), (1, 1, 1), (3, 64, 65)])
def test_layer_norm_forward_backward(lead, C, dtype):
    from lmnet_b200.layernorm import layer_norm

    g = torch.Generator().manual_seed(1)
    ln = torch.nn.LayerNorm(C)
    with torch.no_grad():
        ln.weight.copy_(1 + 0.3 * torch.randn(C, generator=g))
        ln.bias.copy_(0.2 * torch.randn(C, generator=g))
    ref = copy.deepcopy(ln).double()
    x = (2 * torch.randn(*lead, C, generator=g) + 0.5).to(dtype)
    go = torch.randn(*lead, C, generator=g).to(dtype)
    xr = x.double().requires_grad_()
    yr = ref(xr)
    yr.backward(go.double())
    ln = ln.cuda()
    xc = x.cuda().requires_grad_()
    y = layer_norm(ln, xc)
    assert y.dtype == dtype and y.shape == x.shape
    y.backward(go.cuda())
    tol = TOL[dtype]
    assert rel_err(y.float().cpu(), yr) < tol
    assert rel_err(xc.grad.float().cpu(), xr.grad) < 2 * tol
    assert rel_err(ln.weight.grad.cpu(), ref.weight.grad) < 2 * tol
    assert rel_err(ln.bias.grad.cpu(), ref.bias.grad) < 2 * tol


def test_unsupported_width_uses_the_stock_module():
    from lmnet_b200.layernorm import layer_norm, supported

    assert not supported(372)
    ln = torch.nn.LayerNorm(372).cuda()
    x = torch.randn(2, 5, 372, device="cuda")
    assert torch.equal(layer_norm(ln, x), ln(x))


def test_full_size_rows_are_normalised():
    """BASELINE size: 16*352*352 rows of 12 channels in bf16 — every output row has mean beta-weighted 0 and
    unit variance when gamma = 1, beta = 0."""
    from lmnet_b200.layernorm import layer_norm

    ln = torch.nn.LayerNorm(12).cuda()
    x = (torch.randn(16, 352, 352, 12, device="cuda") * 5 + 2).to(torch.bfloat16)
    with torch.no_grad():
        y = layer_norm(ln, x).float()
    assert float(y.mean(-1).abs().max()) < 2e-2
    assert float((y.var(-1, unbiased=False) - 1).abs().max()) < 5e-2
