"""GPU parity for the bilinear x2 (align_corners=True) kernels vs F.interpolate in fp64."""
import pytest
import torch
import torch.nn.functional as F

from _helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("shape", [(2, 3, 5, 7), (1, 12, 22, 22), (2, 4, 1, 9), (1, 2, 33, 16)])
def test_upsample2x_fwd_bwd(shape, dtype, tol):
    from lmnet_b200.upsample import Upsample2x

    g = torch.Generator().manual_seed(0)
    x = torch.randn(*shape, generator=g).to(dtype)
    go = torch.randn(shape[0], shape[1], 2 * shape[2], 2 * shape[3], generator=g).to(dtype)
    xr = x.double().requires_grad_()
    yr = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
    yr.backward(go.double())
    xc = x.cuda().requires_grad_()
    y = Upsample2x()(xc)
    assert y.dtype == dtype and y.shape == yr.shape
    y.backward(go.cuda())
    assert rel_err(y.float().cpu(), yr) < tol
    assert rel_err(xc.grad.float().cpu(), xr.grad) < tol


def test_upsample2x_matches_stock_op_bitwise_in_fp32_forward():
    from lmnet_b200.upsample import Upsample2x

    x = torch.randn(2, 6, 44, 44, device="cuda")
    ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    assert rel_err(Upsample2x()(x), ref) < 1e-6
