"""GPU parity for nn.Linear over channels-last pixels (lmnet_b200.linear: pixel-GEMM forward / input gradient +
mixed-layout weight-gradient kernel) against the stock module evaluated in fp64 on the same bf16-rounded inputs.
Reference call sites: qkv / proj of natten's NeighborhoodAttention2D (/root/reference/core/modules.py:509, 517) and
fc1 / fc2 of Mlp (/root/reference/core/modules.py:42-56)."""
import copy

import pytest
import torch

from _helpers import rel_err

pytestmark = pytest.mark.gpu

# (B, H, W, K, N): qkv / proj / fc1 / fc2 of the four NeighborhoodTransformer levels (N > 96 runs as grid.z chunks)
CASES = [(2, 40, 36, 12, 36), (2, 40, 36, 12, 12), (2, 24, 24, 12, 24), (2, 24, 24, 24, 12), (2, 20, 16, 24, 72),
         (2, 12, 12, 48, 144), (2, 12, 12, 48, 96), (2, 8, 8, 96, 288), (2, 8, 8, 96, 192), (2, 8, 8, 192, 96),
         (1, 352, 352, 12, 36)]


@pytest.mark.parametrize("case", CASES, ids=lambda s: "x".join(map(str, s)))
def test_linear_forward_backward_vs_module(case):
    from lmnet_b200 import _lib
    from lmnet_b200.linear import linear, linear_supported

    B, H, W, K, N = case
    torch.manual_seed(3)
    lin = torch.nn.Linear(K, N)
    ref = copy.deepcopy(lin).double()
    x = torch.randn(B, H, W, K).to(torch.bfloat16)
    go = torch.randn(B, H, W, N)
    xr = x.double().requires_grad_()
    yr = ref(xr)
    (yr * go.double()).sum().backward()

    lin = lin.cuda()
    xc = x.cuda().requires_grad_()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert linear_supported(xc, lin)
        before = _lib.launch_count()
        y = linear(lin, xc)
        assert _lib.launch_count() > before
    assert y.dtype == torch.bfloat16 and y.shape == (B, H, W, N)
    (y.float() * go.cuda()).sum().backward()
    assert rel_err(y.float().cpu(), yr) < 6e-3
    assert rel_err(xc.grad.float().cpu(), xr.grad) < 1e-2
    assert lin.weight.grad.dtype == torch.float32
    assert rel_err(lin.weight.grad.cpu(), ref.weight.grad) < 1e-2        # dout is rounded to bf16 first
    assert rel_err(lin.bias.grad.cpu(), ref.bias.grad) < 1e-2


def test_linear_falls_back_for_fp32_and_odd_shapes():
    from lmnet_b200.linear import linear

    lin = torch.nn.Linear(12, 36).cuda()
    x = torch.randn(2, 8, 8, 12, device="cuda")
    assert torch.equal(linear(lin, x), lin(x))                            # fp32 storage: stock module
    lin7 = torch.nn.Linear(7, 5).cuda()
    x7 = torch.randn(2, 8, 8, 7, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert torch.equal(linear(lin7, x7), lin7(x7))
