"""GPU parity for the dense 3x3 convolutions (lmnet_b200.conv3x3 on csrc/conv3x3.cu) against F.conv2d evaluated in fp64
on the same bf16-rounded operands.  Reference layers: down1-4 / up1-4 (/root/reference/core/LM_Net.py:14-39, 58-74),
M2Skip / M3Skip (/root/reference/core/modules.py:83-143), OverlapPatchEmbed (/root/reference/core/modules.py:30-39)."""
import copy

import pytest
import torch
import torch.nn.functional as F

from _helpers import rel_err

pytestmark = pytest.mark.gpu

# (B, H, W, Cin, Cout, stride): every (Cin, Cout, stride) LM-Net runs on its two largest resolutions, ragged / odd sizes
# (partial tiles, odd stride-2 inputs), a one-tile image, and one full-size level-1 layer
CASES = [(2, 40, 36, 12, 12, 1), (2, 40, 36, 24, 12, 1), (2, 33, 47, 12, 24, 2), (2, 24, 24, 24, 24, 1), (3, 17, 31, 48, 24, 1),
         (2, 24, 24, 72, 24, 1), (2, 24, 24, 24, 48, 2), (2, 12, 12, 48, 48, 1), (1, 5, 7, 12, 12, 1), (2, 16, 16, 12, 24, 2),
         (2, 18, 34, 48, 96, 2), (1, 352, 352, 24, 12, 1), (1, 352, 352, 12, 24, 2)]


@pytest.mark.parametrize("case", CASES, ids=lambda s: "x".join(map(str, s)))
def test_conv3x3_forward_backward_vs_fp64(case):
    from lmnet_b200 import _lib
    from lmnet_b200.conv3x3 import Conv3x3, fwd_supported, wgrad_supported

    B, H, W, Cin, Cout, s = case
    torch.manual_seed(4)
    conv = Conv3x3(Cin, Cout, 3, s, 1)
    with torch.no_grad():
        conv.weight.copy_(conv.weight.to(torch.bfloat16).float())       # bf16-exact weights: the comparison isolates the kernel
    ref = copy.deepcopy(conv).double()
    x = torch.randn(B, Cin, H, W).to(torch.bfloat16)
    xr = x.double().requires_grad_()
    yr = F.conv2d(xr, ref.weight, ref.bias, s, 1)
    go = torch.randn_like(yr).to(torch.bfloat16)
    (yr * go.double()).sum().backward()

    sup = fwd_supported(B, H, W, Cin, Cout, s, torch.bfloat16)
    assert sup or Cout >= 96                                             # wide layers: weights exceed shared memory -> cuDNN
    conv = conv.cuda()
    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    before = _lib.launch_count()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = conv(xc)
    assert (_lib.launch_count() > before) == sup
    assert y.dtype == torch.bfloat16 and y.shape == yr.shape and y.is_contiguous(memory_format=torch.channels_last)
    y.backward(go.cuda())
    assert rel_err(y.float().cpu(), yr) < 6e-3                           # one bf16 rounding of the output
    assert rel_err(xc.grad.float().cpu(), xr.grad) < 6e-3
    if sup and wgrad_supported(B, H, W, Cin, Cout, s, torch.bfloat16):
        tol = 1e-5                                                       # exact bf16 products, fp32 accumulation
    else:
        tol = 1e-2                                                       # cuDNN returns the gradient in bf16
    assert conv.weight.grad.dtype == torch.float32
    assert rel_err(conv.weight.grad.cpu(), ref.weight.grad) < tol
    assert rel_err(conv.bias.grad.cpu(), ref.bias.grad) < tol


def test_conv3x3_fallbacks_match_the_stock_module():
    from lmnet_b200.conv3x3 import Conv3x3, convert_conv3x3

    torch.manual_seed(0)
    conv = Conv3x3(12, 12, 3, 1, 1).cuda()
    x = torch.randn(2, 12, 9, 9, device="cuda")
    assert torch.equal(conv(x), F.conv2d(x, conv.weight, conv.bias, 1, 1))           # fp32 storage: stock path
    wide = Conv3x3(372, 372, 3, 1, 1).cuda()                                         # GFT patch embedding: cuDNN
    xw = torch.randn(2, 372, 6, 6, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        a = wide(xw)
        b = F.conv2d(xw, wide.weight, wide.bias, 1, 1)
    assert rel_err(a.float(), b.float()) < 1e-2
    seq = torch.nn.Sequential(torch.nn.Conv2d(12, 24, 3, 2, 1), torch.nn.Conv2d(24, 24, 1), torch.nn.Conv2d(24, 24, 3, 1, 1, groups=24))
    keys = list(seq.state_dict().keys())
    convert_conv3x3(seq)
    assert type(seq[0]) is Conv3x3 and type(seq[1]) is torch.nn.Conv2d and type(seq[2]) is torch.nn.Conv2d
    assert list(seq.state_dict().keys()) == keys
