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


@pytest.mark.parametrize("case", [(2, 40, 36, 12, 12, 1, 24, 12), (2, 24, 24, 24, 24, 1, 48, 0), (2, 33, 47, 12, 24, 2, 48, 24),
                                  (1, 20, 28, 48, 48, 1, 144, 48), (2, 16, 16, 12, 12, 1, 36, 12)],
                         ids=lambda s: "x".join(map(str, s)))
def test_conv3x3_backward_reads_concatenation_gradient_slices_in_place(case):
    """The output of a convolution that goes into torch.cat receives its gradient as a channel slice of the concatenation's
    channels-last gradient (pixel pitch = total channels).  The kernels read that slice in place; the result must equal the
    backward on a densified copy bit for bit, and no densifying copy may be launched."""
    from lmnet_b200.conv3x3 import Conv3x3, _pixel_pitch

    B, H, W, Cin, Cout, s, Ctot, off = case
    torch.manual_seed(6)
    conv = Conv3x3(Cin, Cout, 3, s, 1).cuda()
    x = torch.randn(B, Cin, H, W, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    wide = torch.randn(B, Ctot, Ho, Wo, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    sl = wide[:, off:off + Cout]                                        # what torch.cat's backward produces
    assert _pixel_pitch(sl) == Ctot

    def run(go):
        xc = x.clone().requires_grad_()
        conv.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = conv(xc)
        y.backward(go)
        return xc.grad, conv.weight.grad.clone(), conv.bias.grad.clone()

    a = run(sl)
    b = run(sl.contiguous(memory_format=torch.channels_last))
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    # through autograd's own cat: the gradients of both producers match the stock module
    conv2 = Conv3x3(Cin, Ctot - Cout, 3, s, 1).cuda() if Ctot > Cout else None
    if conv2 is not None:
        xc = x.clone().requires_grad_()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            z = torch.cat([conv(xc), conv2(xc)], 1)
        conv.zero_grad(set_to_none=True)
        z.backward(wide)
        g_own = conv.weight.grad.clone()
        xs = x.clone().requires_grad_()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            zs = torch.cat([F.conv2d(xs, conv.weight, conv.bias, s, 1), F.conv2d(xs, conv2.weight, conv2.bias, s, 1)], 1)
        conv.zero_grad(set_to_none=True)
        zs.backward(wide)
        assert rel_err(g_own, conv.weight.grad) < 2e-2
        assert rel_err(xc.grad.float(), xs.grad.float()) < 2e-2
