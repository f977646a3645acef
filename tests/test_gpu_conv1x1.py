"""GPU parity for the 1x1-convolution path of ReparamConv (plane-wise GEMMs + csrc/wgrad_1x1.cu) against
stock nn.Conv2d modules evaluated in fp64 (reference ops: /root/reference/core/modules.py:587, 598-599)."""
import copy

import pytest
import torch

from _helpers import rel_err

pytestmark = pytest.mark.gpu


# (B, Cin, E, Cout, H, W): LM-Net's four levels, the 3-channel stem, and a ragged pixel count
SHAPES = [(2, 12, 24, 12, 24, 40), (2, 3, 24, 12, 16, 24), (2, 24, 48, 24, 20, 16), (2, 48, 96, 48, 12, 12),
          (2, 96, 192, 96, 8, 8), (3, 12, 24, 12, 17, 8)]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_raw_wgrad_kernel_vs_einsum(shape):
    from lmnet_b200.conv1x1 import _wgrad, wgrad_supported

    B, Cin, E, Cout, H, W = shape
    P = H * W
    assert wgrad_supported(B, Cout, E, Cin, P, torch.bfloat16)
    g = torch.Generator().manual_seed(0)
    A = torch.randn(B, Cout, P, generator=g).to(torch.bfloat16)
    B1 = torch.randn(B, E, P, generator=g).to(torch.bfloat16)
    B2 = torch.randn(B, Cin, P, generator=g).to(torch.bfloat16)
    dW, drow = _wgrad(A.cuda(), B1.cuda(), B2.cuda())
    ref = torch.einsum("bmp,bnp->bmn", A.double(), torch.cat([B1, B2], 1).double())
    assert rel_err(dW.cpu(), ref) < 1e-5          # bf16 products are exact in fp32; only summation order differs
    assert rel_err(drow.cpu(), A.double().sum(-1)) < 1e-5
    dW1, _ = _wgrad(A.cuda(), B2.cuda(), None)
    assert rel_err(dW1.cpu(), torch.einsum("bmp,bnp->bmn", A.double(), B2.double())) < 1e-5


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_expand_and_pointwise_shortcut_vs_modules(shape):
    from lmnet_b200.conv1x1 import expand_1x1, pointwise_shortcut

    B, Cin, E, Cout, H, W = shape
    torch.manual_seed(1)
    ex, pw, sc = torch.nn.Conv2d(Cin, E, 1), torch.nn.Conv2d(E, Cout, 1), torch.nn.Conv2d(Cin, Cout, 1)
    rex, rpw, rsc = (copy.deepcopy(m).double() for m in (ex, pw, sc))
    x = torch.randn(B, Cin, H, W)
    z = torch.randn(B, E, H, W)
    gate = torch.rand(B, E, 1, 1)
    go1, go2 = torch.randn(B, E, H, W), torch.randn(B, Cout, H, W)
    # fp64 reference on the bf16-rounded activations
    xr, zr, gr = (t.to(torch.bfloat16).double().requires_grad_() for t in (x, z, gate))
    y1 = rex(xr)
    y2 = rpw(gr * zr) + rsc(xr)
    (y1 * go1.double()).sum().add((y2 * go2.double()).sum()).backward()

    ex, pw, sc = ex.cuda(), pw.cuda(), sc.cuda()
    xc, zc, gc = (t.cuda().requires_grad_() for t in (x, z, gate))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o1 = expand_1x1(ex, xc)
        o2 = pointwise_shortcut(pw, sc, zc, gc, xc)
    assert o1.dtype == torch.bfloat16 and o2.dtype == torch.bfloat16
    (o1.float() * go1.cuda()).sum().add((o2.float() * go2.cuda()).sum()).backward()
    tol = 2e-2
    assert rel_err(o1.float().cpu(), y1) < tol and rel_err(o2.float().cpu(), y2) < tol
    assert rel_err(xc.grad.cpu(), xr.grad) < tol
    assert rel_err(zc.grad.cpu(), zr.grad) < tol
    assert rel_err(gc.grad.cpu(), gr.grad) < tol
    for m, r in ((ex, rex), (pw, rpw), (sc, rsc)):
        assert m.weight.grad.dtype == torch.float32
        assert rel_err(m.weight.grad.cpu(), r.weight.grad) < tol
        assert rel_err(m.bias.grad.cpu(), r.bias.grad) < tol


def test_fp32_inputs_take_the_module_path():
    from lmnet_b200.conv1x1 import expand_1x1

    conv = torch.nn.Conv2d(12, 24, 1).cuda()
    x = torch.randn(2, 12, 8, 8, device="cuda")
    assert torch.equal(expand_1x1(conv, x), conv(x))


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_raw_mixed_layout_wgrad_kernel_vs_einsum(shape):
    """csrc/wgrad_1x1.cu mixed-layout kernel: channels-last operands are read through ldmatrix.trans, no transposes."""
    from lmnet_b200.conv1x1 import _wgrad_cl, wgrad_cl_supported

    B, Cin, E, Cout, H, W = shape
    Cin = (Cin + 3) // 4 * 4                     # the host pads the 3-channel stem to 4
    P = H * W
    g = torch.Generator().manual_seed(0)
    dy = torch.randn(B, E, P, generator=g).to(torch.bfloat16)          # planes
    x = torch.randn(B, P, Cin, generator=g).to(torch.bfloat16)         # channels-last
    do = torch.randn(B, P, Cout, generator=g).to(torch.bfloat16)       # channels-last
    z = torch.randn(B, E, P, generator=g).to(torch.bfloat16)           # planes
    # expand: A planes, B1 channels-last
    assert wgrad_cl_supported(B, E, Cin, 0, P, False, True, torch.bfloat16)
    dW, drow = _wgrad_cl(dy.cuda(), x.cuda(), None, False, True)
    assert rel_err(dW.cpu(), torch.einsum("bmp,bpn->bmn", dy.double(), x.double())) < 1e-5
    assert rel_err(drow.cpu(), dy.double().sum(-1)) < 1e-5
    dWs, drows = _wgrad_cl(dy.cuda(), x.cuda(), None, False, True, batch_sum=True)       # summed over the batch in the reduction
    assert dWs.shape == (E, Cin) and drows.shape == (E,)
    assert rel_err(dWs.cpu(), torch.einsum("bmp,bpn->mn", dy.double(), x.double())) < 1e-5
    assert rel_err(drows.cpu(), dy.double().sum((0, 2))) < 1e-5
    # pointwise + shortcut: A channels-last, B1 planes, B2 channels-last
    assert wgrad_cl_supported(B, Cout, E, Cin, P, True, False, torch.bfloat16)
    dW, drow = _wgrad_cl(do.cuda(), z.cuda(), x.cuda(), True, False)
    ref = torch.cat([torch.einsum("bpm,bnp->bmn", do.double(), z.double()),
                     torch.einsum("bpm,bpn->bmn", do.double(), x.double())], dim=2)
    assert rel_err(dW.cpu(), ref) < 1e-5
    assert rel_err(drow.cpu(), do.double().sum(1)) < 1e-5


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_channels_last_expand_and_pointwise_shortcut_vs_modules(shape):
    """Same comparison as above with a CHANNELS-LAST block input: the expand output must come back as NCHW planes, the
    block output and the input gradient as channels-last, all without a layout copy of any activation."""
    from lmnet_b200.conv1x1 import expand_1x1, is_channels_last, pointwise_shortcut

    B, Cin, E, Cout, H, W = shape
    torch.manual_seed(1)
    ex, pw, sc = torch.nn.Conv2d(Cin, E, 1), torch.nn.Conv2d(E, Cout, 1), torch.nn.Conv2d(Cin, Cout, 1)
    rex, rpw, rsc = (copy.deepcopy(m).double() for m in (ex, pw, sc))
    x = torch.randn(B, Cin, H, W)
    z = torch.randn(B, E, H, W)
    gate = torch.rand(B, E, 1, 1)
    go1, go2 = torch.randn(B, E, H, W), torch.randn(B, Cout, H, W)
    xr, zr, gr = (t.to(torch.bfloat16).double().requires_grad_() for t in (x, z, gate))
    y1 = rex(xr)
    y2 = rpw(gr * zr) + rsc(xr)
    (y1 * go1.double()).sum().add((y2 * go2.double()).sum()).backward()

    ex, pw, sc = ex.cuda(), pw.cuda(), sc.cuda()
    xc = x.to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    zc, gc = (t.cuda().requires_grad_() for t in (z.to(torch.bfloat16), gate))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o1 = expand_1x1(ex, xc)
        o2 = pointwise_shortcut(pw, sc, zc, gc, xc)
    assert o1.is_contiguous() and o1.dtype == torch.bfloat16
    assert is_channels_last(o2) or Cout == 1
    (o1.float() * go1.cuda()).sum().add((o2.float() * go2.cuda()).sum()).backward()
    tol = 2e-2
    assert rel_err(o1.float().cpu(), y1) < tol and rel_err(o2.float().cpu(), y2) < tol
    assert rel_err(xc.grad.float().cpu(), xr.grad) < tol
    assert rel_err(zc.grad.float().cpu(), zr.grad) < tol
    assert rel_err(gc.grad.cpu(), gr.grad) < tol
    for m, r in ((ex, rex), (pw, rpw), (sc, rsc)):
        assert rel_err(m.weight.grad.cpu(), r.weight.grad) < tol
        assert rel_err(m.bias.grad.cpu(), r.bias.grad) < tol


# (B, P, N, K1, K2): the five roles of csrc/pixel_gemm.cu at LM-Net's channel counts, incl. a ragged last tile
PG_CASES = [(2, 960, 24, 12, 0), (2, 352, 24, 4, 0), (2, 1000, 48, 24, 0), (3, 136, 96, 48, 0), (2, 64, 192, 96, 0),
            (1, 123904, 24, 12, 0)]


@pytest.mark.parametrize("case", PG_CASES, ids=lambda s: "x".join(map(str, s)))
def test_raw_pixel_gemm_all_roles_vs_einsum(case):
    """out = W1.in1 (+ W2.in2) (+ bias) for every operand-layout combination the block uses; fp64 einsum of the same
    bf16-rounded operands, so the only error is the fp32 accumulation order and the final rounding to bf16."""
    from lmnet_b200.conv1x1 import pgemm_supported, pixel_gemm

    B, P, E, C, _ = case                       # E = expanded channels, C = block channels
    g = torch.Generator().manual_seed(5)
    bf = torch.bfloat16

    def rnd(*shape):
        return torch.randn(*shape, generator=g).to(bf)

    def close(got, want):
        assert rel_err(got.float().cpu(), want) < 6e-3      # one bf16 rounding of the output (2^-9) + accumulation order

    x_cl, dy_pl, z_pl, do_cl = rnd(B, P, C), rnd(B, E, P), rnd(B, E, P), rnd(B, P, C)
    # weights are fp32 parameters (read in place, rounded to bf16 by the kernel); the gate multiplies W_pw per image
    w_ex, w_pw, w_sc = (torch.randn(*sh, generator=g) for sh in ((E, C), (C, E), (C, C)))
    gate = torch.rand(B, E, generator=g)
    bias_e, bias_c = torch.randn(E, generator=g), torch.randn(C, generator=g)
    w_ex_r, w_sc_r = w_ex.to(bf).double(), w_sc.to(bf).double()                       # what the kernel contracts with
    w_pw_r = (w_pw.unsqueeze(0) * gate.unsqueeze(1)).to(bf).double()                  # [B, C, E]
    # expand forward: channels-last -> planes, with BatchNorm partial sums
    assert pgemm_supported(B, P, E, C, 0, True, False, True, bf)
    y, part = pixel_gemm(x_cl.cuda(), True, w_ex.cuda(), bias=bias_e.cuda(), out_cl=False, stats=True)
    want = torch.einsum("nk,bpk->bnp", w_ex_r, x_cl.double()) + bias_e.double().view(1, E, 1)
    close(y, want)
    yd = y.double().cpu()
    tot = part.double().sum(1).cpu()
    assert rel_err(tot[:, 0], yd.sum((0, 2))) < 1e-5 and rel_err(tot[:, 1], (yd * yd).sum((0, 2))) < 1e-5
    # expand input gradient: planes -> channels-last (the weight is read through a transposed view)
    if pgemm_supported(B, P, C, E, 0, False, True, False, bf):
        dx, _ = pixel_gemm(dy_pl.cuda(), False, w_ex.cuda().t(), out_cl=True)
        close(dx, torch.einsum("nk,bnp->bpk", w_ex_r, dy_pl.double()))
    # pointwise (gated weights, planes) + shortcut (channels-last) -> channels-last
    if pgemm_supported(B, P, C, E, C, False, True, False, bf):
        o, _ = pixel_gemm(z_pl.cuda(), False, w_pw.cuda(), x_cl.cuda(), w_sc.cuda(), bias_c.cuda(), out_cl=True, gate=gate.cuda())
        want = (torch.einsum("bne,bep->bpn", w_pw_r, z_pl.double()) +
                torch.einsum("nk,bpk->bpn", w_sc_r, x_cl.double()) + bias_c.double().view(1, 1, C))
        close(o, want)
    # pointwise input gradient: channels-last, gated weights (gate indexed by the OUTPUT channel) -> planes
    assert pgemm_supported(B, P, E, C, 0, True, False, False, bf)
    dz, _ = pixel_gemm(do_cl.cuda(), True, w_pw.cuda().t(), out_cl=False, gate=gate.cuda(), gate_on_n=True)
    close(dz, torch.einsum("bne,bpn->bep", w_pw_r, do_cl.double()))
    # shortcut input gradient: channels-last -> channels-last
    assert pgemm_supported(B, P, C, C, 0, True, True, False, bf)
    dxs, _ = pixel_gemm(do_cl.cuda(), True, w_sc.cuda().t(), out_cl=True)
    close(dxs, torch.einsum("nk,bpn->bpk", w_sc_r, do_cl.double()))


def test_expand_bn_hardswish_with_epilogue_statistics_matches_separate_statistics_pass():
    """BatchNorm driven by the GEMM epilogue's partial sums == BatchNorm with its own statistics pass (same y)."""
    from lmnet_b200.bnact import bn_act
    from lmnet_b200.conv1x1 import expand_1x1

    torch.manual_seed(2)
    conv = torch.nn.Conv2d(12, 24, 1).cuda()
    bn1, bn2 = torch.nn.BatchNorm2d(24).cuda().train(), torch.nn.BatchNorm2d(24).cuda().train()
    x = torch.randn(4, 12, 40, 36, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y, part = expand_1x1(conv, x, want_stats=True)
        assert part is not None and part.shape[0] == 24
        a = bn_act(bn1, y, "hardswish", stats=part)
        b = bn_act(bn2, y, "hardswish")
    assert rel_err(a.float(), b.float()) < 1e-6 * 0 + 4e-3       # coefficients differ by fp32 summation order only
    assert torch.allclose(bn1.running_mean, bn2.running_mean, rtol=1e-5, atol=1e-7)
    assert torch.allclose(bn1.running_var, bn2.running_var, rtol=1e-5, atol=1e-7)
    assert int(bn1.num_batches_tracked) == 1
