"""GPU parity tests for the fused depthwise-branch section of ReparamConv (through the C ABI) against
(a) vectors produced by the unmodified reference class (tests/golden/reparam_golden.pt) and
(b) the plain-torch oracle restatement (oracle/reparam_ref.py) run in fp64 on the CPU."""
import copy
import os

import pytest
import torch

from _helpers import GOLDEN_DIR, fill_deterministic, rel_err

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


@pytest.fixture(autouse=True)
def _no_tf32():
    # the 1x1 convolutions around the fused section are cuDNN calls; TF32 would break the 1e-4 bound
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _block(cin, e, cout, seed):
    from lmnet_b200.model import ReparamConv

    blk = ReparamConv(cin, e, cout)
    fill_deterministic(blk, seed=seed)
    return blk


def test_block_against_reference_golden_fp32():
    """fp32 CUDA path vs the reference class's own fp64 outputs / gradients / running statistics."""
    gold = torch.load(os.path.join(GOLDEN_DIR, "reparam_golden.pt"))
    blk = _block(6, 8, 4, seed=7).cuda()
    x = gold["x"].float().cuda().requires_grad_()
    blk.train()
    y = blk(x)
    y.backward(gold["go"].float().cuda())
    assert rel_err(y.cpu(), gold["train_out"]) < 1e-4
    assert rel_err(x.grad.cpu(), gold["dx"]) < 1e-4
    for k, p in blk.named_parameters():
        if float(gold["grads"][k].abs().max()) < 1e-9:     # bias in front of a BatchNorm: exactly zero gradient
            assert float(p.grad.abs().max()) < 1e-4, k
            continue
        assert rel_err(p.grad.cpu(), gold["grads"][k]) < 2e-4, k
    for k, b in blk.named_buffers():
        assert torch.allclose(b.double().cpu(), gold["buffers_after"][k].double(), rtol=1e-5, atol=1e-6), k
    blk.eval()
    with torch.no_grad():
        ye = blk(gold["x"].float().cuda())
        assert rel_err(ye.cpu(), gold["eval_out"]) < 1e-4
        blk.switch_to_deploy()
        yd = blk(gold["x"].float().cuda())
        assert rel_err(yd.cpu(), gold["deploy_out"]) < 1e-4
        assert rel_err(yd, ye) < 1e-5                       # eval == deploy (reference's own identity)


# (B, Cin, E, Cout, H, W)
SHAPES = [
    (2, 3, 24, 12, 40, 36),      # LM-Net level-1 widths
    (2, 12, 24, 12, 33, 67),     # odd W (scalar store path), > 1 column stripe, ragged row tile
    (3, 24, 48, 24, 16, 130),    # three stripes
    (2, 96, 192, 96, 11, 11),    # LM-Net level-4 widths, tiny plane
    (1, 4, 6, 4, 70, 5),         # narrower than the 5-wide window halo, several row tiles
    (2, 5, 7, 3, 9, 8),          # odd channel count
]


def _run_block(blk, x, go, dtype, forward):
    xc = x.cuda().requires_grad_()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
        y = forward(blk, xc)
    y.backward(go.cuda().to(y.dtype))
    return y, xc


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_train_forward_backward_vs_oracle(shape, dtype):
    """Whole block, training mode.  fp32: within 1e-4 of the fp64 oracle.  bf16 autocast: the block also
    contains stock cuDNN/ATen bf16 ops (1x1 convs, BatchNorm, Hardswish) whose own rounding dominates,
    so the bound is "no worse than the stock bf16 path": err(ours vs fp64) <= max(2e-2, 1.5 x err(stock
    torch bf16 on the same GPU vs fp64)).  The fused op alone is held to 2e-2 in test_raw_op_outputs_and_pool."""
    from oracle.reparam_ref import reparam_forward_ref

    B, cin, e, cout, H, W = shape
    blk = _block(cin, e, cout, seed=11)
    ref = copy.deepcopy(blk).double()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, cin, H, W, generator=g)
    go = torch.randn(B, cout, H, W, generator=g)
    ref.train()
    xr = x.double().requires_grad_()
    yr = reparam_forward_ref(ref, xr)
    yr.backward(go.double())

    stock = copy.deepcopy(blk).cuda().train()
    blk = blk.cuda().train()
    y, xc = _run_block(blk, x, go, dtype, lambda m, t: m(t))
    tol = TOL[dtype]

    def bound(name, got, want, mult):
        if dtype == torch.float32:
            return tol * mult
        return max(tol * mult, 1.5 * rel_err(got, want) + 1e-3)

    ys, xs = _run_block(stock, x, go, dtype, reparam_forward_ref)
    assert rel_err(y.float().cpu(), yr) < bound("y", ys.float().cpu(), yr, 1)
    assert rel_err(xc.grad.cpu(), xr.grad) < bound("dx", xs.grad.cpu(), xr.grad, 2)
    stock_params = dict(stock.named_parameters())
    for (n, p), (_, pr) in zip(blk.named_parameters(), ref.named_parameters()):
        scale = float(pr.grad.abs().max())
        if scale < 1e-12:          # e.g. biases in front of a BatchNorm have exactly zero gradient
            # exact cancellation in exact arithmetic; rounding noise grows like sqrt(#pixels)
            assert float(p.grad.abs().max()) < tol * (B * H * W) ** 0.5, n
            continue
        assert rel_err(p.grad.cpu(), pr.grad) < bound(n, stock_params[n].grad.cpu(), pr.grad, 3), n
    for (n, b), (_, br) in zip(blk.named_buffers(), ref.named_buffers()):
        if n.endswith("num_batches_tracked"):
            assert int(b) == int(br) == 1, n
        else:
            assert torch.allclose(b.double().cpu(), br, rtol=5 * tol, atol=tol), n


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_raw_op_outputs_and_pool(dtype):
    """z, pool of the fused op itself (not the whole block) vs the oracle, incl. the pool gradient path."""
    from lmnet_b200.reparam import fused_dw_bn_gelu
    from oracle.reparam_ref import dw_bn_gelu

    blk = _block(4, 16, 4, seed=3)
    ref = copy.deepcopy(blk).double().train()
    g = torch.Generator().manual_seed(9)
    x1 = torch.randn(2, 16, 37, 70, generator=g)
    gz = torch.randn(2, 16, 37, 70, generator=g)
    gp = torch.randn(2, 16, generator=g)
    xr = x1.double().requires_grad_()
    zr, pr = dw_bn_gelu(ref, xr)
    (zr * gz.double()).sum().add((pr * gp.double()).sum()).backward()
    blk = blk.cuda().train()
    xc = x1.to(dtype).cuda().requires_grad_()
    z, p = fused_dw_bn_gelu(blk, xc)
    assert z.dtype == dtype and p.dtype == torch.float32 and p.shape == (2, 16)
    (z.float() * gz.cuda()).sum().add((p * gp.cuda()).sum()).backward()
    tol = TOL[dtype]
    assert rel_err(z.float().cpu(), zr) < tol
    assert rel_err(p.cpu(), pr) < tol
    assert rel_err(xc.grad.float().cpu(), xr.grad) < tol * 2
    assert rel_err(blk.large_conv.conv.weight.grad.cpu(), ref.large_conv.conv.weight.grad) < tol * 3
    assert rel_err(blk.hor_conv.bn.weight.grad.cpu(), ref.hor_conv.bn.weight.grad) < tol * 3


def test_eval_and_deploy_at_full_size_identity():
    """Size-independent property at the BASELINE size (B=16, E=24, 352x352, bf16): eval-mode output ==
    output after switch_to_deploy() (the reference's own known-answer identity), and
    train-mode statistics: every BN branch output has batch mean beta, so mean(u) == sum of betas."""
    blk = _block(12, 24, 12, seed=2).cuda()
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(16, 12, 352, 352, device="cuda", generator=g)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        blk.eval()
        ye = blk(x)
        blk2 = copy.deepcopy(blk)
        blk2.switch_to_deploy()
        yd = blk2(x)
        assert rel_err(yd.float(), ye.float()) < 2e-2
    # train-mode property on the raw op: per-channel mean of u equals sum_br beta_br
    from lmnet_b200 import reparam

    blk.train()
    with torch.no_grad():
        x1 = blk.expand_conv(x).to(torch.bfloat16)
        z, pool = reparam.fused_dw_bn_gelu(blk, x1)
        assert torch.isfinite(z.float()).all()
        assert rel_err(pool, z.float().mean(dim=(2, 3))) < 1e-3          # pool is the spatial mean of z


def test_no_grad_in_eval_mode_is_enforced():
    blk = _block(4, 8, 4, seed=1).cuda().eval()
    x = torch.randn(1, 4, 16, 16, device="cuda", requires_grad=True)
    with pytest.raises(NotImplementedError):
        blk(x)
    with torch.no_grad():
        assert blk(x).shape == (1, 4, 16, 16)


# shapes that take the TMA + mbarrier pipeline kernels (16-bit storage, W % 8 == 0): (B, E, H, W)
TMA_SHAPES = [
    (2, 16, 40, 72),      # two forward stripes (64 + 8), two dx stripes (56 + 16), ragged row tiles
    (1, 8, 70, 136),      # three stripes, three row tiles (32, 32, 6), dx tiles of 28 rows
    (2, 8, 33, 64),       # exactly one forward stripe; dx needs two (56 + 8)
    (1, 6, 9, 8),         # plane smaller than the TMA box in both directions
    (2, 24, 88, 88),      # LM-Net level-3 map
    (3, 4, 64, 128),      # tiles align exactly with the image (no masks anywhere)
    (1, 5, 120, 56),      # several row tiles per band, one dx stripe exactly
    (2, 4, 3, 16),        # every pixel is on the 2-pixel frame of the composite backward
    (1, 8, 2, 8),         # two rows (with one row the 3x1 branch is a single tap under BatchNorm: its weight
                          # gradient is exactly zero and only rounding noise could be compared)
    (1, 8, 352, 352),     # LM-Net level-1 plane: 6 stripes, 11 row tiles
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", TMA_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_tma_pipeline_kernels_vs_oracle(shape, dtype):
    """The fused op on the TMA path vs the fp64 oracle: z, pool, dx and EVERY parameter gradient (the depthwise weight
    gradients come out of the dx kernel's Gram products), training-mode statistics, plus eval mode; and agreement with
    the first-generation (non-TMA) kernels on the same inputs (LMNET_DW_NO_TMA=1)."""
    from lmnet_b200.reparam import fused_dw_bn_gelu
    from oracle.reparam_ref import dw_bn_gelu

    B, E, H, W = shape
    blk = _block(4, E, 4, seed=13)
    ref = copy.deepcopy(blk).double().train()
    g = torch.Generator().manual_seed(17)
    x1 = torch.randn(B, E, H, W, generator=g).to(dtype)
    gz = torch.randn(B, E, H, W, generator=g)
    gp = torch.randn(B, E, generator=g)
    xr = x1.double().requires_grad_()
    zr, pr = dw_bn_gelu(ref, xr)
    (zr * gz.double()).sum().add((pr * gp.double()).sum()).backward()
    tol = 2e-2 if dtype == torch.bfloat16 else 5e-3

    def run(no_tma):
        if no_tma:
            os.environ["LMNET_DW_NO_TMA"] = "1"
        try:
            m = copy.deepcopy(blk).cuda().train()
            xc = x1.cuda().requires_grad_()
            z, p = fused_dw_bn_gelu(m, xc)
            (z.float() * gz.cuda()).sum().add((p * gp.cuda()).sum()).backward()
            torch.cuda.synchronize()
            return m, xc, z, p
        finally:
            os.environ.pop("LMNET_DW_NO_TMA", None)

    m, xc, z, p = run(False)
    assert rel_err(z.float().cpu(), zr) < tol
    assert rel_err(p.cpu(), pr) < tol
    assert rel_err(xc.grad.float().cpu(), xr.grad) < tol * 2
    # the 2-pixel frame of every plane is patched by a separate kernel in the composite backward: check it on its own
    fm = torch.zeros(H, W, dtype=torch.bool)
    fm[:2], fm[-2:], fm[:, :2], fm[:, -2:] = True, True, True, True
    assert rel_err(xc.grad.float().cpu()[:, :, fm], xr.grad[:, :, fm]) < tol * 2
    names = ("large_conv", "square_conv", "ver_conv", "hor_conv")
    for n in names:
        a, b = getattr(m, n), getattr(ref, n)
        assert rel_err(a.conv.weight.grad.cpu(), b.conv.weight.grad) < tol * 3, n
        assert rel_err(a.bn.weight.grad.cpu(), b.bn.weight.grad) < tol * 3, n
        assert rel_err(a.bn.bias.grad.cpu(), b.bn.bias.grad) < tol * 3, n
        assert torch.allclose(a.bn.running_mean.double().cpu(), b.bn.running_mean, rtol=5 * tol, atol=tol), n
        assert torch.allclose(a.bn.running_var.double().cpu(), b.bn.running_var, rtol=5 * tol, atol=tol), n
    # composite backward (Gram by-product of the forward) vs the recomputing TMA backward on the same inputs
    os.environ["LMNET_DW_NO_GRAM"] = "1"
    try:
        m1, xc1, z1, p1 = run(False)
    finally:
        os.environ.pop("LMNET_DW_NO_GRAM", None)
    assert rel_err(z.float(), z1.float()) < 1e-2 and rel_err(p, p1) < 1e-3
    assert rel_err(xc.grad.float(), xc1.grad.float()) < 2e-2
    for n in names:
        assert rel_err(getattr(m, n).conv.weight.grad, getattr(m1, n).conv.weight.grad) < 2e-2, n
        assert rel_err(getattr(m, n).bn.weight.grad, getattr(m1, n).bn.weight.grad) < 1e-4, n
    m0, xc0, z0, p0 = run(True)
    assert rel_err(z.float(), z0.float()) < 1e-2 and rel_err(p, p0) < 1e-3
    assert rel_err(xc.grad.float(), xc0.grad.float()) < 2e-2
    for n in names:
        assert rel_err(getattr(m, n).conv.weight.grad, getattr(m0, n).conv.weight.grad) < 2e-2, n
    # eval mode (running statistics folded into one 5x5 kernel)
    ref.eval()
    m.eval()
    ref.load_state_dict({k: v.double().cpu() for k, v in m.state_dict().items()})
    with torch.no_grad():
        ze, pe = fused_dw_bn_gelu(m, x1.cuda())
        zre, pre = dw_bn_gelu(ref, x1.double())
    assert rel_err(ze.float().cpu(), zre) < tol and rel_err(pe.cpu(), pre) < tol
