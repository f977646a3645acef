"""GPU parity tests for the neighbourhood-attention kernels (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): 1e-4 relative in fp32, 2e-2 in bf16 — relative to the largest
magnitude of the reference tensor (tests/_helpers.rel_err)."""
import itertools

import pytest
import torch

from _helpers import rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2, torch.float16: 5e-3}

# (B, H, W, heads, D, K, d)
FUSED_CASES = [
    (2, 11, 13, 12, 1, 3, 1),    # LM-Net stage 1 head dim (hd=1), ragged sizes
    (2, 12, 10, 12, 2, 3, 1),    # stage 2
    (1, 9, 16, 12, 4, 3, 1),     # stage 3
    (1, 8, 7, 12, 8, 3, 1),      # stage 4
    (1, 3, 3, 12, 1, 3, 1),      # minimum size H == W == K
    (1, 44, 44, 12, 8, 3, 1),    # real stage-4 map of the 352x352 config
    (1, 20, 18, 12, 1, 7, 1),    # BASELINE microbench kernel 7
    (1, 17, 14, 12, 2, 7, 2),    # kernel 7, dilation 2, H,W barely >= K*d and ragged residues
    (1, 16, 15, 6, 4, 5, 1),
    (1, 13, 12, 5, 4, 3, 3),     # heads not divisible by the vector grouping, dilation 3
    (1, 19, 20, 2, 16, 9, 1),    # runtime-K path
    (1, 14, 14, 3, 32, 5, 2),    # NAT-style head dim
    (1, 14, 15, 3, 1, 13, 1),    # largest kernel
    # row-streaming kernels (16-bit storage): several stripes / bands, stripe boundary moved off the last K-1 columns
    (1, 40, 70, 12, 1, 3, 1),    # 2 stripes of 62 + 8, 5 bands
    (1, 27, 64, 12, 1, 3, 1),    # 64 - 62 < K: boundary moves to W - K; bands 8,8,8,3 -> last boundary moves too
    (1, 33, 61, 12, 2, 3, 1),    # stripes of 30: 61 - 60 < K
    (1, 30, 30, 12, 4, 3, 1),    # stripes of 14
    (1, 26, 29, 12, 2, 3, 2),    # dilation 2: four sub-grids of different extents
    (2, 24, 40, 12, 1, 5, 1),    # streaming forward only (K = 5), backward on the general path
    (1, 30, 33, 12, 4, 7, 1),    # streaming forward, K = 7
]


def _mk(shape, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64).to(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("case", FUSED_CASES, ids=lambda c: "x".join(map(str, c)))
@pytest.mark.parametrize("use_rpb", [True, False])
def test_fused_forward_backward_vs_oracle(case, dtype, use_rpb):
    from natten.functional import na2d
    from oracle.na2d_ref import c_oracle

    B, H, W, heads, D, K, d = case
    q, k, v, go = (_mk((B, H, W, heads, D), dtype, s) for s in (1, 2, 3, 4))
    rpb = (0.3 * _mk((heads, 2 * K - 1, 2 * K - 1), torch.float32, 5)) if use_rpb else None
    # oracle on exactly the values the kernel sees (inputs rounded to dtype), in fp64
    o = c_oracle()
    q64, k64, v64, go64 = (t.double() for t in (q, k, v, go))
    r64 = rpb.double() if use_rpb else None
    ref = o.fused_fwd(q64, k64, v64, r64, K, d)
    rdq, rdk, rdv, rdrpb = o.fused_bwd(q64, k64, v64, r64, go64, K, d)

    qc, kc, vc = (t.cuda().requires_grad_() for t in (q, k, v))
    rc = rpb.cuda().requires_grad_() if use_rpb else None
    out = na2d(qc, kc, vc, K, d, rel_pos_bias=rc)
    assert out.dtype == dtype and out.shape == q.shape
    tol = TOL[dtype]
    assert rel_err(out.cpu(), ref) < tol
    out.backward(go.cuda())
    assert rel_err(qc.grad.cpu(), rdq) < tol
    assert rel_err(kc.grad.cpu(), rdk) < tol
    assert rel_err(vc.grad.cpu(), rdv) < tol
    if use_rpb:
        assert rc.grad.dtype == torch.float32
        assert rel_err(rc.grad.cpu(), rdrpb) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("K,d,D", [(3, 1, 1), (3, 1, 8), (3, 2, 2), (7, 1, 4)])
def test_packed_qkv_equals_separate(dtype, K, d, D):
    from natten.functional import na2d, na2d_qkvpacked

    B, H, W, heads = 2, 15, 16, 12
    qkv = _mk((B, H, W, 3, heads, D), dtype, 7).cuda().requires_grad_()
    rpb = (0.3 * _mk((heads, 2 * K - 1, 2 * K - 1), torch.float32, 8)).cuda().requires_grad_()
    go = _mk((B, H, W, heads, D), dtype, 9).cuda()
    out = na2d_qkvpacked(qkv, K, d, rel_pos_bias=rpb)
    out.backward(go)
    g_packed, g_rpb = qkv.grad.clone(), rpb.grad.clone()
    qkv.grad = rpb.grad = None
    q, k, v = (qkv[:, :, :, i].contiguous() for i in range(3))
    out2 = na2d(q, k, v, K, d, rel_pos_bias=rpb)
    assert torch.equal(out, out2)          # same kernel, different strides: bit-identical
    out2.backward(go)
    assert torch.equal(g_packed, qkv.grad)
    assert rel_err(g_rpb, rpb.grad) < 1e-5  # drpb partial sums depend on the CTA partition only


UNFUSED_CASES = [(2, 3, 9, 11, 4, 3, 1), (1, 12, 12, 14, 1, 3, 2), (1, 2, 15, 14, 8, 7, 1), (1, 2, 21, 14, 5, 7, 2),
                 (1, 2, 12, 13, 20, 5, 1), (1, 1, 13, 13, 3, 11, 1)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", UNFUSED_CASES, ids=lambda c: "x".join(map(str, c)))
def test_unfused_ops_vs_oracle(case, dtype):
    from natten.functional import na2d_av, na2d_qk
    from oracle.na2d_ref import c_oracle

    B, heads, H, W, D, K, d = case
    o = c_oracle()
    q, k, v = (_mk((B, heads, H, W, D), dtype, s) for s in (1, 2, 3))
    rpb = 0.3 * _mk((heads, 2 * K - 1, 2 * K - 1), torch.float32, 4)
    gattn = _mk((B, heads, H, W, K * K), dtype, 5)
    gout = _mk((B, heads, H, W, D), dtype, 6)
    tol = TOL[dtype]
    qc, kc, vc, rc = (t.cuda().requires_grad_() for t in (q, k, v, rpb))
    attn = na2d_qk(qc, kc, K, d, rel_pos_bias=rc)
    ref_attn = o.qk_fwd(q.double(), k.double(), rpb.double(), K, d)
    assert rel_err(attn.cpu(), ref_attn) < tol
    attn.backward(gattn.cuda())
    rdq, rdk, rdrpb = o.qk_bwd(q.double(), k.double(), gattn.double(), K, d)
    assert rel_err(qc.grad.cpu(), rdq) < tol and rel_err(kc.grad.cpu(), rdk) < tol
    assert rel_err(rc.grad.cpu(), rdrpb) < tol
    p = _mk((B, heads, H, W, K * K), torch.float32, 7).softmax(-1).to(dtype)
    pc = p.cuda().requires_grad_()
    out = na2d_av(pc, vc, K, d)
    assert rel_err(out.cpu(), o.av_fwd(p.double(), v.double(), K, d)) < tol
    out.backward(gout.cuda())
    rdp, rdv = o.av_bwd(p.double(), v.double(), gout.double(), K, d)
    assert rel_err(pc.grad.cpu(), rdp) < tol and rel_err(vc.grad.cpu(), rdv) < tol


def test_unfused_accepts_transposed_views_like_dinat():
    """transformers' DiNAT feeds `.view(B,H,W,heads,D).permute(...)`-style non-contiguous tensors."""
    from natten.functional import natten2dav, natten2dqkrpb
    from oracle.na2d_ref import c_oracle

    B, H, W, heads, D, K = 1, 9, 10, 4, 8, 3
    base = [_mk((B, H, W, heads, D), torch.float32, s).cuda() for s in (1, 2, 3)]
    q, k, v = (t.permute(0, 3, 1, 2, 4) for t in base)
    assert not q.is_contiguous()
    rpb = (0.2 * _mk((heads, 5, 5), torch.float32, 4)).cuda()
    out = natten2dav(natten2dqkrpb(q, k, rpb, K, 1).softmax(-1), v, K, 1)
    o = c_oracle()
    ref = o.fused_fwd(*(t.cpu().double() for t in base), rpb.cpu().double(), K, 1, scale=1.0)
    assert rel_err(out.permute(0, 2, 3, 1, 4).cpu(), ref) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_module_vs_oracle_module(dtype):
    """natten.NeighborhoodAttention2D (LM-Net's construction) against the oracle module, fwd + bwd."""
    import natten
    from oracle.na2d_ref import OracleNeighborhoodAttention2D

    torch.manual_seed(0)
    for C, H, W in [(12, 17, 19), (96, 11, 9)]:
        m = natten.NeighborhoodAttention2D(dim=C, num_heads=12, kernel_size=3)
        ref = OracleNeighborhoodAttention2D(C, 12, 3).double()
        ref.load_state_dict(m.state_dict())
        x = torch.randn(2, H, W, C)
        xr = x.double().requires_grad_()
        yr = ref(xr)
        g = torch.randn(2, H, W, C)
        yr.backward(g.double())
        m = m.cuda()
        xc = x.cuda().requires_grad_()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
            y = m(xc)
        y.backward(g.cuda().to(y.dtype))
        tol = TOL[dtype]
        assert rel_err(y.float().cpu(), yr) < tol
        assert rel_err(xc.grad.cpu(), xr.grad) < tol * 2
        for (n, p), (_, pr) in zip(m.named_parameters(), ref.named_parameters()):
            assert rel_err(p.grad.cpu(), pr.grad) < tol * 2, n


# ---------------------------------------------------------------------------------------------
# the real BASELINE stage shapes (352x352 config) against the C oracle: B = 1 (B = 2 at the two small stages),
# kernel 3 (what LM-Net runs) and BASELINE's microbench kernel 7 with dilation 1 and 2, fwd + bwd + drpb.
# The C oracle needs < 1 s per case at these sizes.
# ---------------------------------------------------------------------------------------------
STAGE_SHAPES = [(1, 352, 1), (1, 176, 2), (2, 88, 4), (2, 44, 8)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("K,d", [(3, 1), (7, 1), (7, 2)])
@pytest.mark.parametrize("B,R,D", STAGE_SHAPES, ids=lambda v: str(v))
def test_stage_shapes_vs_oracle(B, R, D, K, d, dtype):
    from natten.functional import na2d
    from oracle.na2d_ref import c_oracle

    heads = 12
    q, k, v, go = (_mk((B, R, R, heads, D), dtype, s) for s in (11, 12, 13, 14))
    rpb = 0.3 * _mk((heads, 2 * K - 1, 2 * K - 1), torch.float32, 15)
    o = c_oracle()
    q64, k64, v64, go64 = (t.double() for t in (q, k, v, go))
    ref = o.fused_fwd(q64, k64, v64, rpb.double(), K, d)
    rdq, rdk, rdv, rdrpb = o.fused_bwd(q64, k64, v64, rpb.double(), go64, K, d)
    qc, kc, vc = (t.cuda().requires_grad_() for t in (q, k, v))
    rc = rpb.cuda().requires_grad_()
    out = na2d(qc, kc, vc, K, d, rel_pos_bias=rc)
    out.backward(go.cuda())
    tol = TOL[dtype]
    assert rel_err(out.cpu(), ref) < tol
    assert rel_err(qc.grad.cpu(), rdq) < tol
    assert rel_err(kc.grad.cpu(), rdk) < tol
    assert rel_err(vc.grad.cpu(), rdv) < tol
    # drpb sums B*R*R*9..49 products per bin: rounding noise of 16-bit dS grows like sqrt(N) against a sum that can
    # cancel; the bound is relative to the largest bin as everywhere else
    assert rel_err(rc.grad.cpu(), rdrpb) < tol


# ---------------------------------------------------------------------------------------------
# size-independent properties at the full BASELINE batch (B = 16)
# ---------------------------------------------------------------------------------------------
STAGES = [(16, 352, 1), (16, 176, 2), (16, 88, 4), (16, 44, 8)]


@pytest.mark.parametrize("B,R,D", STAGES)
@pytest.mark.parametrize("K,d", [(3, 1), (7, 2)])
def test_full_size_properties(B, R, D, K, d):
    from natten.functional import na2d

    heads = 12
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v, v2 = (torch.randn(B, R, R, heads, D, device="cuda", dtype=torch.bfloat16, generator=g) for _ in range(4))
    rpb = 0.1 * torch.randn(heads, 2 * K - 1, 2 * K - 1, device="cuda", generator=g)
    # (1) softmax weights sum to one: v == 1  =>  out == 1
    ones = torch.ones_like(v)
    out1 = na2d(q, k, ones, K, d, rel_pos_bias=rpb)
    assert float((out1.float() - 1).abs().max()) < 1e-2
    # (2) linear in v
    a = na2d(q, k, v, K, d, rel_pos_bias=rpb).float()
    b = na2d(q, k, v2, K, d, rel_pos_bias=rpb).float()
    ab = na2d(q, k, (v.float() + v2.float()).to(torch.bfloat16), K, d, rel_pos_bias=rpb).float()
    assert rel_err(ab, a + b) < 2e-2
    # (3) dilation == independent sub-grids
    if d > 1:
        sub = na2d(q[:, 1::d, 0::d].contiguous(), k[:, 1::d, 0::d].contiguous(), v[:, 1::d, 0::d].contiguous(), K, 1,
                   rel_pos_bias=rpb)
        assert torch.equal(sub, na2d(q, k, v, K, d, rel_pos_bias=rpb)[:, 1::d, 0::d])
    # (4) adjoint identity of the backward: <dout, d/dv out . w> == <dv, w>   (out is linear in v)
    vq = v.clone().requires_grad_()
    out = na2d(q, k, vq, K, d, rel_pos_bias=rpb)
    dout = torch.randn_like(out)
    out.backward(dout)
    terms = dout.float() * b                        # b = NA(q,k,v2): the linear map applied to v2
    lhs = float(terms.double().sum())
    rhs = float((vq.grad.float() * v2.float()).double().sum())
    # both sides carry bf16 rounding (2^-9 per term, random sign): the noise scales with the 2-norm
    assert abs(lhs - rhs) <= 4 * 2.0 ** -8 * float(terms.double().norm())


def test_k_equal_to_map_size_is_global_attention():
    """Known-answer property on the GPU path: K == H == W  =>  plain softmax attention (+ Swin bias)."""
    from natten.functional import na2d

    L_, heads, D = 7, 4, 8
    K = L_
    g = torch.Generator().manual_seed(3)
    q, k, v = (torch.randn(2, L_, L_, heads, D, generator=g).cuda() for _ in range(3))
    rpb = (0.5 * torch.randn(heads, 2 * K - 1, 2 * K - 1, generator=g)).cuda()
    ii, jj = torch.meshgrid(torch.arange(L_), torch.arange(L_), indexing="ij")
    ii, jj = ii.reshape(-1).cuda(), jj.reshape(-1).cuda()
    bias = rpb[:, (ii[None, :] - ii[:, None]) + K - 1, (jj[None, :] - jj[:, None]) + K - 1]
    qf, kf, vf = (t.reshape(2, L_ * L_, heads, D).transpose(1, 2) for t in (q, k, v))
    ref = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf, attn_mask=bias[None]).transpose(1, 2)
    out = na2d(q, k, v, K, rel_pos_bias=rpb)
    assert rel_err(out.reshape(2, L_ * L_, heads, D), ref) < 1e-4


def test_general_kernels_still_cover_16bit_storage():
    """The row-streaming kernels take the eligible bf16/fp16 shapes; LMNET_NA_V1=1 (read once per process) routes
    everything to the general kernels, whose vectorised 16-bit paths must stay parity-green.  Also checks that the
    two generations agree with each other on a stage-1 shaped problem."""
    import os
    import subprocess
    import sys

    code = r"""
import sys, torch
sys.path[:0] = [%r, %r, %r]
from natten.functional import na2d
from oracle.na2d_ref import c_oracle
from _helpers import rel_err
o = c_oracle()
for (B, H, W, heads, D, K, d) in [(2, 11, 13, 12, 1, 3, 1), (1, 12, 10, 12, 2, 3, 1), (1, 9, 16, 12, 4, 3, 1), (1, 26, 29, 12, 2, 3, 2)]:
    g = torch.Generator().manual_seed(3)
    q, k, v, go = (torch.randn(B, H, W, heads, D, generator=g, dtype=torch.float64).to(torch.bfloat16) for _ in range(4))
    rpb = 0.3 * torch.randn(heads, 2 * K - 1, 2 * K - 1, generator=g)
    ref = o.fused_fwd(q.double(), k.double(), v.double(), rpb.double(), K, d)
    rdq, rdk, rdv, rdrpb = o.fused_bwd(q.double(), k.double(), v.double(), rpb.double(), go.double(), K, d)
    qc, kc, vc = (t.cuda().requires_grad_() for t in (q, k, v))
    rc = rpb.cuda().requires_grad_()
    out = na2d(qc, kc, vc, K, d, rel_pos_bias=rc)
    out.backward(go.cuda())
    for got, want in ((out, ref), (qc.grad, rdq), (kc.grad, rdk), (vc.grad, rdv), (rc.grad, rdrpb)):
        assert rel_err(got.float().cpu(), want) < 2e-2
print("ok")
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = code % (root, os.path.join(root, "lm-net_b200"), os.path.join(root, "tests"))
    env = dict(os.environ, LMNET_NA_V1="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
