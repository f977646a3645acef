"""Pins the CPU oracle (oracle/) — the real natten is absent offline, so the oracle is pinned by
(a) agreement of two independently written index formulations (C sub-sequence/clamp vs the
closed-form rules quoted in SURVEY.md §8 c3), (b) analytic backward vs autograd, and
(c) known-answer properties published with the algorithm (NAT / DiNAT papers)."""
import itertools

import pytest
import torch
import torch.nn.functional as F

from oracle import na2d_ref as R

torch.manual_seed(0)

CASES = [
    # (B, heads, H, W, D, K, d)
    (2, 3, 7, 9, 4, 3, 1),
    (1, 2, 3, 3, 1, 3, 1),      # minimum size H == W == K
    (1, 2, 6, 11, 2, 3, 2),     # H == K*d
    (2, 1, 13, 10, 3, 5, 2),
    (1, 2, 14, 15, 8, 7, 1),
    (1, 1, 14, 17, 2, 7, 2),    # H == K*d, ragged residues on W
    (1, 2, 12, 10, 1, 3, 3),    # W: residue classes of unequal length
]


def _rand(*shape):
    return torch.randn(*shape, dtype=torch.float64)


@pytest.mark.parametrize("B,Hd,H,W,D,K,d", CASES)
@pytest.mark.parametrize("use_rpb", [True, False])
def test_c_oracle_matches_closed_form_gather(B, Hd, H, W, D, K, d, use_rpb):
    o = R.c_oracle()
    q, k, v = _rand(B, Hd, H, W, D), _rand(B, Hd, H, W, D), _rand(B, Hd, H, W, D)
    rpb = _rand(Hd, 2 * K - 1, 2 * K - 1) if use_rpb else None
    attn_c = o.qk_fwd(q, k, rpb, K, d)
    attn_g = R.na2d_qk_gather(q, k, rpb, K, d)
    torch.testing.assert_close(attn_c, attn_g, rtol=1e-12, atol=1e-12)
    p = attn_g.softmax(-1)
    torch.testing.assert_close(o.av_fwd(p.contiguous(), v, K, d), R.na2d_av_gather(p, v, K, d),
                               rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("B,Hd,H,W,D,K,d", CASES)
def test_c_oracle_backward_matches_autograd(B, Hd, H, W, D, K, d):
    o = R.c_oracle()
    q, k, v = (_rand(B, Hd, H, W, D).requires_grad_() for _ in range(3))
    rpb = _rand(Hd, 2 * K - 1, 2 * K - 1).requires_grad_()
    attn = R.na2d_qk_gather(q, k, rpb, K, d)
    g = _rand(*attn.shape)
    dq, dk, drpb = torch.autograd.grad(attn, (q, k, rpb), g)
    cdq, cdk, cdrpb = o.qk_bwd(q.detach(), k.detach(), g, K, d)
    for a, b in ((dq, cdq), (dk, cdk), (drpb, cdrpb)):
        torch.testing.assert_close(a, b, rtol=1e-11, atol=1e-11)
    p = attn.detach().softmax(-1).requires_grad_()
    out = R.na2d_av_gather(p, v, K, d)
    go = _rand(*out.shape)
    dp, dv = torch.autograd.grad(out, (p, v), go)
    cdp, cdv = o.av_bwd(p.detach().contiguous(), v.detach(), go, K, d)
    torch.testing.assert_close(dp, cdp, rtol=1e-11, atol=1e-11)
    torch.testing.assert_close(dv, cdv, rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("B,Hd,H,W,D,K,d", CASES)
@pytest.mark.parametrize("use_rpb", [True, False])
def test_fused_oracle_fwd_bwd(B, Hd, H, W, D, K, d, use_rpb):
    """Fused C oracle == composition of the unfused ops == autograd of the gather form."""
    o = R.c_oracle()
    q, k, v = (_rand(B, H, W, Hd, D).requires_grad_() for _ in range(3))
    rpb = _rand(Hd, 2 * K - 1, 2 * K - 1).requires_grad_() if use_rpb else None
    out = R.na2d_gather(q, k, v, K, d, rpb)
    go = _rand(*out.shape)
    grads = torch.autograd.grad(out, (q, k, v) + ((rpb,) if use_rpb else ()), go)
    rp = rpb.detach() if use_rpb else None
    cout, lse = o.fused_fwd(q.detach(), k.detach(), v.detach(), rp, K, d, want_lse=True)
    torch.testing.assert_close(out.detach().contiguous(), cout, rtol=1e-11, atol=1e-11)
    cg = o.fused_bwd(q.detach(), k.detach(), v.detach(), rp, go, K, d)
    for a, b in zip(grads, cg):
        torch.testing.assert_close(a, b, rtol=1e-10, atol=1e-10)
    # lse really is the log-sum-exp of the scaled scores
    attn = R.na2d_qk_gather(q.detach().permute(0, 3, 1, 2, 4) * D ** -0.5, k.detach().permute(0, 3, 1, 2, 4), rp, K, d)
    torch.testing.assert_close(lse, attn.logsumexp(-1).permute(0, 2, 3, 1).contiguous(), rtol=1e-11, atol=1e-11)


def test_fp32_oracle_close_to_fp64():
    o = R.c_oracle()
    q, k, v = (_rand(2, 9, 8, 12, 4) for _ in range(3))
    rpb = _rand(12, 5, 5) * 0.02
    ref = o.fused_fwd(q, k, v, rpb, 3, 1)
    out = o.fused_fwd(q.float(), k.float(), v.float(), rpb.float(), 3, 1)
    torch.testing.assert_close(out.double(), ref, rtol=1e-5, atol=1e-5)


# ----------------------------- known-answer properties ----------------------------- #
@pytest.mark.parametrize("L", [3, 5, 7])
def test_kat_full_window_equals_global_attention_with_swin_bias(L):
    """NAT paper: with the neighbourhood as large as the feature map NA *is* self-attention.
    With rpb it is self-attention + Swin's relative position bias B[h, (ki-i)+K-1, (kj-j)+K-1]."""
    K = L
    Hd, D = 2, 4
    q, k, v = (_rand(1, L, L, Hd, D) for _ in range(3))
    rpb = _rand(Hd, 2 * K - 1, 2 * K - 1)
    ii, jj = torch.meshgrid(torch.arange(L), torch.arange(L), indexing="ij")
    ii, jj = ii.reshape(-1), jj.reshape(-1)
    bias = rpb[:, (ii[None, :] - ii[:, None]) + K - 1, (jj[None, :] - jj[:, None]) + K - 1]  # [Hd, Nq, Nk]
    qf, kf, vf = (t.reshape(1, L * L, Hd, D).transpose(1, 2) for t in (q, k, v))
    ref = F.scaled_dot_product_attention(qf, kf, vf, attn_mask=bias[None]).transpose(1, 2).reshape(1, L, L, Hd, D)
    out = R.c_oracle().fused_fwd(q, k, v, rpb, K, 1)
    torch.testing.assert_close(out, ref.contiguous(), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("K", [3, 5])
def test_kat_interior_equals_unfold_sliding_window(K):
    """Away from the borders NA is plain sliding-window attention (F.unfold gives the windows)."""
    B, Hd, H, W, D = 1, 2, 9, 10, 3
    ns = K // 2
    q, k, v = (_rand(B, Hd, H, W, D) for _ in range(3))
    attn = R.c_oracle().qk_fwd(q, k, None, K, 1)
    # unfold k: [B*Hd, D, H, W] -> [B*Hd, D*K*K, Hi*Wi]
    ku = F.unfold(k.reshape(B * Hd, H, W, D).permute(0, 3, 1, 2), K).reshape(B, Hd, D, K * K, H - 2 * ns, W - 2 * ns)
    ref = torch.einsum("bhijd,bhdnij->bhijn", q[:, :, ns:H - ns, ns:W - ns], ku)
    torch.testing.assert_close(attn[:, :, ns:H - ns, ns:W - ns], ref, rtol=1e-12, atol=1e-12)
    # central rpb entry sits at the query itself
    keys, pbs = R.neighbour_table(H, K, 1)
    assert all(pbs[i, ns] == K - 1 for i in range(ns, H - ns))


def test_kat_corner_window_is_shifted_not_padded():
    """NAT paper fig. 2: a corner query attends the K x K block in that corner (no zero padding)."""
    for K, d, L in [(3, 1, 8), (7, 1, 9), (3, 2, 9), (5, 3, 17)]:
        keys, pbs = R.neighbour_table(L, K, d)
        assert keys[0].tolist() == [m * d for m in range(K)]
        last_r = (L - 1) % d
        sub = list(range(last_r, L, d))
        assert keys[L - 1].tolist() == sub[-K:]
        # every window has exactly K distinct in-range keys in the query's residue class
        for i in range(L):
            assert len(set(keys[i])) == K and keys[i].min() >= 0 and keys[i].max() < L
            assert all(kk % d == i % d for kk in keys[i])
            assert (keys[i] - i).tolist() == ((pbs[i] - (K - 1)) * d).tolist()


@pytest.mark.parametrize("K,d", [(3, 2), (3, 3), (5, 2)])
def test_kat_dilation_is_independent_subgrids(K, d):
    """DiNAT: dilation d == running undilated NA on each of the d*d interleaved sub-grids."""
    B, H, W, Hd, D = 1, K * d + 3, K * d + 2, 2, 2
    q, k, v = (_rand(B, H, W, Hd, D) for _ in range(3))
    rpb = _rand(Hd, 2 * K - 1, 2 * K - 1)
    o = R.c_oracle()
    out = o.fused_fwd(q, k, v, rpb, K, d)
    for ri, rj in itertools.product(range(d), range(d)):
        sub = [t[:, ri::d, rj::d].contiguous() for t in (q, k, v)]
        torch.testing.assert_close(out[:, ri::d, rj::d].contiguous(), o.fused_fwd(*sub, rpb, K, 1),
                                   rtol=1e-12, atol=1e-12)


def test_oracle_module_state_dict_keys():
    m = R.OracleNeighborhoodAttention2D(dim=24, num_heads=12, kernel_size=3)
    assert list(m.state_dict().keys()) == ["rpb", "qkv.weight", "qkv.bias", "proj.weight", "proj.bias"]
    x = torch.randn(1, 5, 6, 24)
    y = m(x)
    assert y.shape == x.shape
    y.sum().backward()
    assert m.rpb.grad is not None and m.rpb.grad.abs().sum() > 0


def test_cpu_reference_model_is_self_contained():
    """bench.py's cpu_baseline / `--impl reference` arm: the oracle's LM-Net must run forward + backward on CPU
    tensors WITHOUT the test backend that stands in for the CUDA library — i.e. it may not reach any lmnet_b200
    operator (those raise on CPU tensors).  Guards the widened modules (LayerNorm, BN+act, up-sampling)."""
    import torch

    from lmnet_b200 import _lib
    from lmnet_b200.train import build_training, synthetic_batches, train_step
    from oracle.lmnet_ref import build_cpu_reference

    with pytest.raises(RuntimeError):          # the product really is CUDA-only in this process
        _lib.require_cuda(torch.zeros(1))
    net = build_cpu_reference(3, 2, seed=42).train()
    opt, crit, dice = build_training(net, "cpu", fused=False)
    images, labels = synthetic_batches(1, 1, 32, seed=0, pin=False)[0]
    loss, out = train_step(net, opt, images, labels, crit, dice, amp_dtype=None)
    assert torch.isfinite(loss) and out.shape == (1, 2, 32, 32)
    assert all(p.grad is None or torch.isfinite(p.grad).all() for p in net.parameters())


@pytest.mark.parametrize("H,W,K,heads,D", [(9, 11, 3, 2, 4), (7, 7, 7, 1, 2), (12, 8, 5, 3, 1), (6, 13, 3, 12, 1)])
def test_oracle_against_flex_attention_with_the_published_natten_mask(H, W, K, heads, D):
    """Independent pin of the forward: PyTorch's own FlexAttention (unfused CPU path; it has no CPU backward) with the
    neighbourhood mask as published for NATTEN in PyTorch's attention-gym examples — window centre =
    clamp(query, K//2, L-1-K//2) per axis, keys within K//2 of the centre — plus the relative positional bias as a
    score_mod indexed by (key - query + K - 1).  The oracle's backward is pinned to its forward by autograd elsewhere
    in this file."""
    import warnings

    from torch.nn.attention.flex_attention import create_block_mask, flex_attention

    def mask_mod(b, h, q_idx, kv_idx):
        qy, qx, ky, kx = q_idx // W, q_idx % W, kv_idx // W, kv_idx % W
        cx = qx.clamp(K // 2, (W - 1) - K // 2)
        cy = qy.clamp(K // 2, (H - 1) - K // 2)
        return ((cx - kx).abs() <= K // 2) & ((cy - ky).abs() <= K // 2)

    g = torch.Generator().manual_seed(H * 100 + W)
    q, k, v = (torch.randn(1, heads, H * W, D, generator=g, dtype=torch.float64) for _ in range(3))
    rpb = 0.5 * torch.randn(heads, 2 * K - 1, 2 * K - 1, generator=g, dtype=torch.float64)

    def score_mod(score, b, h, q_idx, kv_idx):
        dy = (kv_idx // W - q_idx // W + K - 1).clamp(0, 2 * K - 2)     # masked-out pairs may index anything valid
        dx = (kv_idx % W - q_idx % W + K - 1).clamp(0, 2 * K - 2)
        return score + rpb[h, dy, dx]

    def nhwc(t):   # [1, heads, H*W, D] -> [1, H, W, heads, D]
        return t.view(1, heads, H, W, D).permute(0, 2, 3, 1, 4).contiguous()

    o = R.c_oracle()
    with warnings.catch_warnings(), torch.no_grad():
        warnings.simplefilter("ignore")
        bm = create_block_mask(mask_mod, 1, heads, H * W, H * W, device="cpu")
        plain = flex_attention(q, k, v, block_mask=bm)
        biased = flex_attention(q, k, v, score_mod=score_mod, block_mask=bm)
    assert (nhwc(plain) - o.fused_fwd(nhwc(q), nhwc(k), nhwc(v), None, K, 1)).abs().max() < 1e-10
    assert (nhwc(biased) - o.fused_fwd(nhwc(q), nhwc(k), nhwc(v), rpb, K, 1)).abs().max() < 1e-10


def test_two_formulations_agree_on_random_shapes():
    """Hypothesis sweep over (H, W, K, dilation, heads, head dim): the C oracle (sub-sequence / clamp formulation) and
    the closed-form gather tables must give the same logits and the same aggregation on every legal shape — the
    parametrised cases above only sample the edge conditions by hand."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    o = R.c_oracle()

    @st.composite
    def shapes(draw):
        K = draw(st.sampled_from([3, 5, 7, 9]))
        d = draw(st.integers(1, 3))
        H = draw(st.integers(K * d, K * d + 9))
        W = draw(st.integers(K * d, K * d + 9))
        return H, W, K, d, draw(st.integers(1, 3)), draw(st.sampled_from([1, 2, 4]))

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(shapes())
    def check(s):
        H, W, K, d, heads, D = s
        g = torch.Generator().manual_seed(H * 1000 + W * 10 + K + d)
        q, k, v = (torch.randn(1, heads, H, W, D, generator=g, dtype=torch.float64) for _ in range(3))
        rpb = torch.randn(heads, 2 * K - 1, 2 * K - 1, generator=g, dtype=torch.float64)
        attn_c, attn_g = o.qk_fwd(q, k, rpb, K, d), R.na2d_qk_gather(q, k, rpb, K, d)
        torch.testing.assert_close(attn_c, attn_g, rtol=1e-12, atol=1e-12)
        p = attn_g.softmax(-1).contiguous()
        torch.testing.assert_close(o.av_fwd(p, v, K, d), R.na2d_av_gather(p, v, K, d), rtol=1e-12, atol=1e-12)

    check()


def _dense_neighbourhood_mask_and_bias_index(H, W, K, d):
    """Independent dense statement of (dilated) neighbourhood attention, written without any of the oracle's index
    helpers: query (qy, qx) attends key (ky, kx) iff, per axis, both lie in the same residue class mod d and — in the
    coordinates of that sub-sequence (index // d, length ceil((L - r) / d)) — the key is within K//2 of the window
    centre clamp(query, K//2, Lsub-1-K//2) (the attention-gym NATTEN mask applied to each of the d sub-sequences, which
    is DiNAT's definition of dilation).  Returns mask [HW, HW] (bool) and the two bias index maps [HW, HW]."""
    def axis(L):
        i = torch.arange(L)
        r, s = i % d, i // d
        lsub = (L - r + d - 1) // d                                       # length of the query's sub-sequence
        centre = torch.minimum(torch.maximum(s, torch.full_like(s, K // 2)), lsub - 1 - K // 2)
        same = r[:, None] == r[None, :]
        near = (centre[:, None] - s[None, :]).abs() <= K // 2
        rel = (s[None, :] - s[:, None]) + K - 1                           # (key - query) / d + K - 1
        return same & near, rel.clamp(0, 2 * K - 2)
    my, ry = axis(H)
    mx, rx = axis(W)
    mask = (my[:, None, :, None] & mx[None, :, None, :]).reshape(H * W, H * W)
    by = ry[:, None, :, None].expand(H, W, H, W).reshape(H * W, H * W)
    bx = rx[None, :, None, :].expand(H, W, H, W).reshape(H * W, H * W)
    return mask, by, bx


@pytest.mark.parametrize("H,W,K,d,heads,D", [(9, 11, 3, 1, 2, 4), (13, 12, 3, 2, 3, 2), (11, 14, 3, 3, 2, 1),
                                             (15, 17, 7, 2, 2, 2), (21, 22, 7, 3, 1, 4), (12, 10, 5, 2, 12, 1)])
def test_oracle_forward_and_backward_against_dense_masked_sdpa_autograd(H, W, K, d, heads, D):
    """Pins dilation (d = 2, 3) and EVERY backward output (dq, dk, dv, drpb) of the C oracle against torch's own
    autograd through a dense masked softmax attention (fp64): F.scaled_dot_product_attention (math backend on CPU) for
    the forward, and the same attention written with explicit matmul / softmax for the gradients."""
    g = torch.Generator().manual_seed(H * 131 + W * 7 + K + d)
    q, k, v, go = (torch.randn(1, heads, H * W, D, generator=g, dtype=torch.float64) for _ in range(4))
    rpb = 0.5 * torch.randn(heads, 2 * K - 1, 2 * K - 1, generator=g, dtype=torch.float64)
    mask, by, bx = _dense_neighbourhood_mask_and_bias_index(H, W, K, d)
    assert int(mask.sum(1).min()) == K * K and int(mask.sum(1).max()) == K * K     # every query sees exactly K*K keys

    def nhwc(t):
        return t.view(1, heads, H, W, D).permute(0, 2, 3, 1, 4).contiguous()

    qr, kr, vr, rr = (t.clone().requires_grad_() for t in (q, k, v, rpb))
    bias = rr[:, by, bx].masked_fill(~mask, float("-inf"))                    # [heads, HW, HW]
    scale = D ** -0.5
    attn = ((qr * scale) @ kr.transpose(-2, -1) + bias[None]).softmax(-1)
    out = attn @ vr
    out.backward(go)
    with torch.no_grad():
        sdpa = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=bias.detach()[None], scale=scale)
    assert (sdpa - out.detach()).abs().max() < 1e-12

    o = R.c_oracle()
    ref = o.fused_fwd(nhwc(q), nhwc(k), nhwc(v), rpb, K, d)
    assert (nhwc(out.detach()) - ref).abs().max() < 1e-11
    dq, dk, dv, drpb = o.fused_bwd(nhwc(q), nhwc(k), nhwc(v), rpb, nhwc(go), K, d)
    assert (nhwc(qr.grad) - dq).abs().max() < 1e-10
    assert (nhwc(kr.grad) - dk).abs().max() < 1e-10
    assert (nhwc(vr.grad) - dv).abs().max() < 1e-10
    assert (rr.grad - drpb).abs().max() < 1e-9


@pytest.mark.parametrize("H,W,K,d", [(13, 12, 3, 2), (15, 16, 5, 3), (21, 15, 7, 2)])
def test_oracle_against_flex_attention_with_dilation(H, W, K, d):
    """FlexAttention (PyTorch's own implementation) with the dense dilated mask above as mask_mod: third independent
    implementation agreeing on dilation > 1 (forward; FlexAttention has no CPU backward)."""
    import warnings

    from torch.nn.attention.flex_attention import create_block_mask, flex_attention

    heads, D = 2, 4
    mask, by, bx = _dense_neighbourhood_mask_and_bias_index(H, W, K, d)
    g = torch.Generator().manual_seed(H + W + K + d)
    q, k, v = (torch.randn(1, heads, H * W, D, generator=g, dtype=torch.float64) for _ in range(3))
    rpb = 0.5 * torch.randn(heads, 2 * K - 1, 2 * K - 1, generator=g, dtype=torch.float64)

    def nhwc(t):
        return t.view(1, heads, H, W, D).permute(0, 2, 3, 1, 4).contiguous()

    with warnings.catch_warnings(), torch.no_grad():
        warnings.simplefilter("ignore")
        bm = create_block_mask(lambda b, h, qi, ki: mask[qi, ki], 1, heads, H * W, H * W, device="cpu")
        got = flex_attention(q, k, v, score_mod=lambda s, b, h, qi, ki: s + rpb[h, by[qi, ki], bx[qi, ki]], block_mask=bm)
    ref = R.c_oracle().fused_fwd(nhwc(q), nhwc(k), nhwc(v), rpb, K, d)
    assert (nhwc(got) - ref).abs().max() < 1e-10


@pytest.mark.parametrize("K", [3, 5, 7])
def test_oracle_against_the_stock_torch_gather_formulation_of_the_baseline_tool(K):
    """tools/stock_baselines.py times a plain-torch neighbourhood attention (advanced-indexing gather of the clamped
    window + softmax + einsum) as the same-device stock baseline; it is a third, independently written formulation
    and must agree with the C oracle (forward and, through autograd, all four gradients)."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from stock_baselines import stock_na

    from oracle.na2d_ref import c_oracle

    torch.manual_seed(K)
    q, k, v, go = (torch.randn(2, 2 * K + 1, 2 * K + 3, 3, 2, dtype=torch.float64) for _ in range(4))
    rpb = torch.randn(3, 2 * K - 1, 2 * K - 1, dtype=torch.float64)
    o = c_oracle()
    ref = o.fused_fwd(q, k, v, rpb, K, 1, scale=0.7)
    rg = o.fused_bwd(q, k, v, rpb, go, K, 1, scale=0.7)
    qs, ks, vs, rs = (t.clone().requires_grad_() for t in (q, k, v, rpb))
    got = stock_na(qs, ks, vs, rs, K, 0.7)
    assert float((got - ref).abs().max()) < 1e-12
    for g_, w_ in zip(torch.autograd.grad(got, (qs, ks, vs, rs), go), rg):
        assert float((g_ - w_).abs().max()) < 1e-11
