"""The three 1x1 convolutions of ReparamConv on NCHW planes (widening step f1 of SURVEY.md §8;
reference: expand_conv[0] /root/reference/core/modules.py:537, pointwise_conv :576-579, shortcut :581-584,
used at :587 and :598-599).

Forward and input gradients are GEMMs whose large operand is already in the right layout (NCHW planes
= the [K, N] operand), so they go to cuBLAS through torch.bmm — unlike cuDNN's bf16 path no NCHW<->NHWC
transposes are needed.  The weight and bias gradients (a small matrix reduced over up to 124k pixels) run on
the hand-written split-K tensor-core kernel of csrc/wgrad_1x1.cu.  The squeeze-excite gate is folded into
the pointwise weights per sample (W_b = W * gate_b), which removes the gate multiply and its backward pass
over [B,E,H,W]; pointwise + shortcut accumulate into one output and share one read of grad_output.
"""
from __future__ import annotations

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L


def _wgrad(A, B1, B2):
    """A [B,M,P], B1 [B,N1,P], B2 [B,N2,P] or None -> (dW [B,M,N1+N2] fp32, drow [B,M] fp32)."""
    Bn, M, P = A.shape
    N1 = B1.shape[1]
    N2 = 0 if B2 is None else B2.shape[1]
    dims = L.WgradDims(Bn, M, N1, N2, P)
    dW = torch.empty(Bn, M, N1 + N2, dtype=torch.float32, device=A.device)
    drow = torch.empty(Bn, M, dtype=torch.float32, device=A.device)
    n = L.lib().lmnet_wgrad_1x1_workspace_bytes(L.byref(dims))
    ws = torch.empty(max(int(n), 16), dtype=torch.uint8, device=A.device)
    rc = L.lib().lmnet_wgrad_1x1(L.ptr(A), L.ptr(B1), L.ptr(B2), L.ptr(dW), L.ptr(drow), L.ptr(ws), ws.numel(),
                                 L.byref(dims), L.dtype_code(A), L.stream_ptr())
    L.check(rc, "wgrad_1x1")
    return dW, drow


def wgrad_supported(B, M, N1, N2, P, dtype) -> bool:
    if dtype not in (torch.bfloat16, torch.float16):
        return False
    dims = L.WgradDims(B, M, N1, N2, P)
    return bool(L.lib().lmnet_wgrad_1x1_supported(L.byref(dims), L._DTYPES[dtype]))


class _Expand1x1(torch.autograd.Function):
    """y[b] = W @ x[b] + bias on NCHW planes."""

    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, x, w, bias):
        B, K, H, Wd = x.shape
        M = w.shape[0]
        xf = x.reshape(B, K, H * Wd)
        wc = w.to(x.dtype).unsqueeze(0).expand(B, M, K)
        y = torch.baddbmm(bias.to(x.dtype).view(1, M, 1), wc, xf) if bias is not None else torch.bmm(wc, xf)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.bias_dtype = None if bias is None else bias.dtype
        return y.view(B, M, H, Wd)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        B, K, H, Wd = x.shape
        M = w.shape[0]
        dyf = dy.to(x.dtype).contiguous().view(B, M, H * Wd)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.bmm(w.to(x.dtype).t().unsqueeze(0).expand(B, K, M), dyf).view(B, K, H, Wd)
        dW, drow = _wgrad(dyf, x.view(B, K, H * Wd), None)
        dw = dW.sum(0).to(w.dtype)
        db = drow.sum(0).to(ctx.bias_dtype) if ctx.has_bias else None
        return dx, dw, db


class _PointwiseShortcut(torch.autograd.Function):
    """out[b] = (Wpw * gate[b]) @ z[b] + Wsc @ x[b] + bias  — pointwise(gate*z) + shortcut(x) of ReparamConv."""

    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, z, gate, x, wpw, wsc, bias):
        B, E, H, Wd = z.shape
        Cin, Cout = x.shape[1], wpw.shape[0]
        P = H * Wd
        dt = z.dtype
        wg = (wpw.float().unsqueeze(0) * gate.float().unsqueeze(1)).to(dt)            # [B, Cout, E]
        out = torch.baddbmm(bias.to(dt).view(1, Cout, 1), wsc.to(dt).unsqueeze(0).expand(B, Cout, Cin), x.reshape(B, Cin, P))
        out = torch.baddbmm(out, wg, z.reshape(B, E, P))
        ctx.save_for_backward(z, gate, x, wpw, wsc, wg)
        ctx.bias_dtype = bias.dtype
        return out.view(B, Cout, H, Wd)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        z, gate, x, wpw, wsc, wg = ctx.saved_tensors
        B, E, H, Wd = z.shape
        Cin, Cout = x.shape[1], wpw.shape[0]
        P = H * Wd
        dt = z.dtype
        do = dout.to(dt).contiguous().view(B, Cout, P)
        dz = torch.bmm(wg.transpose(1, 2), do).view(B, E, H, Wd)
        dx = None
        if ctx.needs_input_grad[2]:
            dx = torch.bmm(wsc.to(dt).t().unsqueeze(0).expand(B, Cin, Cout), do).view(B, Cin, H, Wd)
        dW, drow = _wgrad(do, z.view(B, E, P), x.reshape(B, Cin, P))                  # [B, Cout, E + Cin]
        dWg = dW[:, :, :E]
        dwpw = (dWg * gate.float().unsqueeze(1)).sum(0).to(wpw.dtype)
        dgate = (dWg * wpw.float().unsqueeze(0)).sum(1).to(gate.dtype)                # [B, E]
        dwsc = dW[:, :, E:].sum(0).to(wsc.dtype)
        dbias = drow.sum(0).to(ctx.bias_dtype)
        return dz, dgate, dx, dwpw, dwsc, dbias


def _wgrad_cl(A, B1, B2, a_cl, b1_cl, batch_sum=False):
    """Mixed-layout weight gradient.  A: planes [B,M,P] or channels-last [B,P,M]; B1 likewise with N1; B2 (or None) is
    channels-last [B,P,N2].  Returns (dW [B,M,N1+N2] fp32, drow [B,M] fp32), or with batch_sum the sums over the batch
    ([M,N1+N2], [M]) straight out of the fixed-order reduction kernel."""
    Bn = A.shape[0]
    M, P = (A.shape[2], A.shape[1]) if a_cl else (A.shape[1], A.shape[2])
    N1 = B1.shape[2] if b1_cl else B1.shape[1]
    N2 = 0 if B2 is None else B2.shape[2]
    dims = L.WgradDims(Bn, M, N1, N2, P)
    lead = () if batch_sum else (Bn,)
    dW = torch.empty(*lead, M, N1 + N2, dtype=torch.float32, device=A.device)
    drow = torch.empty(*lead, M, dtype=torch.float32, device=A.device)
    n = L.lib().lmnet_wgrad_1x1_cl_workspace_bytes(L.byref(dims), int(a_cl), int(b1_cl))
    ws = torch.empty(max(int(n), 16), dtype=torch.uint8, device=A.device)
    fn = L.lib().lmnet_wgrad_1x1_cl_sum if batch_sum else L.lib().lmnet_wgrad_1x1_cl
    rc = fn(L.ptr(A), L.ptr(B1), L.ptr(B2), L.ptr(dW), L.ptr(drow), L.ptr(ws), ws.numel(),
                                    L.byref(dims), int(a_cl), int(b1_cl), L.dtype_code(A), L.stream_ptr())
    L.check(rc, "wgrad_1x1_cl")
    return dW, drow


def wgrad_cl_supported(B, M, N1, N2, P, a_cl, b1_cl, dtype) -> bool:
    if dtype not in (torch.bfloat16, torch.float16):
        return False
    dims = L.WgradDims(B, M, N1, N2, P)
    return bool(L.lib().lmnet_wgrad_1x1_cl_supported(L.byref(dims), int(a_cl), int(b1_cl), L._DTYPES[dtype]))


def _pgemm_dims(B, P, N, K1, K2):
    return L.PgemmDims(B, P, N, K1, K2)


def pgemm_supported(B, P, N, K1, K2, in1_cl, out_cl, stats, dtype) -> bool:
    if dtype not in (torch.bfloat16, torch.float16):
        return False
    return bool(L.lib().lmnet_pixel_gemm_supported(L.byref(_pgemm_dims(B, P, N, K1, K2)), int(in1_cl), int(out_cl),
                                                   int(stats), L._DTYPES[dtype]))


def _f32_param(w):
    w = w.detach()
    return w if w.dtype == torch.float32 else w.float()


def pixel_gemm(in1, in1_cl, w1, in2=None, w2=None, bias=None, out_cl=True, stats=False, gate=None, gate_on_n=False):
    """out[b] = W1 . in1[b] (+ W2 . in2[b]) (+ bias) over the pixels of every image (csrc/pixel_gemm.cu).

    in1: [B,P,K1] (in1_cl) or [B,K1,P] planes; w1: fp32 [N,K1] with ANY strides (a `.t()` view is read in place); in2:
    [B,P,K2] channels-last with w2 fp32 [N,K2]; bias fp32 [N]; gate: fp32 [B,K1] (or [B,N] with gate_on_n) multiplied into
    W1 per image while the kernel stages it.  The kernel rounds the weights to the activation dtype itself, so no cast,
    transpose or gate-multiply kernel runs around the call.  Returns out ([B,P,N] if out_cl else [B,N,P]) and, with
    stats=True, the per-CTA (sum, sum of squares) partials [N, ctas, 2] of the stored output for lmnet_bn_act_fwd_stats."""
    B = in1.shape[0]
    P, K1 = (in1.shape[1], in1.shape[2]) if in1_cl else (in1.shape[2], in1.shape[1])
    N = w1.shape[0]
    K2 = 0 if in2 is None else in2.shape[2]
    dt = in1.dtype
    dims = _pgemm_dims(B, P, N, K1, K2)
    w1 = _f32_param(w1)
    w2 = None if w2 is None else _f32_param(w2)
    gate = None if gate is None else gate.detach().float().contiguous()
    wts = L.PgemmWeights(w1.data_ptr(), w1.stride(0), w1.stride(1), None if gate is None else gate.data_ptr(), int(gate_on_n),
                         None if w2 is None else w2.data_ptr(), 0 if w2 is None else w2.stride(0), 0 if w2 is None else w2.stride(1))
    bias = None if bias is None else bias.detach().float().contiguous()
    out = torch.empty((B, P, N) if out_cl else (B, N, P), dtype=dt, device=in1.device)
    part = None
    if stats:
        ctas = L.lib().lmnet_pixel_gemm_stats_ctas(L.byref(dims), int(in1_cl), int(out_cl))
        part = torch.empty(N, ctas, 2, dtype=torch.float32, device=in1.device)
    rc = L.lib().lmnet_pixel_gemm(L.ptr(in1), int(in1_cl), L.ptr(in2), L.byref(wts), L.ptr(bias), L.ptr(out), int(out_cl),
                                  L.ptr(part), L.byref(dims), L.dtype_code(in1), L.stream_ptr())
    L.check(rc, "pixel_gemm")
    return out, part


def is_channels_last(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous()


def _pixels(x_cl):
    """[B,C,H,W] channels-last tensor -> its memory as [B, H*W, C] (a view)."""
    B, C, H, W = x_cl.shape
    return x_cl.permute(0, 2, 3, 1).reshape(B, H * W, C)


def _as_cl(t_bpc, H, W):
    """[B, H*W, C] contiguous -> logical [B,C,H,W] in channels-last memory (a view)."""
    B, P, C = t_bpc.shape
    return t_bpc.view(B, H, W, C).permute(0, 3, 1, 2)


class _Expand1x1Cl(torch.autograd.Function):
    """x channels-last [B,K,H,W] -> y[b] = W @ x[b]^T + bias as NCHW planes [B,M,H,W] (what the fused BatchNorm and
    the depthwise section read).  The layout change rides on the GEMM's operand order: no transposed copy exists.
    Second output: BatchNorm (sum, sum of squares) partials of y from the GEMM's epilogue (or None)."""

    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, x, w, bias, want_stats):
        B, K, H, Wd = x.shape
        M = w.shape[0]
        P = H * Wd
        xp = _pixels(x)                                                    # [B, P, K]
        part = None
        if pgemm_supported(B, P, M, K, 0, True, False, want_stats, x.dtype):
            y, part = pixel_gemm(xp, True, w, bias=bias, out_cl=False, stats=want_stats)
        else:
            wc = w.to(x.dtype).unsqueeze(0).expand(B, M, K)
            xt = xp.transpose(1, 2)                                        # [B, K, P] view
            y = torch.baddbmm(bias.to(x.dtype).view(1, M, 1), wc, xt) if bias is not None else torch.bmm(wc, xt)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.bias_dtype = None if bias is None else bias.dtype
        if part is not None:
            ctx.mark_non_differentiable(part)
        return y.view(B, M, H, Wd), part

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy, _dpart=None):
        x, w = ctx.saved_tensors
        B, K, H, Wd = x.shape
        M = w.shape[0]
        P = H * Wd
        dyf = dy.to(x.dtype).contiguous().view(B, M, P)
        dx = None
        if ctx.needs_input_grad[0]:
            if pgemm_supported(B, P, K, M, 0, False, True, False, x.dtype):
                dx = _as_cl(pixel_gemm(dyf, False, w.t(), out_cl=True)[0], H, Wd)
            else:
                dx = _as_cl(torch.bmm(dyf.transpose(1, 2), w.to(x.dtype).unsqueeze(0).expand(B, M, K)), H, Wd)
        dW, drow = _wgrad_cl(dyf, _pixels(x), None, False, True, batch_sum=True)
        dw = dW.to(w.dtype)
        db = drow.to(ctx.bias_dtype) if ctx.has_bias else None
        return dx, dw, db, None


class _PointwiseShortcutCl(torch.autograd.Function):
    """out = (Wpw * gate[b]) @ z[b] + Wsc @ x[b] + bias with z as NCHW planes, x and out channels-last."""

    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, z, gate, x, wpw, wsc, bias):
        B, E, H, Wd = z.shape
        Cin, Cout = x.shape[1], wpw.shape[0]
        P = H * Wd
        dt = z.dtype
        xp = _pixels(x)                                                                # [B, P, Cin]
        if pgemm_supported(B, P, Cout, E, Cin, False, True, False, dt):
            out = pixel_gemm(z.view(B, E, P), False, wpw, xp, wsc, bias, out_cl=True, gate=gate)[0]   # W_b = Wpw * gate_b in-kernel
        else:
            wg = (wpw.float().unsqueeze(0) * gate.float().unsqueeze(1)).to(dt)        # [B, Cout, E]
            out = torch.baddbmm(bias.to(dt).view(1, 1, Cout), xp, wsc.to(dt).t().unsqueeze(0).expand(B, Cin, Cout))
            out = torch.baddbmm(out, z.view(B, E, P).transpose(1, 2), wg.transpose(1, 2))  # [B, P, Cout]
        ctx.save_for_backward(z, gate, x, wpw, wsc)
        ctx.bias_dtype = bias.dtype
        return _as_cl(out, H, Wd)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        z, gate, x, wpw, wsc = ctx.saved_tensors
        B, E, H, Wd = z.shape
        Cin, Cout = x.shape[1], wpw.shape[0]
        P = H * Wd
        dt = z.dtype
        do = _pixels(dout.to(dt).contiguous(memory_format=torch.channels_last))        # [B, P, Cout]
        if pgemm_supported(B, P, E, Cout, 0, True, False, False, dt):
            dz = pixel_gemm(do, True, wpw.t(), out_cl=False, gate=gate, gate_on_n=True)[0].view(B, E, H, Wd)
        else:
            wg = (wpw.float().unsqueeze(0) * gate.float().unsqueeze(1)).to(dt)
            dz = torch.bmm(wg.transpose(1, 2), do.transpose(1, 2)).view(B, E, H, Wd)   # planes
        dx = None
        if ctx.needs_input_grad[2]:
            if pgemm_supported(B, P, Cin, Cout, 0, True, True, False, dt):
                dx = _as_cl(pixel_gemm(do, True, wsc.t(), out_cl=True)[0], H, Wd)
            else:
                dx = _as_cl(torch.bmm(do, wsc.to(dt).unsqueeze(0).expand(B, Cout, Cin)), H, Wd)
        dW, drow = _wgrad_cl(do, z.view(B, E, P), _pixels(x), True, False)             # [B, Cout, E + Cin]
        dev = z.device
        w32, g32 = _f32_param(wpw), gate.detach().float().contiguous()
        dwpw = torch.empty(Cout, E, dtype=torch.float32, device=dev)
        dgate = torch.empty(B, E, dtype=torch.float32, device=dev)
        dwsc = torch.empty(Cout, Cin, dtype=torch.float32, device=dev)
        dbias = torch.empty(Cout, dtype=torch.float32, device=dev)
        rc = L.lib().lmnet_pointwise_grads(L.ptr(dW), L.ptr(drow), L.ptr(g32), L.ptr(w32), w32.stride(0), w32.stride(1),
                                           L.ptr(dwpw), L.ptr(dgate), L.ptr(dwsc), L.ptr(dbias), B, Cout, E, Cin,
                                           L.stream_ptr())
        L.check(rc, "pointwise_grads")
        return dz, dgate.to(gate.dtype), dx, dwpw.to(wpw.dtype), dwsc.to(wsc.dtype), dbias.to(ctx.bias_dtype)


def _pad_channels(x, w, mult=4):
    """Zero-pad the channel dimension of x [B,K,H,W] and the matching weight columns [M,K] to a multiple of `mult`
    (the RGB input of the first block: 3 -> 4 channels, so that a pixel is one 8-byte vector)."""
    K = x.shape[1]
    pad = (-K) % mult
    if pad == 0:
        return x, w
    return torch.nn.functional.pad(x, (0, 0, 0, 0, 0, pad)), torch.nn.functional.pad(w, (0, pad))


def _compute_dtype(x):
    if torch.is_autocast_enabled("cuda"):
        return torch.get_autocast_dtype("cuda")
    return x.dtype


def to_channels_last(x: torch.Tensor, dt=None) -> torch.Tensor:
    dt = x.dtype if dt is None else dt
    return x.to(dt).contiguous(memory_format=torch.channels_last)


def expand_1x1(conv: torch.nn.Conv2d, x: torch.Tensor, want_stats: bool = False):
    """conv(x) for a 1x1 convolution on a CUDA tensor; returns NCHW planes.  Channels-last inputs take the mixed-layout
    path (no transposed copy); plain NCHW inputs the plane-wise one; cuDNN through the module for uncovered shapes.
    With want_stats=True returns (y, partials): BatchNorm (sum, sum of squares) partials of y out of the GEMM's epilogue
    for bnact.bn_act(..., stats=partials), or None where the producing kernel does not emit them."""
    L.require_cuda(x)
    dt = _compute_dtype(x)
    B, K, H, W = x.shape
    M = conv.out_channels
    if is_channels_last(x):
        w = conv.weight.view(M, K)
        xpad, wpad = _pad_channels(x.to(dt), w)
        if wgrad_cl_supported(B, M, xpad.shape[1], 0, H * W, False, True, dt):
            if xpad is not x:
                xpad = xpad.contiguous(memory_format=torch.channels_last)
            y, part = _Expand1x1Cl.apply(xpad, wpad, conv.bias, want_stats)
            return (y, part) if want_stats else y
    if wgrad_supported(B, M, K, 0, H * W, dt):
        y = _Expand1x1.apply(x.to(dt).contiguous(), conv.weight.view(M, K), conv.bias)
    else:
        y = conv(x)
    return (y, None) if want_stats else y


def pointwise_shortcut(pw: torch.nn.Conv2d, sc: torch.nn.Conv2d, z, gate, x):
    """pw(gate * z) + sc(x) with gate [B,E] or [B,E,1,1] on CUDA tensors (module calls for uncovered shapes)."""
    L.require_cuda(z, x)
    dt = _compute_dtype(z)
    B, E, H, W = z.shape
    Cin, Cout = x.shape[1], pw.out_channels
    if pw.bias is not None and sc.bias is not None and is_channels_last(x) and z.is_contiguous():
        wsc = sc.weight.view(Cout, Cin)
        xpad, wscpad = _pad_channels(x.to(dt), wsc)
        if wgrad_cl_supported(B, Cout, E, xpad.shape[1], H * W, True, False, dt):
            if xpad is not x:
                xpad = xpad.contiguous(memory_format=torch.channels_last)
            return _PointwiseShortcutCl.apply(z.to(dt), gate.reshape(B, E), xpad, pw.weight.view(Cout, E), wscpad,
                                              pw.bias + sc.bias)
    if pw.bias is not None and sc.bias is not None and wgrad_supported(B, Cout, E, Cin, H * W, dt):
        bias = pw.bias + sc.bias
        return _PointwiseShortcut.apply(z.to(dt).contiguous(), gate.reshape(B, E), x.to(dt).contiguous(),
                                        pw.weight.view(Cout, E), sc.weight.view(Cout, Cin), bias)
    return pw(gate.reshape(B, E, 1, 1).to(z.dtype) * z) + sc(x)
