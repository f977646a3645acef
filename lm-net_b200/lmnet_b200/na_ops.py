"""Autograd operators for 2-D neighbourhood attention on top of the C ABI.

Mirrors the operator surface of natten 0.14-0.17 that LM-Net reaches through
``NeighborhoodAttention2D`` (/root/reference/core/modules.py:18,509,517):

* ``na2d`` / ``na2d_qkvpacked``  — fused q*scale -> QK+rpb -> softmax -> AV, no attention map,
  backward by recomputation;
* ``na2d_qk`` / ``na2d_av``      — the unfused pair (natten.functional.na2d_qk / na2d_av,
  0.14 spellings natten2dqkrpb / natten2dav), layout [B, heads, H, W, D].

Argument checking follows natten's (odd kernel > 1, H and W >= kernel*dilation, matching shapes)
and raises ValueError before anything is launched.  All compute is CUDA; CPU tensors raise.
"""
from __future__ import annotations

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L


# --------------------------------------------------------------------------------------------
# argument checks (host logic, testable without a GPU)
# --------------------------------------------------------------------------------------------
def check_kernel(kernel_size: int, dilation: int, H: int, W: int) -> None:
    if not isinstance(kernel_size, int) or kernel_size < 3 or kernel_size % 2 != 1:
        raise ValueError(f"kernel_size must be an odd integer > 1, got {kernel_size}")
    if kernel_size > 13:
        raise ValueError(f"kernel_size up to 13 is supported, got {kernel_size}")
    if not isinstance(dilation, int) or dilation < 1:
        raise ValueError(f"dilation must be an integer >= 1, got {dilation}")
    if H < kernel_size * dilation or W < kernel_size * dilation:
        raise ValueError(
            f"input {H}x{W} is smaller than kernel_size*dilation = {kernel_size * dilation}; "
            "pad the input (NeighborhoodAttention2D does this for you)")


def check_rpb(rpb, heads: int, kernel_size: int):
    if rpb is None:
        return None
    want = (heads, 2 * kernel_size - 1, 2 * kernel_size - 1)
    if tuple(rpb.shape) != want:
        raise ValueError(f"rel_pos_bias must have shape {want}, got {tuple(rpb.shape)}")
    return rpb


def _last_contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.stride(-1) == 1 else t.contiguous()


def _rpb32(rpb):
    return None if rpb is None else rpb.detach().float().contiguous()


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------------
# raw launches (no autograd)
# --------------------------------------------------------------------------------------------
def raw_fused_fwd(q, k, v, rpb32, out, K, d, scale, order="bhwnd", lse=None):
    L.require_cuda(q, k, v, out, rpb32)
    B, H, W, heads, D = _logical_shape(q, order)
    dims = L.na_dims(B, H, W, heads, D, K, d)
    vq, vk, vv, vo = (L.view5(t, order) for t in (q, k, v, out))
    rc = L.lib().lmnet_na2d_fwd(L.byref(vq), L.byref(vk), L.byref(vv), L.ptr(rpb32), L.byref(vo), L.ptr(lse),
                                L.byref(dims), float(scale), L.dtype_code(q), L.stream_ptr())
    L.check(rc, "na2d_fwd")


def raw_fused_bwd(q, k, v, rpb32, dout, dq, dk, dv, drpb, K, d, scale, order="bhwnd"):
    L.require_cuda(q, k, v, dout, dq, dk, dv)
    B, H, W, heads, D = _logical_shape(q, order)
    dims = L.na_dims(B, H, W, heads, D, K, d)
    views = [L.view5(t, order) for t in (q, k, v, dout, dq, dk, dv)]
    nbytes = L.lib().lmnet_na2d_bwd_workspace_bytes(L.byref(dims))
    ws = _workspace(nbytes, q.device)
    rc = L.lib().lmnet_na2d_bwd(L.byref(views[0]), L.byref(views[1]), L.byref(views[2]), L.ptr(rpb32),
                                L.byref(views[3]), L.byref(views[4]), L.byref(views[5]), L.byref(views[6]),
                                L.ptr(drpb), L.ptr(ws), ws.numel(), L.byref(dims), float(scale),
                                L.dtype_code(q), L.stream_ptr())
    L.check(rc, "na2d_bwd")


def _logical_shape(t, order):
    s = t.shape
    if order == "bhwnd":
        return s[0], s[1], s[2], s[3], s[4]
    return s[0], s[2], s[3], s[1], s[4]


def _same(q, *others):
    for t in others:
        if t.shape != q.shape:
            raise ValueError(f"q, k, v must have identical shapes, got {tuple(q.shape)} and {tuple(t.shape)}")
        if t.dtype != q.dtype:
            raise ValueError(f"q, k, v must have identical dtypes, got {q.dtype} and {t.dtype}")


# --------------------------------------------------------------------------------------------
# fused op, separate q / k / v   (layout [B, H, W, heads, D])
# --------------------------------------------------------------------------------------------
class _NA2DFused(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, q, k, v, rpb, K, d, scale):
        q, k, v = _last_contig(q), _last_contig(k), _last_contig(v)
        rpb32 = _rpb32(rpb)
        out = torch.empty(q.shape, dtype=q.dtype, device=q.device)
        try:
            raw_fused_fwd(q, k, v, rpb32, out, K, d, scale)
        except RuntimeError as e:  # misaligned views: retry on packed copies (still the CUDA path)
            if "unsupported" not in str(e):
                raise
            q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
            raw_fused_fwd(q, k, v, rpb32, out, K, d, scale)
        ctx.save_for_backward(q, k, v, rpb32)
        ctx.cfg = (K, d, scale, rpb is not None, None if rpb is None else rpb.dtype)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        q, k, v, rpb32 = ctx.saved_tensors
        K, d, scale, has_rpb, rpb_dtype = ctx.cfg
        dout = dout.to(q.dtype).contiguous()
        dq, dk, dv = (torch.empty(q.shape, dtype=q.dtype, device=q.device) for _ in range(3))
        drpb = torch.empty_like(rpb32) if has_rpb else None
        try:
            raw_fused_bwd(q, k, v, rpb32, dout, dq, dk, dv, drpb, K, d, scale)
        except RuntimeError as e:
            if "unsupported" not in str(e):
                raise
            q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
            raw_fused_bwd(q, k, v, rpb32, dout, dq, dk, dv, drpb, K, d, scale)
        if drpb is not None:
            drpb = drpb.to(rpb_dtype)
        return dq, dk, dv, drpb, None, None, None


def na2d(query, key, value, kernel_size, dilation=1, rel_pos_bias=None, scale=None):
    """Fused neighbourhood attention.  query/key/value: [B, H, W, heads, head_dim]."""
    if query.dim() != 5:
        raise ValueError("na2d expects [B, H, W, heads, head_dim] tensors")
    _same(query, key, value)
    B, H, W, heads, D = query.shape
    check_kernel(kernel_size, dilation, H, W)
    check_rpb(rel_pos_bias, heads, kernel_size)
    scale = float(D) ** -0.5 if scale is None else float(scale)
    return _NA2DFused.apply(query, key, value, rel_pos_bias, kernel_size, dilation, scale)


# --------------------------------------------------------------------------------------------
# fused op on a packed qkv tensor [B, H, W, 3, heads, D] (what the module's Linear produces)
# --------------------------------------------------------------------------------------------
class _NA2DPacked(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, qkv, rpb, K, d, scale):
        qkv = qkv.contiguous()
        B, H, W, _, heads, D = qkv.shape
        rpb32 = _rpb32(rpb)
        out = torch.empty(B, H, W, heads, D, dtype=qkv.dtype, device=qkv.device)
        raw_fused_fwd(qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2], rpb32, out, K, d, scale)
        ctx.save_for_backward(qkv, rpb32)
        ctx.cfg = (K, d, scale, rpb is not None, None if rpb is None else rpb.dtype)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        qkv, rpb32 = ctx.saved_tensors
        K, d, scale, has_rpb, rpb_dtype = ctx.cfg
        dout = dout.to(qkv.dtype).contiguous()
        dqkv = torch.empty_like(qkv)
        drpb = torch.empty_like(rpb32) if has_rpb else None
        raw_fused_bwd(qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2], rpb32, dout,
                      dqkv[:, :, :, 0], dqkv[:, :, :, 1], dqkv[:, :, :, 2], drpb, K, d, scale)
        if drpb is not None:
            drpb = drpb.to(rpb_dtype)
        return dqkv, drpb, None, None, None


def na2d_qkvpacked(qkv, kernel_size, dilation=1, rel_pos_bias=None, scale=None):
    """Fused neighbourhood attention reading q, k, v straight out of a packed projection.

    qkv: [B, H, W, 3, heads, head_dim] (the natural output layout of the module's qkv Linear);
    returns [B, H, W, heads, head_dim]."""
    if qkv.dim() != 6 or qkv.shape[3] != 3:
        raise ValueError("na2d_qkvpacked expects [B, H, W, 3, heads, head_dim]")
    B, H, W, _, heads, D = qkv.shape
    check_kernel(kernel_size, dilation, H, W)
    check_rpb(rel_pos_bias, heads, kernel_size)
    scale = float(D) ** -0.5 if scale is None else float(scale)
    return _NA2DPacked.apply(qkv, rel_pos_bias, kernel_size, dilation, scale)


# --------------------------------------------------------------------------------------------
# unfused ops, natten layout [B, heads, H, W, D] / [B, heads, H, W, K*K]
# --------------------------------------------------------------------------------------------
def raw_qk_fwd(q, k, rpb32, attn, K, d):
    L.require_cuda(q, k, attn, rpb32)
    B, heads, H, W, D = q.shape
    dims = L.na_dims(B, H, W, heads, D, K, d)
    vq, vk = L.view5(q, "bnhwd"), L.view5(k, "bnhwd")
    L.check(L.lib().lmnet_na2d_qk_fwd(L.byref(vq), L.byref(vk), L.ptr(rpb32), L.ptr(attn), L.byref(dims),
                                      L.dtype_code(q), L.stream_ptr()), "na2d_qk_fwd")


def raw_qk_bwd(q, k, dattn, dq, dk, drpb, K, d):
    L.require_cuda(q, k, dattn, dq, dk)
    B, heads, H, W, D = q.shape
    dims = L.na_dims(B, H, W, heads, D, K, d)
    vq, vk, vdq, vdk = (L.view5(t, "bnhwd") for t in (q, k, dq, dk))
    ws = _workspace(L.lib().lmnet_na2d_qk_bwd_workspace_bytes(L.byref(dims)) if drpb is not None else 0, q.device)
    L.check(L.lib().lmnet_na2d_qk_bwd(L.byref(vq), L.byref(vk), L.ptr(dattn), L.byref(vdq), L.byref(vdk), L.ptr(drpb),
                                      L.ptr(ws), ws.numel(), L.byref(dims), L.dtype_code(q), L.stream_ptr()),
            "na2d_qk_bwd")


def raw_av_fwd(attn, v, out, K, d):
    L.require_cuda(attn, v, out)
    B, heads, H, W, D = v.shape
    dims = L.na_dims(B, H, W, heads, D, K, d)
    vv, vo = L.view5(v, "bnhwd"), L.view5(out, "bnhwd")
    L.check(L.lib().lmnet_na2d_av_fwd(L.ptr(attn), L.byref(vv), L.byref(vo), L.byref(dims), L.dtype_code(v),
                                      L.stream_ptr()), "na2d_av_fwd")


def raw_av_bwd(attn, v, dout, dattn, dv, K, d):
    L.require_cuda(attn, v, dout, dattn, dv)
    B, heads, H, W, D = v.shape
    dims = L.na_dims(B, H, W, heads, D, K, d)
    vv, vg, vdv = (L.view5(t, "bnhwd") for t in (v, dout, dv))
    L.check(L.lib().lmnet_na2d_av_bwd(L.ptr(attn), L.byref(vv), L.byref(vg), L.ptr(dattn), L.byref(vdv), L.byref(dims),
                                      L.dtype_code(v), L.stream_ptr()), "na2d_av_bwd")


class _NA2DQK(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, q, k, rpb, K, d):
        q, k = _last_contig(q), _last_contig(k)
        rpb32 = _rpb32(rpb)
        B, heads, H, W, D = q.shape
        attn = torch.empty(B, heads, H, W, K * K, dtype=q.dtype, device=q.device)
        raw_qk_fwd(q, k, rpb32, attn, K, d)
        ctx.save_for_backward(q, k)
        ctx.cfg = (K, d, rpb is not None, None if rpb is None else rpb.dtype)
        return attn

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dattn):
        q, k = ctx.saved_tensors
        K, d, has_rpb, rpb_dtype = ctx.cfg
        dattn = dattn.to(q.dtype).contiguous()
        dq = torch.empty(q.shape, dtype=q.dtype, device=q.device)
        dk = torch.empty(q.shape, dtype=q.dtype, device=q.device)
        heads = q.shape[1]
        drpb = torch.empty(heads, 2 * K - 1, 2 * K - 1, dtype=torch.float32, device=q.device) if has_rpb else None
        raw_qk_bwd(q, k, dattn, dq, dk, drpb, K, d)
        return dq, dk, (None if drpb is None else drpb.to(rpb_dtype)), None, None


class _NA2DAV(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, attn, v, K, d):
        v = _last_contig(v)
        attn = attn.to(v.dtype).contiguous()
        out = torch.empty(v.shape, dtype=v.dtype, device=v.device)
        raw_av_fwd(attn, v, out, K, d)
        ctx.save_for_backward(attn, v)
        ctx.cfg = (K, d)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        attn, v = ctx.saved_tensors
        K, d = ctx.cfg
        dout = _last_contig(dout.to(v.dtype))
        dattn = torch.empty_like(attn)
        dv = torch.empty(v.shape, dtype=v.dtype, device=v.device)
        raw_av_bwd(attn, v, dout, dattn, dv, K, d)
        return dattn, dv, None, None


def na2d_qk(query, key, kernel_size, dilation=1, rel_pos_bias=None):
    """attn[b,h,i,j,n] = q . k_n (+ rpb).  query/key: [B, heads, H, W, head_dim] -> [B, heads, H, W, K*K]."""
    if query.dim() != 5:
        raise ValueError("na2d_qk expects [B, heads, H, W, head_dim] tensors")
    _same(query, key)
    B, heads, H, W, D = query.shape
    check_kernel(kernel_size, dilation, H, W)
    check_rpb(rel_pos_bias, heads, kernel_size)
    return _NA2DQK.apply(query, key, rel_pos_bias, kernel_size, dilation)


def na2d_av(attn, value, kernel_size, dilation=1):
    """out = sum_n attn_n * v_n.  attn: [B, heads, H, W, K*K], value: [B, heads, H, W, head_dim]."""
    if value.dim() != 5 or attn.dim() != 5:
        raise ValueError("na2d_av expects attn [B, heads, H, W, K*K] and value [B, heads, H, W, head_dim]")
    B, heads, H, W, D = value.shape
    check_kernel(kernel_size, dilation, H, W)
    if tuple(attn.shape) != (B, heads, H, W, kernel_size * kernel_size):
        raise ValueError(f"attn must have shape {(B, heads, H, W, kernel_size * kernel_size)}, got {tuple(attn.shape)}")
    return _NA2DAV.apply(attn, value, kernel_size, dilation)
