"""nn.Linear over channels-last pixels on the pixel-GEMM kernels (widening step f2 of SURVEY.md §8).

The Linear layers around the neighbourhood-attention core — ``qkv`` / ``proj`` of NeighborhoodAttention2D
(/root/reference/core/modules.py:509, 517 -> natten's module) and ``fc1`` / ``fc2`` of the block's Mlp
(/root/reference/core/modules.py:42-56, 511, 519) — act on [B, H, W, C] tensors with C in {12 .. 192}: per pixel a
K <= 192, N <= 288 contraction, i.e. 24 .. 960 B of traffic against at most 110 kFLOP.  They are HBM-bound streaming
kernels, not GEMMs in the cuBLAS sense (whose 128-wide tiles waste > 80 % of the tensor pipe at these widths and whose
split-K weight gradients over 2 M pixels are the slowest kernels of the stock step):

    forward          out = x . W^T + b          csrc/pixel_gemm.cu   (channels-last -> channels-last)
    input gradient   dx  = dout . W             csrc/pixel_gemm.cu
    weight gradient  dW  = dout^T . x, db       csrc/wgrad_1x1.cu    (both operands channels-last, split over pixels)

Anything the kernels do not cover (fp32 storage, odd channel counts, CPU tensors) goes to ``F.linear``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L
from .conv1x1 import _compute_dtype, _wgrad_cl, pgemm_supported, pixel_gemm, wgrad_cl_supported


def _as_bpc(t):
    """[..., C] contiguous -> [B, P, C] view with B = the leading dimension (the grid's batch axis)."""
    C = t.shape[-1]
    B = t.shape[0] if t.dim() > 2 else 1
    return t.reshape(B, t.numel() // (B * C), C)


class _LinearPx(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, x, w, bias):
        xp = _as_bpc(x)
        out, _ = pixel_gemm(xp, True, w, bias=bias, out_cl=True)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.bias_dtype = None if bias is None else bias.dtype
        return out.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        N, K = w.shape
        xp = _as_bpc(x)
        B, P, _ = xp.shape
        do = _as_bpc(dout.to(x.dtype).contiguous())
        dx = None
        if ctx.needs_input_grad[0]:
            if pgemm_supported(B, P, K, N, 0, True, True, False, x.dtype):
                dx = pixel_gemm(do, True, w.t(), out_cl=True)[0].view(x.shape)
            else:
                dx = (do @ w.to(x.dtype)).view(x.shape)
        if wgrad_cl_supported(B, N, K, 0, P, True, True, x.dtype):
            dW, drow = _wgrad_cl(do, xp, None, True, True, batch_sum=True)
            dw = dW.to(w.dtype)
            db = drow.to(ctx.bias_dtype) if ctx.has_bias else None
        else:
            dw = (do.reshape(-1, N).t() @ xp.reshape(-1, K)).to(w.dtype)
            db = do.reshape(-1, N).sum(0).to(ctx.bias_dtype) if ctx.has_bias else None
        return dx, dw, db


def linear_supported(x: torch.Tensor, lin: torch.nn.Linear) -> bool:
    if not x.is_cuda or x.dim() < 2 or x.shape[-1] != lin.in_features:
        return False
    dt = _compute_dtype(x)
    if dt not in (torch.bfloat16, torch.float16) or x.dtype != dt:
        return False
    B = x.shape[0] if x.dim() > 2 else 1
    P = x.numel() // (B * lin.in_features)
    return pgemm_supported(B, P, lin.out_features, lin.in_features, 0, True, True, False, dt)


def linear(lin: torch.nn.Linear, x: torch.Tensor) -> torch.Tensor:
    """lin(x) for a Linear over the last dimension of a 16-bit CUDA tensor; F.linear for everything else."""
    if type(lin) is torch.nn.Linear and linear_supported(x, lin):
        L.require_cuda(x)
        return _LinearPx.apply(x.contiguous(), lin.weight, lin.bias)
    return lin(x)
