"""Bilinear x2 up-sampling (align_corners=True) on the sm_100a kernels of ``csrc/upsample.cu`` — widening
step f3 of SURVEY.md §8: the ``nn.Upsample`` of the decoder (/root/reference/core/LM_Net.py:58-74) and of the
skip-fusion blocks (/root/reference/core/modules.py:93-95, 129-131).  Stays in the storage type (autocast's
stock op converts to fp32 and back), interpolation arithmetic in fp32 with ATen's index rule."""
from __future__ import annotations

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L


class _Up2x(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, x):
        L.require_cuda(x)
        x = x.contiguous()
        B, C, H, W = x.shape
        y = torch.empty(B, C, 2 * H, 2 * W, dtype=x.dtype, device=x.device)
        dims = L.UpsampleDims(B * C, H, W)
        L.check(L.lib().lmnet_upsample2x_fwd(L.ptr(x), L.ptr(y), L.byref(dims), L.dtype_code(x), L.stream_ptr()),
                "upsample2x_fwd")
        ctx.shape = (B, C, H, W)
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        B, C, H, W = ctx.shape
        dy = dy.contiguous()
        dx = torch.empty(B, C, H, W, dtype=dy.dtype, device=dy.device)
        dims = L.UpsampleDims(B * C, H, W)
        L.check(L.lib().lmnet_upsample2x_bwd(L.ptr(dy), L.ptr(dx), L.byref(dims), L.dtype_code(dy), L.stream_ptr()),
                "upsample2x_bwd")
        return dx


def is_channels_last(t: torch.Tensor) -> bool:
    """True when a 4-D tensor's memory is [B, H, W, C] (and it is not also plain-contiguous, e.g. C == 1)."""
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous()


class _Up2xCl(torch.autograd.Function):
    """Same operator on channels-last memory ([B,H,W,C]); output and gradient are channels-last too."""

    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, x):
        L.require_cuda(x)
        B, C, H, W = x.shape
        y = torch.empty(B, C, 2 * H, 2 * W, dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        dims = L.UpsampleClDims(B, H, W, C)
        L.check(L.lib().lmnet_upsample2x_cl_fwd(L.ptr(x), L.ptr(y), L.byref(dims), L.dtype_code(x), L.stream_ptr()),
                "upsample2x_cl_fwd")
        ctx.shape = (B, C, H, W)
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        B, C, H, W = ctx.shape
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty(B, C, H, W, dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        dims = L.UpsampleClDims(B, H, W, C)
        L.check(L.lib().lmnet_upsample2x_cl_bwd(L.ptr(dy), L.ptr(dx), L.byref(dims), L.dtype_code(dy), L.stream_ptr()),
                "upsample2x_cl_bwd")
        return dx


class Upsample2x(torch.nn.Module):
    """Drop-in for nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) (no parameters, so
    state_dicts are unaffected).  CUDA only, like every operator of the package: a CPU tensor raises.  fp32 / bf16 /
    fp16 NCHW tensors run the fused kernel; other CUDA dtypes or ranks go to ATen's kernel."""

    scale_factor, mode, align_corners = 2, "bilinear", True

    def forward(self, x):
        L.require_cuda(x)
        if x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16, torch.float16):
            return _Up2xCl.apply(x) if is_channels_last(x) else _Up2x.apply(x)
        return torch.nn.functional.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)

    def extra_repr(self):
        return "scale_factor=2, mode='bilinear', align_corners=True"
