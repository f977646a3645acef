"""Average pooling by an integer factor on channels-last 16-bit tensors (``csrc/pool_nhwc.cu``): the
``nn.AdaptiveAvgPool2d`` calls of PyramidPool (/root/reference/core/modules.py:454-498), whose input sizes are exact
multiples (16 / 8 / 4 / 2) of the bottleneck resolution.  Anything else goes to ``F.adaptive_avg_pool2d``."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L


class _AvgPoolCl(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, x, f):
        B, C, H, W = x.shape
        Ho, Wo = H // f, W // f
        y = torch.empty((B, C, Ho, Wo), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        dims = L.PoolDims(B, Ho, Wo, C, f)
        L.check(L.lib().lmnet_avgpool_cl_fwd(L.ptr(x), L.ptr(y), L.byref(dims), L.dtype_code(x), L.stream_ptr()), "avgpool_cl_fwd")
        ctx.f, ctx.shape = f, x.shape
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        B, C, H, W = ctx.shape
        f = ctx.f
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((B, C, H, W), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        dims = L.PoolDims(B, H // f, W // f, C, f)
        L.check(L.lib().lmnet_avgpool_cl_bwd(L.ptr(dy), L.ptr(dx), L.byref(dims), L.dtype_code(dy), L.stream_ptr()), "avgpool_cl_bwd")
        return dx, None


def adaptive_avg_pool(x: torch.Tensor, size) -> torch.Tensor:
    """F.adaptive_avg_pool2d(x, size); the sm_100a kernel when x is a channels-last 16-bit CUDA tensor whose height and
    width are the same exact multiple of `size`."""
    Ho, Wo = (size, size) if isinstance(size, int) else tuple(size)
    if x.is_cuda and x.dim() == 4 and x.dtype in (torch.bfloat16, torch.float16) and x.shape[1] % 4 == 0:
        B, C, H, W = x.shape
        f = H // Ho if Ho > 0 else 0
        if f >= 1 and H == Ho * f and W == Wo * f and x.is_contiguous(memory_format=torch.channels_last) and (C > 1):
            if f == 1:
                return x
            return _AvgPoolCl.apply(x, f)
    return F.adaptive_avg_pool2d(x, (Ho, Wo))
