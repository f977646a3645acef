"""LayerNorm over short channels-last rows on the sm_100a kernels of ``csrc/layer_norm.cu`` (widening
step f2 of SURVEY.md §8): the ``norm1`` / ``norm2`` of NeighborhoodTransformer
(/root/reference/core/modules.py:507, 510, 515-518), C in {12, 24, 48, 96}.

The output keeps the input's storage type.  Under autocast the reference's LayerNorm returns fp32 that the
next Linear immediately casts back to bf16; writing bf16 directly yields the same values into that Linear
and saves the two cast passes."""
from __future__ import annotations

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L


def _f32(t):
    if t is None:
        return None
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, x, gamma, beta, eps):
        L.require_cuda(x)
        x = x.contiguous()
        C = x.shape[-1]
        rows = x.numel() // C
        g32, b32 = _f32(gamma), _f32(beta)
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        L.check(L.lib().lmnet_layer_norm_fwd(L.ptr(x), L.ptr(g32), L.ptr(b32), L.ptr(y), L.ptr(mean), L.ptr(rstd), rows,
                                             C, float(eps), L.dtype_code(x), L.stream_ptr()), "layer_norm_fwd")
        ctx.save_for_backward(x, g32, mean, rstd)
        ctx.meta = (None if gamma is None else gamma.dtype, None if beta is None else beta.dtype)
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, g32, mean, rstd = ctx.saved_tensors
        C = x.shape[-1]
        rows = x.numel() // C
        dy = dy.to(x.dtype).contiguous()
        dx = torch.empty_like(x)
        dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
        nbytes = L.lib().lmnet_layer_norm_workspace_bytes(rows, C)
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=x.device)
        L.check(L.lib().lmnet_layer_norm_bwd(L.ptr(x), L.ptr(dy), L.ptr(g32), L.ptr(mean), L.ptr(rstd), L.ptr(dx),
                                             L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws), ws.numel(), rows, C,
                                             L.dtype_code(x), L.stream_ptr()), "layer_norm_bwd")
        gd, bd = ctx.meta
        return dx, (None if gd is None else dgamma.to(gd)), (None if bd is None else dbeta.to(bd)), None


def supported(C: int) -> bool:
    return bool(L.lib().lmnet_layer_norm_supported(int(C)))


def layer_norm(ln: torch.nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    """ln(x) for a LayerNorm over the last dimension; uses the fused kernel for the channel counts of
    LM-Net's neighbourhood transformers and the stock module otherwise."""
    if len(ln.normalized_shape) == 1 and x.shape[-1] == ln.normalized_shape[0] and supported(x.shape[-1]):
        return _LayerNorm.apply(x, ln.weight, ln.bias, ln.eps)
    return ln(x)
