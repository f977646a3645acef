"""The squeeze-excite gate of ReparamConv on one kernel per direction (``csrc/se_gate.cu``).

Reference: ``SE.forward`` (/root/reference/core/modules.py:1030-1036).  The average pool comes out of the depthwise
kernel and the multiply is folded into the pointwise weights, so this module covers what is left:
``gate = Hardsigmoid(fc2(ReLU(fc1(pool))))`` on [B, E] numbers — ~27 tiny stock launches per block and direction otherwise.
fp32 arithmetic; parameters and their gradients fp32.
"""
from __future__ import annotations

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L


def _f32(t):
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


class _SEGate(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, pool, w1, b1, w2, b2):
        L.require_cuda(pool)
        B, E = pool.shape
        R = w1.shape[0]
        p32, W1, W2 = _f32(pool), _f32(w1).view(R, E), _f32(w2).view(E, R)
        B1 = None if b1 is None else _f32(b1)
        B2 = None if b2 is None else _f32(b2)
        gate = torch.empty(B, E, dtype=torch.float32, device=pool.device)
        h1 = torch.empty(B, R, dtype=torch.float32, device=pool.device)
        pre2 = torch.empty(B, E, dtype=torch.float32, device=pool.device)
        L.dtype_code(p32)                     # device contract
        rc = L.lib().lmnet_se_gate_fwd(L.ptr(p32), L.ptr(W1), L.ptr(B1), L.ptr(W2), L.ptr(B2), L.ptr(gate), L.ptr(h1),
                                       L.ptr(pre2), B, E, R, L.stream_ptr())
        L.check(rc, "se_gate_fwd")
        ctx.save_for_backward(p32, h1, pre2, W1, W2)
        ctx.meta = (pool.dtype, w1.shape, w1.dtype, None if b1 is None else b1.dtype, w2.shape, w2.dtype,
                    None if b2 is None else b2.dtype)
        return gate

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dgate):
        p32, h1, pre2, W1, W2 = ctx.saved_tensors
        B, E = p32.shape
        R = W1.shape[0]
        dev = p32.device
        dg = _f32(dgate)
        dpool = torch.empty(B, E, dtype=torch.float32, device=dev)
        dW1 = torch.empty(R, E, dtype=torch.float32, device=dev)
        db1 = torch.empty(R, dtype=torch.float32, device=dev)
        dW2 = torch.empty(E, R, dtype=torch.float32, device=dev)
        db2 = torch.empty(E, dtype=torch.float32, device=dev)
        rc = L.lib().lmnet_se_gate_bwd(L.ptr(dg), L.ptr(p32), L.ptr(h1), L.ptr(pre2), L.ptr(W1), L.ptr(W2), L.ptr(dpool),
                                       L.ptr(dW1), L.ptr(db1), L.ptr(dW2), L.ptr(db2), B, E, R, L.stream_ptr())
        L.check(rc, "se_gate_bwd")
        pd, w1s, w1d, b1d, w2s, w2d, b2d = ctx.meta
        return (dpool.to(pd), dW1.view(w1s).to(w1d), None if b1d is None else db1.to(b1d), dW2.view(w2s).to(w2d),
                None if b2d is None else db2.to(b2d))


def _is_1x1(conv) -> bool:
    return (type(conv) is torch.nn.Conv2d and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.padding == (0, 0)
            and conv.groups == 1 and conv.dilation == (1, 1))


def se_gate(se: torch.nn.Module, pool: torch.Tensor) -> torch.Tensor:
    """gate [B, E] (fp32) = se.scale_activation(se.fc2(se.activation(se.fc1(pool)))) for pool [B, E]."""
    B, E = pool.shape
    ok = (pool.is_cuda and _is_1x1(se.fc1) and _is_1x1(se.fc2) and type(se.activation) is torch.nn.ReLU
          and type(se.scale_activation) is torch.nn.Hardsigmoid and se.fc1.in_channels == E and se.fc2.out_channels == E
          and se.fc2.in_channels == se.fc1.out_channels
          and bool(L.lib().lmnet_se_gate_supported(B, E, se.fc1.out_channels)))
    if not ok:
        x = pool.view(B, E, 1, 1)
        return se.scale_activation(se.fc2(se.activation(se.fc1(x)))).reshape(B, E)
    return _SEGate.apply(pool, se.fc1.weight, se.fc1.bias, se.fc2.weight, se.fc2.bias)
