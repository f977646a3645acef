"""Fused BatchNorm2d + activation (widening step f1/f3 of SURVEY.md §8) on the sm_100a kernels of
``csrc/bn_act.cu``.  Drop-in for ``activation(bn(y))`` where ``bn`` is an ``nn.BatchNorm2d``:

* ReparamConv's ``expand_conv``: Conv1x1 -> BatchNorm2d -> Hardswish (/root/reference/core/modules.py:537-539)
* M2Skip / M3Skip ``fuse_conv``: Conv3x3 -> BatchNorm2d -> GELU (/root/reference/core/modules.py:97-100, 122-137)

Semantics are nn.BatchNorm2d's: batch statistics + in-place running-stat / counter update in training,
running statistics in eval; parameters, statistics and gradients of gamma/beta in fp32.
"""
from __future__ import annotations

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L


def _dims(y):
    B, C = y.shape[0], y.shape[1]
    return L.BnDims(B, C, y.numel() // (B * C))


def _ws(dims, dev):
    n = L.lib().lmnet_bn_act_workspace_bytes(L.byref(dims))
    return torch.empty(max(int(n), 16), dtype=torch.uint8, device=dev)


def _f32(t):
    if t is None:
        return None
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


def _is_cl(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous()


def _cl_ok(y) -> bool:
    """channels-last tensor that the channels-last kernels cover (otherwise: NCHW kernels on a contiguous copy)."""
    if not _is_cl(y) or y.dtype not in L._DTYPES:
        return False
    return bool(L.lib().lmnet_bn_act_cl_supported(L.byref(_dims(y)), L._DTYPES[y.dtype]))


class _BNActTrain(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, y, gamma, beta, running, eps, momentum, act, stats=None):
        L.require_cuda(y)
        ctx.cl = _cl_ok(y)
        if not ctx.cl:
            y = y.contiguous()
        g32, b32 = _f32(gamma), _f32(beta)
        rmean, rvar, nbt = running
        C = y.shape[1]
        dims = _dims(y)
        out = torch.empty_like(y)          # preserves the memory format (channels-last stays channels-last)
        save_mean = torch.empty(C, dtype=torch.float32, device=y.device)
        save_rstd = torch.empty(C, dtype=torch.float32, device=y.device)
        ws = _ws(dims, y.device)
        if stats is not None and not ctx.cl:
            # the producer's epilogue already reduced (sum, sum of squares) per CTA: finalize + apply only
            if stats.dim() != 3 or stats.shape[0] != C or stats.shape[2] != 2 or stats.dtype != torch.float32:
                raise ValueError("stats must be fp32 [C, chunks, 2] partial sums")
            rc = L.lib().lmnet_bn_act_fwd_stats(
                L.ptr(y), L.ptr(stats.contiguous()), int(stats.shape[1]), L.ptr(g32), L.ptr(b32), L.ptr(rmean), L.ptr(rvar),
                L.ptr(nbt), L.ptr(out), L.ptr(save_mean), L.ptr(save_rstd), float(eps), float(momentum), act,
                L.ptr(ws), ws.numel(), L.byref(dims), L.dtype_code(y), L.stream_ptr())
        else:
            fn = L.lib().lmnet_bn_act_cl_fwd if ctx.cl else L.lib().lmnet_bn_act_fwd
            rc = fn(L.ptr(y), L.ptr(g32), L.ptr(b32), L.ptr(rmean), L.ptr(rvar), L.ptr(nbt),
                    L.ptr(out), L.ptr(save_mean), L.ptr(save_rstd), float(eps), float(momentum),
                    1, act, L.ptr(ws), ws.numel(), L.byref(dims), L.dtype_code(y), L.stream_ptr())
        L.check(rc, "bn_act_fwd")
        ctx.save_for_backward(y, g32, b32, save_mean, save_rstd)
        ctx.act = act
        ctx.meta = (None if gamma is None else gamma.dtype, None if beta is None else beta.dtype)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        y, g32, b32, save_mean, save_rstd = ctx.saved_tensors
        dims = _dims(y)
        C = y.shape[1]
        dout = dout.to(y.dtype)
        dout = dout.contiguous(memory_format=torch.channels_last) if ctx.cl else dout.contiguous()
        dy = torch.empty_like(y)
        dgamma = torch.empty(C, dtype=torch.float32, device=y.device)
        dbeta = torch.empty(C, dtype=torch.float32, device=y.device)
        ws = _ws(dims, y.device)
        fn = L.lib().lmnet_bn_act_cl_bwd if ctx.cl else L.lib().lmnet_bn_act_bwd
        rc = fn(L.ptr(y), L.ptr(dout), L.ptr(g32), L.ptr(b32), L.ptr(save_mean),
                L.ptr(save_rstd), L.ptr(dy), L.ptr(dgamma), L.ptr(dbeta), ctx.act,
                L.ptr(ws), ws.numel(), L.byref(dims), L.dtype_code(y), L.stream_ptr())
        L.check(rc, "bn_act_bwd")
        gd, bd = ctx.meta
        return (dy, None if gd is None else dgamma.to(gd), None if bd is None else dbeta.to(bd), None, None, None, None, None)


def bn_act(bn: torch.nn.BatchNorm2d, y: torch.Tensor, act: str = "none", stats=None) -> torch.Tensor:
    """activation(bn(y)) in one pass pair.  `act` in {"none", "hardswish", "gelu", "relu"}.  `stats`: optional fp32
    [C, chunks, 2] per-chunk (sum, sum of squares) of y already reduced by its producer (conv1x1.expand_1x1); used in
    training mode on NCHW planes, where it replaces the statistics pass."""
    code = L.ACT_CODES[act]
    if y.dim() < 2 or y.shape[1] != bn.num_features:
        raise ValueError(f"expected {bn.num_features} channels, got input of shape {tuple(y.shape)}")
    tracked = bn.running_mean is not None
    if bn.training or not tracked:
        if bn.momentum is None:
            raise NotImplementedError("cumulative-average BatchNorm (momentum=None) is not supported by the fused path")
        update = bn.training and tracked
        if update:
            for t in (bn.running_mean, bn.running_var):
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise NotImplementedError(f"fused BatchNorm update needs contiguous float32 running statistics, got "
                                              f"{t.dtype} (keep BatchNorm buffers in fp32: autocast, not model.half())")
            if bn.num_batches_tracked is not None and bn.num_batches_tracked.dtype != torch.int64:
                raise NotImplementedError("num_batches_tracked must be int64")
        running = (bn.running_mean, bn.running_var, bn.num_batches_tracked) if update else (None, None, None)
        return _BNActTrain.apply(y, bn.weight, bn.bias, running, bn.eps, bn.momentum, code, stats)
    if torch.is_grad_enabled() and (y.requires_grad or (bn.weight is not None and bn.weight.requires_grad)):
        raise NotImplementedError("gradients through eval-mode fused BatchNorm are not implemented; use torch.no_grad()")
    L.require_cuda(y)
    cl = _cl_ok(y)
    if not cl:
        y = y.contiguous()
    dims = _dims(y)
    out = torch.empty_like(y)
    ws = _ws(dims, y.device)
    fn = L.lib().lmnet_bn_act_cl_fwd if cl else L.lib().lmnet_bn_act_fwd
    rc = fn(L.ptr(y), L.ptr(_f32(bn.weight)), L.ptr(_f32(bn.bias)), L.ptr(_f32(bn.running_mean)),
            L.ptr(_f32(bn.running_var)), None, L.ptr(out), None, None, float(bn.eps), 0.0, 0, code,
            L.ptr(ws), ws.numel(), L.byref(dims), L.dtype_code(y), L.stream_ptr())
    L.check(rc, "bn_act_fwd")
    return out


def conv_bn_act(seq: torch.nn.Sequential, x: torch.Tensor) -> torch.Tensor:
    """Runs a (Conv2d, BatchNorm2d, activation) Sequential with the BN + activation fused.  Falls back to the
    Sequential itself for any other structure (e.g. user-modified blocks)."""
    if len(seq) == 3 and isinstance(seq[0], torch.nn.Conv2d) and type(seq[1]) is torch.nn.BatchNorm2d:
        act = {torch.nn.Hardswish: "hardswish", torch.nn.GELU: "gelu", torch.nn.ReLU: "relu"}.get(type(seq[2]))
        if act == "gelu" and getattr(seq[2], "approximate", "none") != "none":
            act = None
        if act is not None:
            return bn_act(seq[1], seq[0](x), act)
    return seq(x)
