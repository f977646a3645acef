"""Fused middle section of LM-Net's ``ReparamConv`` on the sm_100a kernels.

Reference semantics being replaced (/root/reference/core/modules.py:592-597, 1030-1036):

    out  = large_conv(x1) + square_conv(x1) + ver_conv(x1) + hor_conv(x1)   # 4 x (depthwise conv + BN)
    x1   = se(active(out))                                                   # GELU, avg-pool, gate, multiply

``fused_dw_bn_gelu`` returns ``z = GELU(out)`` and ``pool = mean_hw(z)`` from one C-ABI call per
direction; the squeeze-excite gate (two tiny 1x1 convs on [B,E,1,1]) and the surrounding 1x1
convolutions stay on torch (SURVEY.md §8 f1 lists them as the next widening step).

``patch_reparam_conv(cls)`` swaps ``forward`` on a ReparamConv class object (the reference's or
ours) — constructor, sub-module names, ``fuse_bn`` / ``get_equivalent_kernel_bias`` /
``switch_to_deploy`` and therefore checkpoints are untouched.
"""
from __future__ import annotations

import ctypes

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L
from .bnact import bn_act, conv_bn_act
from .conv1x1 import _compute_dtype, expand_1x1, is_channels_last, pointwise_shortcut, to_channels_last
from .se import se_gate


def _dims(x):
    B, E, H, W = x.shape
    return L.DwDims(B, E, H, W)


def _ws(dims, x):
    n = L.lib().lmnet_reparam_dw_workspace_bytes(L.byref(dims), L.dtype_code(x))
    return torch.empty(max(int(n), 16), dtype=torch.uint8, device=x.device)


def _f32(t):
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


class _FusedDwTrain(torch.autograd.Function):
    """x, 4 depthwise weights, 4 x (gamma, beta) -> z = GELU(sum_br BN_br(conv_br(x))), pool = mean_hw z.

    Batch statistics; running_mean / running_var / num_batches_tracked of the four BatchNorms are
    updated in place on the device exactly like nn.BatchNorm2d in training mode."""

    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, x, w5, w3, w31, w13, g0, b0, g1, b1, g2, b2, g3, b3, running, eps, momentum):
        L.require_cuda(x)
        x = x.contiguous()
        ws_f = [_f32(w5), _f32(w3), _f32(w31), _f32(w13)]
        gam = [_f32(g0), _f32(g1), _f32(g2), _f32(g3)]
        bet = [_f32(b0), _f32(b1), _f32(b2), _f32(b3)]
        rmean, rvar, nbt = running
        B, E, H, W = x.shape
        dims = _dims(x)
        u = torch.empty_like(x)
        z = torch.empty_like(x)
        pool = torch.empty(B, E, dtype=torch.float32, device=x.device)
        save_mean = torch.empty(4, E, dtype=torch.float32, device=x.device)
        save_rstd = torch.empty(4, E, dtype=torch.float32, device=x.device)
        params = L.dw_params(ws_f, gam, bet, rmean, rvar)
        nbt_arr = None
        if nbt is not None:
            nbt_arr = (ctypes.c_void_p * 4)(*[None if t is None else t.data_ptr() for t in nbt])
        ws = _ws(dims, x)
        # lag sums of the branch outputs taken by the statistics pass: with them the backward is one composite stencil
        # pass instead of a recomputation of the four branches (csrc/reparam_dw_tma2.cuh); the library says whether
        # this shape / dtype / alignment has that path
        gram = torch.empty(E, L.lib().lmnet_reparam_dw_gram_floats(), dtype=torch.float32, device=x.device)
        gram_saved = ctypes.c_int(0)
        rc = L.lib().lmnet_reparam_dw_train_fwd_gram(
            L.ptr(x), L.byref(params), L.ptr(u), L.ptr(z), L.ptr(pool), L.ptr(save_mean), L.ptr(save_rstd),
            float(eps), float(momentum), nbt_arr, L.ptr(gram), ctypes.byref(gram_saved), L.ptr(ws), ws.numel(),
            L.byref(dims), L.dtype_code(x), L.stream_ptr())
        L.check(rc, "reparam_dw_train_fwd_gram")
        ctx.has_gram = bool(gram_saved.value)
        ctx.save_for_backward(x, u, save_mean, save_rstd, *ws_f, *gam, *([gram] if ctx.has_gram else []))
        ctx.param_meta = [(t.shape, t.dtype) for t in (w5, w3, w31, w13, g0, b0, g1, b1, g2, b2, g3, b3)]
        ctx.set_materialize_grads(False)
        return z, pool

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dz, dpool):
        x, u, save_mean, save_rstd, w5, w3, w31, w13, g0, g1, g2, g3 = ctx.saved_tensors[:12]
        gram = ctx.saved_tensors[12] if ctx.has_gram else None
        B, E, H, W = x.shape
        dims = _dims(x)
        dz = torch.zeros_like(x) if dz is None else dz.to(x.dtype).contiguous()
        dpool = None if dpool is None else dpool.float().contiguous()
        dx = torch.empty_like(x)
        dev = x.device
        dws = [torch.empty(E, n, dtype=torch.float32, device=dev) for n in (25, 9, 3, 3)]
        dgs = [torch.empty(E, dtype=torch.float32, device=dev) for _ in range(4)]
        dbs = [torch.empty(E, dtype=torch.float32, device=dev) for _ in range(4)]
        zeros = [torch.zeros(E, dtype=torch.float32, device=dev)] * 4  # beta is not needed by the backward
        params = L.dw_params([w5, w3, w31, w13], [g0, g1, g2, g3], zeros)
        grads = L.dw_grads(dws, dgs, dbs)
        ws = _ws(dims, x)
        rc = L.lib().lmnet_reparam_dw_train_bwd_gram(
            L.ptr(x), L.ptr(u), L.ptr(dz), L.ptr(dpool), L.byref(params), L.ptr(save_mean), L.ptr(save_rstd),
            L.ptr(gram), L.ptr(dx), L.byref(grads), L.ptr(ws), ws.numel(), L.byref(dims), L.dtype_code(x), L.stream_ptr())
        L.check(rc, "reparam_dw_train_bwd_gram")
        meta = ctx.param_meta
        outs = [dx]
        for i, t in enumerate(dws):
            outs.append(t.view(meta[i][0]).to(meta[i][1]))
        for k in range(4):
            outs.append(dgs[k].to(meta[4 + 2 * k][1]))
            outs.append(dbs[k].to(meta[5 + 2 * k][1]))
        return (*outs, None, None, None)


def _eval_call(x, params, bias, eps):
    L.require_cuda(x)
    x = x.contiguous()
    B, E, H, W = x.shape
    dims = _dims(x)
    z = torch.empty_like(x)
    pool = torch.empty(B, E, dtype=torch.float32, device=x.device)
    ws = _ws(dims, x)
    rc = L.lib().lmnet_reparam_dw_eval_fwd(L.ptr(x), L.byref(params), L.ptr(bias), float(eps), L.ptr(z), L.ptr(pool),
                                           L.ptr(ws), ws.numel(), L.byref(dims), L.dtype_code(x), L.stream_ptr())
    L.check(rc, "reparam_dw_eval_fwd")
    return z, pool


def _branches(mod):
    return [mod.large_conv, mod.square_conv, mod.ver_conv, mod.hor_conv]


def _check_branch_shapes(mod):
    E = mod.large_conv.conv.weight.shape[0]
    want = [(E, 1, 5, 5), (E, 1, 3, 3), (E, 1, 3, 1), (E, 1, 1, 3)]
    for br, shape in zip(_branches(mod), want):
        c = br.conv
        if tuple(c.weight.shape) != shape or c.stride != (1, 1) or c.groups != E or c.bias is not None:
            raise NotImplementedError(
                "fused ReparamConv supports the LM-Net configuration only: depthwise 5x5 / 3x3 / 3x1 / 1x3, "
                f"stride 1, no bias (got weight {tuple(c.weight.shape)}, stride {c.stride}, groups {c.groups})")


def _check_running_buffers(bns):
    """The finalize kernels update running_mean / running_var as fp32 and num_batches_tracked as int64 in place."""
    for bn in bns:
        for t in (bn.running_mean, bn.running_var):
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise NotImplementedError(f"fused BatchNorm update needs contiguous float32 running statistics, got {t.dtype} "
                                          "(keep BatchNorm buffers in fp32, e.g. use autocast instead of model.half())")
        nbt = bn.num_batches_tracked
        if nbt is not None and nbt.dtype != torch.int64:
            raise NotImplementedError(f"num_batches_tracked must be int64, got {nbt.dtype}")


def fused_dw_bn_gelu(mod, x1):
    """z, pool for the non-deploy ReparamConv `mod` (four conv+BN branches) applied to x1 [B,E,H,W]."""
    _check_branch_shapes(mod)
    brs = _branches(mod)
    bns = [b.bn for b in brs]
    eps = bns[0].eps
    if any(bn.eps != eps for bn in bns):
        raise NotImplementedError("the four BatchNorms must share eps")
    tracked = all(bn.running_mean is not None for bn in bns)
    # nn.BatchNorm2d decides batch vs running statistics from ITS OWN .training flag (frozen BNs inside a training
    # block keep and use their running statistics); the fused kernel handles the four branches together
    batch_mode = [bn.training or bn.running_mean is None for bn in bns]
    if any(batch_mode) and not all(batch_mode):
        raise NotImplementedError("fused ReparamConv: the four branch BatchNorms must all be in training mode or all in "
                                  "eval mode (mixed frozen / live BatchNorms are not supported)")
    if batch_mode[0]:
        momentum = bns[0].momentum
        if momentum is None or any(bn.momentum != momentum for bn in bns):
            raise NotImplementedError("cumulative-average BatchNorm (momentum=None) is not supported by the fused path")
        update = tracked and all(bn.training for bn in bns)
        _check_running_buffers(bns if update else ())
        running = ([bn.running_mean for bn in bns] if update else None,
                   [bn.running_var for bn in bns] if update else None,
                   [bn.num_batches_tracked for bn in bns] if update else None)
        args = [x1] + [b.conv.weight for b in brs]
        for bn in bns:
            args += [bn.weight, bn.bias]
        return _FusedDwTrain.apply(*args, running, eps, momentum)
    if torch.is_grad_enabled() and (x1.requires_grad or any(p.requires_grad for b in brs for p in b.parameters())):
        raise NotImplementedError("gradients through the eval-mode fused ReparamConv are not implemented; "
                                  "wrap inference in torch.no_grad() (as utils/train_eval_utils.evaluate does)")
    params = L.dw_params([_f32(b.conv.weight) for b in brs], [_f32(bn.weight) for bn in bns],
                         [_f32(bn.bias) for bn in bns], [_f32(bn.running_mean) for bn in bns],
                         [_f32(bn.running_var) for bn in bns])
    return _eval_call(x1, params, None, eps)


def fused_dw_deploy(mod, x1):
    """z, pool for a ReparamConv after switch_to_deploy(): one fused 5x5 depthwise conv + bias, then GELU."""
    fc = mod.fuse_conv
    E = fc.weight.shape[0]
    if tuple(fc.weight.shape) != (E, 1, 5, 5) or fc.stride != (1, 1) or fc.groups != E:
        raise NotImplementedError("deploy-mode fused ReparamConv expects a depthwise 5x5, stride-1 fuse_conv")
    if torch.is_grad_enabled() and (x1.requires_grad or fc.weight.requires_grad):
        raise NotImplementedError("gradients through the deploy-mode fused ReparamConv are not implemented")
    w = _f32(fc.weight)
    params = L.dw_params([w, None, None, None])
    bias = None if fc.bias is None else _f32(fc.bias)
    return _eval_call(x1, params, bias, 0.0)


def _is_1x1(conv) -> bool:
    return (isinstance(conv, torch.nn.Conv2d) and conv.kernel_size == (1, 1) and conv.stride == (1, 1)
            and conv.padding == (0, 0) and conv.groups == 1 and conv.dilation == (1, 1))


def reparam_forward(self, x):
    """Replacement for ReparamConv.forward (/root/reference/core/modules.py:586-600).

    Layout: with 16-bit compute (autocast) the block takes and returns CHANNELS-LAST tensors — the layout cuDNN's 16-bit
    3x3 convolutions and the neighbourhood-attention blocks on either side work in — while the expanded tensor inside the
    block lives as NCHW planes (what the BatchNorm / depthwise kernels want).  Both layout changes ride on the 1x1 GEMMs'
    operand order, so no transposed copy of any activation is made.  An NCHW input is converted once at entry."""
    if x.is_cuda and x.dim() == 4 and _compute_dtype(x) in (torch.bfloat16, torch.float16) and not is_channels_last(x):
        x = to_channels_last(x, _compute_dtype(x))
    ec = self.expand_conv
    if len(ec) == 3 and _is_1x1(ec[0]) and type(ec[1]) is torch.nn.BatchNorm2d and type(ec[2]) is torch.nn.Hardswish:
        # pixel GEMM (BatchNorm sum / sum^2 from its epilogue in training) + fused BatchNorm + Hardswish
        batch_stats = ec[1].training or ec[1].running_mean is None
        y = expand_1x1(ec[0], x, want_stats=batch_stats)
        y, part = y if batch_stats else (y, None)
        x1 = bn_act(ec[1], y, "hardswish", stats=part)
    else:
        x1 = conv_bn_act(ec, x)
    if self.deploy:
        z, pool = fused_dw_deploy(self, x1)
    else:
        z, pool = fused_dw_bn_gelu(self, x1)
    se = self.se
    B, E = pool.shape
    gate = se_gate(se, pool) if pool.is_cuda else se.scale_activation(se.fc2(se.activation(se.fc1(pool.to(z.dtype).view(B, E, 1, 1)))))
    if len(self.pointwise_conv) == 1 and len(self.shortcut) == 1 and _is_1x1(self.pointwise_conv[0]) \
            and _is_1x1(self.shortcut[0]):
        # pointwise(gate * z) + shortcut(x): two accumulating plane-wise GEMMs, the SE multiply rides in the weights
        return pointwise_shortcut(self.pointwise_conv[0], self.shortcut[0], z, gate, x)
    return self.pointwise_conv(gate.reshape(B, E, 1, 1).to(z.dtype) * z) + self.shortcut(x)


def patch_reparam_conv(cls):
    """Swap `forward` on a ReparamConv class object; returns the original forward."""
    original = cls.forward
    cls.forward = reparam_forward
    cls._lmnet_b200_original_forward = original
    return original
