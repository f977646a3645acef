"""Dense 3x3 convolutions of LM-Net on the channels-last implicit-GEMM kernels of ``csrc/conv3x3.cu`` (widening step f3
of SURVEY.md §8): down1-4 / up1-4 (/root/reference/core/LM_Net.py:14-39, 58-74), the M2Skip / M3Skip convolutions
(/root/reference/core/modules.py:83-143) and the OverlapPatchEmbed of the neighbourhood transformers
(/root/reference/core/modules.py:30-39).

``Conv3x3`` is an ``nn.Conv2d`` subclass with the same constructor, parameters and ``state_dict`` keys; only ``forward``
changes.  ``convert_conv3x3(model)`` re-classes the eligible ``nn.Conv2d`` instances of an existing model in place (what
``lmnet_b200.patch`` does to the unmodified reference model).  Shapes the kernels do not cover — fp32 storage, wide
low-resolution layers whose weights do not fit in shared memory, CPU tensors — run the stock cuDNN path.

    forward          csrc/conv3x3.cu  conv3x3_fwd      (12-channel tensors need no padded copy)
    input gradient   stride 1: the same kernel on dy with flipped / transposed weights; stride 2: cuDNN
    weight gradient  csrc/conv3x3.cu  conv3x3_wgrad    (+ bias gradient in the same pass, deterministic)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.amp import custom_bwd, custom_fwd

from . import _lib as L
from .conv1x1 import _compute_dtype


def _dims(B, H, W, Cin, Cout, stride):
    return L.Conv3x3Dims(B, H, W, Cin, Cout, stride)


def fwd_supported(B, H, W, Cin, Cout, stride, dtype) -> bool:
    if dtype not in (torch.bfloat16, torch.float16):
        return False
    return bool(L.lib().lmnet_conv3x3_fwd_supported(L.byref(_dims(B, H, W, Cin, Cout, stride)), L._DTYPES[dtype]))


def wgrad_supported(B, H, W, Cin, Cout, stride, dtype) -> bool:
    if dtype not in (torch.bfloat16, torch.float16):
        return False
    return bool(L.lib().lmnet_conv3x3_wgrad_supported(L.byref(_dims(B, H, W, Cin, Cout, stride)), L._DTYPES[dtype]))


def _pixel_pitch(t):
    """Pixel pitch (elements) of a logical [B,C,H,W] tensor that is dense channels-last (-> C) or a channel slice of a wider
    channels-last tensor (-> the wider tensor's channel count, e.g. the gradient slices torch.cat's backward hands out),
    such that the kernels' 16- / 8-byte vectors stay aligned; None for any other layout."""
    B, C, H, W = t.shape
    sb, sc, sh, sw = t.stride()
    if sc != 1 or sw < C or sh != W * sw or (sb != H * W * sw and B > 1):
        return None
    vec = 16 if C % 8 == 0 else 8
    if (sw * t.element_size()) % vec != 0 or t.data_ptr() % vec != 0:
        return None
    return sw


def _launch_fwd(x, w, transposed, bias, Cout, stride, pitch=0):
    """x: logical [B,Cin,H,W] in channels-last memory with pixel pitch `pitch` (0 = dense); w: the layer's fp32 weight, read
    in place by the kernel (transposed: the stride-1 input gradient uses it flipped and transposed); returns logical
    [B,Cout,Ho,Wo] channels-last."""
    B, Cin, H, W = x.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    y = torch.empty((B, Cout, Ho, Wo), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    dims = _dims(B, H, W, Cin, Cout, stride)
    rc = L.lib().lmnet_conv3x3_fwd_strided(L.ptr(x), int(pitch), L.ptr(w), int(transposed), L.ptr(bias), L.ptr(y), L.byref(dims),
                                           L.dtype_code(x), L.stream_ptr())
    L.check(rc, "conv3x3_fwd")
    return y


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _w32(w):
    w = w.detach()
    return w if (w.dtype == torch.float32 and w.is_contiguous()) else w.float().contiguous()


class _Conv3x3(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, x, w, bias, stride):
        Cout, Cin = w.shape[0], w.shape[1]
        b32 = None if bias is None else bias.detach().float().contiguous()
        y = _launch_fwd(x, _w32(w), False, b32, Cout, stride)
        ctx.save_for_backward(x, w)
        ctx.stride = stride
        ctx.bias_dtype = None if bias is None else bias.dtype
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        s = ctx.stride
        B, Cin, H, W = x.shape
        Cout = w.shape[0]
        dt = x.dtype
        dy = dy.to(dt)
        # the gradient of a convolution whose output was concatenated arrives as a channel slice of the concatenation's
        # channels-last gradient: the kernels read it in place (pixel pitch); any other layout is densified once
        pitch = _pixel_pitch(dy) if dy.dim() == 4 else None
        if pitch is None:
            dy = _cl(dy)
            pitch = 0
        elif pitch == dy.shape[1]:
            pitch = 0
        has_bias = ctx.bias_dtype is not None
        dx = dw = db = None
        dy_dense = None

        def dense():
            nonlocal dy_dense
            if dy_dense is None:
                dy_dense = dy if pitch == 0 else _cl(dy)
            return dy_dense

        def aten_bwd(mask):
            return torch.ops.aten.convolution_backward(dense(), x, w.to(dt), [Cout] if has_bias else None, [s, s], [1, 1], [1, 1],
                                                       False, [0, 0], 1, mask)

        if ctx.needs_input_grad[0]:
            Ho, Wo = dy.shape[2], dy.shape[3]
            if s == 1 and fwd_supported(B, Ho, Wo, Cout, Cin, 1, dt):
                dx = _launch_fwd(dy, _w32(w), True, None, Cin, 1, pitch)
            else:
                dx = aten_bwd([True, False, False])[0]
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            if wgrad_supported(B, H, W, Cin, Cout, s, dt):
                dims = _dims(B, H, W, Cin, Cout, s)
                dW = torch.empty(Cout, Cin, 3, 3, dtype=torch.float32, device=x.device)
                dbias = torch.empty(Cout, dtype=torch.float32, device=x.device) if has_bias else None
                n = L.lib().lmnet_conv3x3_wgrad_workspace_bytes(L.byref(dims))
                ws = torch.empty(max(int(n), 16), dtype=torch.uint8, device=x.device)
                rc = L.lib().lmnet_conv3x3_wgrad_strided(L.ptr(x), L.ptr(dy), int(pitch), L.ptr(dW), L.ptr(dbias), L.ptr(ws),
                                                         ws.numel(), L.byref(dims), L.dtype_code(x), L.stream_ptr())
                L.check(rc, "conv3x3_wgrad")
                dw = dW.to(w.dtype)
                db = dbias.to(ctx.bias_dtype) if has_bias else None
            else:
                _, dw, db = aten_bwd([False, True, has_bias])
                dw = dw.to(w.dtype)
                db = db.to(ctx.bias_dtype) if has_bias else None
        return dx, dw, db, None


def _eligible(conv: torch.nn.Conv2d) -> bool:
    return (conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.stride in ((1, 1), (2, 2)) and conv.groups == 1
            and conv.dilation == (1, 1) and conv.padding_mode == "zeros")


def conv3x3(conv: torch.nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """conv(x) for a dense 3x3 / padding-1 convolution; channels-last in and out on the sm_100a kernel where it applies."""
    if not x.is_cuda or x.dim() != 4 or not _eligible(conv):
        return F.conv2d(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
    dt = _compute_dtype(x)
    B, Cin, H, W = x.shape
    s = conv.stride[0]
    if dt not in (torch.bfloat16, torch.float16):          # fp32 storage: the stock path, layout untouched
        return F.conv2d(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
    if not fwd_supported(B, H, W, Cin, conv.out_channels, s, dt):
        # cuDNN's 16-bit convolutions compute in NHWC: hand them a channels-last tensor (no-op when it already is)
        return F.conv2d(_cl(x), conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
    return _Conv3x3.apply(_cl(x.to(dt)), conv.weight, conv.bias, s)


class Conv3x3(torch.nn.Conv2d):
    """nn.Conv2d whose forward runs on csrc/conv3x3.cu when the layer is a dense 3x3 / padding-1 / stride-1|2 convolution
    on a 16-bit CUDA tensor, and on cuDNN otherwise.  Same parameters and state_dict keys as nn.Conv2d."""

    def forward(self, x):
        return conv3x3(self, x)


def convert_conv3x3(model: torch.nn.Module) -> torch.nn.Module:
    """Re-class every eligible plain nn.Conv2d of `model` to Conv3x3, in place (parameters are untouched)."""
    for m in model.modules():
        if type(m) is torch.nn.Conv2d and _eligible(m):
            m.__class__ = Conv3x3
    return model
