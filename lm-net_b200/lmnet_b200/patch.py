"""One call that puts the UNMODIFIED reference model on the sm_100a kernels.

    import sys; sys.path.insert(0, ".../lm-net_b200")       # `import natten` now resolves to the drop-in
    import core.modules, core.LM_Net
    from lmnet_b200.patch import patch_reference_modules, convert_model
    patch_reference_modules(core.modules)                    # class-level forward swaps, nothing else changes
    model = convert_model(core.LM_Net.LM_Net(3, 2).cuda())   # re-classes nn.Upsample / 3x3 nn.Conv2d instances in place

What is swapped (constructors, sub-module names and therefore checkpoints stay untouched):
  ReparamConv.forward            core/modules.py:586-600   -> lmnet_b200.reparam.reparam_forward
  NeighborhoodTransformer.forward core/modules.py:514-521  -> fused LayerNorm + channels-last patch embedding
  M3Skip.forward / M2Skip.forward core/modules.py:101-107, 138-143 -> fused BatchNorm+GELU, channels-last 3x3 convs
  PyramidPool.forward            core/modules.py:481-498   -> integer-factor channels-last average pooling kernel
  nn.Upsample(scale_factor=2, bilinear, align_corners=True) instances -> lmnet_b200.upsample.Upsample2x
  nn.Conv2d(.., 3, stride 1|2, padding 1) instances -> lmnet_b200.conv3x3.Conv3x3 (same parameters / state_dict keys)
NeighborhoodAttention2D itself comes from the drop-in `natten` package.
"""
from __future__ import annotations

import torch

from .bnact import conv_bn_act
from .conv3x3 import convert_conv3x3
from .layernorm import layer_norm
from .linear import linear
from .pool import adaptive_avg_pool
from .reparam import patch_reparam_conv
from .upsample import Upsample2x


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last) if t.is_cuda else t


def _mlp_forward(mlp, x):
    """Mlp.forward (core/modules.py:50-56) with fc1 / fc2 on the pixel-GEMM kernels; GELU and Dropout as in the module."""
    if not (type(mlp.fc1) is torch.nn.Linear and type(mlp.fc2) is torch.nn.Linear) or not x.is_cuda:
        return mlp(x)
    return mlp.dropout(linear(mlp.fc2, mlp.dropout(mlp.act_fn(linear(mlp.fc1, x)))))


def _natt_forward(self, x):
    """NeighborhoodTransformer.forward (core/modules.py:514-521): 3x3 conv embed -> LN -> NA -> +res -> LN -> MLP -> +res."""
    emb = self.patchembedding(_cl(x))              # NHWC conv output: its [B,H,W,C] permute is a free view
    att = self.att1(layer_norm(self.norm1, emb)) + emb
    y = _mlp_forward(self.mlp, layer_norm(self.norm2, att)) + att
    # logical NCHW on channels-last memory: a free view (every consumer is channels-last as well)
    return y.permute(0, 3, 1, 2) if y.is_cuda else y.permute(0, 3, 1, 2).contiguous()


def _m3skip_forward(self, xl, xm, xs):
    return conv_bn_act(self.fuse_conv, torch.cat([self.convl(_cl(xl)), self.convm(_cl(xm)), self.convs(xs)], dim=1))


def _m2skip_forward(self, xl, xs):
    xs = self.convs(_cl(xs)) if self.model_type == "bottom" else self.convs(xs)
    return conv_bn_act(self.fuse_conv, torch.cat([self.convl(_cl(xl)), xs], dim=1))


def _pyramidpool_forward(self, x1, x2, x3, x4, x5):
    """PyramidPool.forward (core/modules.py:481-498): pool the four encoder outputs to x5's size and stack."""
    size = x5.shape[-2:]
    return torch.cat([adaptive_avg_pool(t, size) for t in (x1, x2, x3, x4)] + [x5], dim=1)


def patch_reference_modules(mods) -> dict:
    """`mods` is the reference's imported `core.modules`.  Returns the original forwards (to undo)."""
    originals = {"ReparamConv": patch_reparam_conv(mods.ReparamConv)}
    for name, fwd in (("NeighborhoodTransformer", _natt_forward), ("M3Skip", _m3skip_forward), ("M2Skip", _m2skip_forward),
                      ("PyramidPool", _pyramidpool_forward)):
        cls = getattr(mods, name, None)
        if cls is None:
            continue
        originals[name] = cls.forward
        cls.forward = fwd
    return originals


def unpatch_reference_modules(mods, originals: dict) -> None:
    for name, fwd in originals.items():
        getattr(mods, name).forward = fwd


def convert_upsample(model: torch.nn.Module) -> torch.nn.Module:
    """Replace every nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) by Upsample2x (in place)."""
    for parent in model.modules():
        for name, child in list(parent.named_children()):
            if isinstance(child, torch.nn.Upsample) and child.mode == "bilinear" and child.align_corners \
                    and child.scale_factor in (2, 2.0, (2, 2), (2.0, 2.0)):
                setattr(parent, name, Upsample2x())
    return model


def convert_model(model: torch.nn.Module) -> torch.nn.Module:
    """Everything instance-level in one call: bilinear x2 up-sampling and the dense 3x3 convolutions (in place)."""
    return convert_conv3x3(convert_upsample(model))
