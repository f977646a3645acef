"""TEST INFRASTRUCTURE — LM-Net on the CPU: the reference op sequence with natten's CPU ops
restated by the C oracle.  This is the `--impl reference` / cpu_baseline arm of bench.py and the
model-level checker of the tests (reference modules in torch CPU, NA through the
CPU oracle restating natten's naive CPU ops, fp32).

The network definition (layer graph, parameter names) is taken from ``lmnet_b200.model`` — which
tests/test_reference_parity.py pins against the unmodified /root/reference/core/LM_Net.py — and the
two hot-path units are swapped for their oracle forms:
  * every ``ReparamConv``            -> oracle.reparam_ref.reparam_forward_ref (stock torch ops)
  * every ``NeighborhoodAttention2D``-> oracle.na2d_ref.OracleNeighborhoodAttention2D (C oracle)
and every module whose forward calls a widened lmnet_b200 operator (LayerNorm, BN + activation, up-sampling:
SURVEY.md §8 f2/f3) gets the reference's stock-torch forward back:
  * ``NeighborhoodTransformer``      -> /root/reference/core/modules.py:514-521
  * ``M3Skip`` / ``M2Skip``          -> /root/reference/core/modules.py:101-107, 137-143
  * ``Upsample2x``                   -> nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
Nothing here touches the CUDA extension (tests/test_oracle.py runs it with the library calls disabled).
"""
import types

import torch

from lmnet_b200 import model as M
from natten import NeighborhoodAttention2D

from .na2d_ref import OracleNeighborhoodAttention2D
from .reparam_ref import reparam_forward_ref


def _transformer_forward_ref(self, x):
    emb = self.patchembedding(x)
    att = self.att1(self.norm1(emb)) + emb
    y = self.mlp(self.norm2(att)) + att
    return y.permute(0, 3, 1, 2).contiguous()


def _m3skip_forward_ref(self, xl, xm, xs):
    return self.fuse_conv(torch.cat([self.convl(xl), self.convm(xm), self.convs(xs)], dim=1))


def _m2skip_forward_ref(self, xl, xs):
    return self.fuse_conv(torch.cat([self.convl(xl), self.convs(xs)], dim=1))


def to_oracle(net: torch.nn.Module, swap_reparam: bool = True) -> torch.nn.Module:
    """In-place: route the hot-path units of `net` (an lmnet_b200.model.LM_Net or any module tree that
    contains its ReparamConv / NeighborhoodAttention2D) through the CPU oracle."""
    for mod in list(net.modules()):
        if swap_reparam and isinstance(mod, M.ReparamConv):
            mod.forward = types.MethodType(reparam_forward_ref, mod)
        elif isinstance(mod, M.NeighborhoodTransformer):
            mod.forward = types.MethodType(_transformer_forward_ref, mod)
        elif isinstance(mod, M.M3Skip):
            mod.forward = types.MethodType(_m3skip_forward_ref, mod)
        elif isinstance(mod, M.M2Skip):
            mod.forward = types.MethodType(_m2skip_forward_ref, mod)
        for name, child in list(mod.named_children()):
            if isinstance(child, M.Upsample2x):
                setattr(mod, name, torch.nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True))
            if isinstance(child, NeighborhoodAttention2D):
                repl = OracleNeighborhoodAttention2D(child.dim, child.num_heads, child.kernel_size, child.dilation,
                                                     bias=child.rpb is not None,
                                                     qkv_bias=child.qkv.bias is not None, qk_scale=child.scale)
                repl.load_state_dict(child.state_dict())
                repl.to(next(child.parameters()).dtype)
                setattr(mod, name, repl)
    return net


def build_cpu_reference(channel=3, n_classes=2, seed=None, dtype=torch.float32):
    if seed is not None:
        torch.manual_seed(seed)
    net = M.LM_Net(channel, n_classes)
    return to_oracle(net).to(dtype)
