"""TEST INFRASTRUCTURE — LM-Net on the CPU: the reference op sequence with natten's CPU ops
restated by the C oracle.  This is the `--impl reference` / cpu_baseline arm of bench.py and the
model-level checker of the tests (reference modules in torch CPU, NA through the
CPU oracle restating natten's naive CPU ops, fp32).

The network definition (layer graph, parameter names) is taken from ``lmnet_b200.model`` — which
tests/test_reference_parity.py pins against the unmodified /root/reference/core/LM_Net.py — and the
two hot-path units are swapped for their oracle forms:
  * every ``ReparamConv``            -> oracle.reparam_ref.reparam_forward_ref (stock torch ops)
  * every ``NeighborhoodAttention2D``-> oracle.na2d_ref.OracleNeighborhoodAttention2D (C oracle)
Nothing here touches the CUDA extension.
"""
import types

import torch

from lmnet_b200 import model as M
from natten import NeighborhoodAttention2D

from .na2d_ref import OracleNeighborhoodAttention2D
from .reparam_ref import reparam_forward_ref


def to_oracle(net: torch.nn.Module, swap_reparam: bool = True) -> torch.nn.Module:
    """In-place: route the hot-path units of `net` (an lmnet_b200.model.LM_Net or any module tree that
    contains its ReparamConv / NeighborhoodAttention2D) through the CPU oracle."""
    for mod in list(net.modules()):
        if swap_reparam and isinstance(mod, M.ReparamConv):
            mod.forward = types.MethodType(reparam_forward_ref, mod)
        for name, child in list(mod.named_children()):
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
