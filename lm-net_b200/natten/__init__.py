"""Drop-in ``natten`` package backed by the B200 (sm_100a) kernels of ``lmnet_b200``.

LM-Net does ``from natten import NeighborhoodAttention2D`` (/root/reference/core/modules.py:18) and
builds ``NeighborhoodAttention2D(dim=C, num_heads=12, kernel_size=3)`` (:509).  Putting
``lm-net_b200/`` on ``sys.path`` makes that import resolve here; nothing in the reference changes.

API generation mirrored: natten 0.14-0.17 (``dim=`` keyword, relative positional bias ``rpb``,
unfused ``na2d_qk`` / ``na2d_av`` plus the 0.14 spellings ``natten2dqkrpb`` / ``natten2dav``, and the
fused ``na2d``).  This is an independent implementation, not a port of NATTEN's kernels.
"""
from .module import NeighborhoodAttention2D
from . import functional
from .functional import na2d, na2d_av, na2d_qk

__version__ = "0.17.0+lmnet.b200"


def has_cuda() -> bool:
    import torch

    return torch.cuda.is_available()


__all__ = ["NeighborhoodAttention2D", "functional", "na2d", "na2d_qk", "na2d_av", "has_cuda"]
