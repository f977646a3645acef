"""CPU, fp64: the identities the composite depthwise backward relies on (csrc/reparam_dw_tma2.cuh) against
torch autograd of the reference op sequence — four depthwise conv + training-mode BatchNorm branches, summed
(/root/reference/core/modules.py:592-595):

    dx       = K5[flip(sum_br c1_br w_br)](du) - K9[C](x) - sum_s Wc0[s] + frame(q)        (frame: 2-pixel border only)
    dw_br[t] = c1_br P[t] - c2_br Q_br[t] - c0_br S[t]

The restatement lives in tools/debug/dx_composite_proto.py (it asserts to 1e-9 relative); this test runs it on an interior +
border case, a plane whose every pixel is on the frame, and a single-row plane."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shape", [(2, 2, 9, 11), (1, 2, 4, 8), (1, 1, 1, 8)], ids=lambda s: "x".join(map(str, s)))
def test_composite_stencil_identities_match_autograd(shape):
    spec = importlib.util.spec_from_file_location("dx_composite_proto", os.path.join(ROOT, "tools", "debug", "dx_composite_proto.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.run(*shape)          # raises AssertionError on any mismatch (dx incl. the frame, and all four weight gradients)
