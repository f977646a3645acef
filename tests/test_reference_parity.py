"""Runs only where the unmodified reference is mounted (/root/reference): the drop-in recipe of
INTEGRATION.md against the real core/LM_Net.py and core/modules.py."""
import copy

import pytest
import torch

from _helpers import fill_deterministic, import_reference, reference_available, rel_err

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present on this machine")


def test_reference_imports_our_natten_and_keys_match():
    lm, mods = import_reference()
    import natten
    from lmnet_b200.model import LM_Net

    assert mods.NeighborhoodAttention2D is natten.NeighborhoodAttention2D      # the drop-in resolved
    ref, ours = lm.LM_Net(3, 2), LM_Net(3, 2)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    ours.load_state_dict(a, strict=True)
    ref.load_state_dict(b, strict=True)


def test_oracle_restatement_of_reparamconv_equals_reference_class():
    from oracle.reparam_ref import reparam_forward_ref

    _, mods = import_reference()
    blk = mods.ReparamConv(5, 10, 7, 5, 3).double()
    fill_deterministic(blk, seed=4)
    ref = copy.deepcopy(blk)
    x = torch.randn(2, 5, 13, 9, dtype=torch.float64)
    for mode in ("train", "eval"):
        getattr(blk, mode)(), getattr(ref, mode)()
        assert rel_err(reparam_forward_ref(blk, x), ref(x)) < 1e-13
    for (n, a), (_, b) in zip(blk.named_buffers(), ref.named_buffers()):
        assert torch.equal(a, b), n


def test_patched_reference_model_equals_reference(cpu_backend):
    """patch_reference_modules + natten drop-in on the REAL reference model == the reference's own forward
    (natten's arithmetic answered by the CPU oracle on both sides)."""
    from lmnet_b200.patch import convert_upsample, patch_reference_modules, unpatch_reference_modules
    from oracle.lmnet_ref import to_oracle

    lm, mods = import_reference()
    ref = lm.LM_Net(3, 2).double()
    fill_deterministic(ref, seed=5)
    ours = copy.deepcopy(ref)                     # same classes, same weights
    ref = to_oracle(ref, swap_reparam=False).eval()
    x = torch.randn(1, 3, 32, 32, dtype=torch.float64)
    with torch.no_grad():
        want = ref(x)
    originals = patch_reference_modules(mods)     # ReparamConv, NeighborhoodTransformer, M2Skip, M3Skip
    try:
        ours = convert_upsample(ours)
        assert sum(type(m).__name__ == "Upsample2x" for m in ours.modules()) == 7
        with torch.no_grad():
            got = ours.eval()(x)                  # reference classes, our forwards + our natten module
    finally:
        unpatch_reference_modules(mods, originals)
    assert rel_err(got, want) < 1e-7
    assert torch.equal(got.argmax(1), want.argmax(1))
