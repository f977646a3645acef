"""lmnet_b200 — B200-native (sm_100a) hot path of LM-Net behind the reference's own Python surface.

* ``lmnet_b200.na_ops``   autograd operators over the C ABI (fused / unfused neighbourhood attention)
* ``lmnet_b200.reparam``  fused depthwise multi-branch conv + BN + GELU for ``ReparamConv``
* ``lmnet_b200.model``    LM-Net assembled on those operators (state_dict-compatible with the reference)
* ``natten`` (sibling package) the drop-in the reference imports at core/modules.py:18
"""
__version__ = "0.1.0"
