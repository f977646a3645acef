"""``natten.functional`` surface on the lmnet_b200 CUDA kernels."""
from lmnet_b200.na_ops import na2d, na2d_av, na2d_qk, na2d_qkvpacked


def natten2dqkrpb(query, key, rpb, kernel_size, dilation=1):
    """natten 0.14 spelling (also what transformers' DiNAT calls)."""
    return na2d_qk(query, key, kernel_size, dilation, rel_pos_bias=rpb)


def natten2dav(attn, value, kernel_size, dilation=1):
    """natten 0.14 spelling."""
    return na2d_av(attn, value, kernel_size, dilation)


__all__ = ["na2d", "na2d_qk", "na2d_av", "na2d_qkvpacked", "natten2dqkrpb", "natten2dav"]
