"""``NeighborhoodAttention2D`` — the module LM-Net instantiates at /root/reference/core/modules.py:509.

Same constructor spellings, parameter names (``qkv``, ``rpb``, ``proj``) and I/O layout
([B, H, W, C] in and out) as natten 0.14-0.17, so reference checkpoints load unchanged.  The inner
q*scale -> QK+rpb -> softmax -> AV sequence runs as ONE fused CUDA kernel that reads q, k and v
straight out of the packed projection and never materialises the attention map; its backward
recomputes.  With attention dropout > 0 in training the unfused operators are used instead (the
dropout mask has to act on a materialised map).
"""
import torch
import torch.nn.functional as F
from torch import nn

from lmnet_b200 import na_ops
from lmnet_b200.linear import linear


class NeighborhoodAttention2D(nn.Module):
    def __init__(self, dim, num_heads, kernel_size, dilation=1, bias=True, qkv_bias=True, qk_scale=None,
                 attn_drop=0.0, proj_drop=0.0, rel_pos_bias=None):
        super().__init__()
        if dim % num_heads != 0:
            raise ValueError(f"dim ({dim}) must be divisible by num_heads ({num_heads})")
        if not isinstance(kernel_size, int) or kernel_size < 3 or kernel_size % 2 != 1:
            raise ValueError(f"kernel_size must be an odd integer > 1, got {kernel_size}")
        dilation = 1 if dilation is None else dilation
        if not isinstance(dilation, int) or dilation < 1:
            raise ValueError(f"dilation must be an integer >= 1, got {dilation}")
        # 0.14 calls the switch `bias`, >= 0.15 `rel_pos_bias`; LM-Net passes neither -> rpb on
        use_rpb = bias if rel_pos_bias is None else rel_pos_bias
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.scale = qk_scale or self.head_dim ** -0.5
        self.kernel_size, self.dilation = kernel_size, dilation
        self.window_size = kernel_size * dilation
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        if use_rpb:
            self.rpb = nn.Parameter(torch.zeros(num_heads, 2 * kernel_size - 1, 2 * kernel_size - 1))
            nn.init.trunc_normal_(self.rpb, std=0.02, mean=0.0, a=-2.0, b=2.0)
        else:
            self.register_parameter("rpb", None)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        if x.dim() != 4 or x.shape[-1] != self.dim:
            raise ValueError(f"expected [B, H, W, {self.dim}] input, got {tuple(x.shape)}")
        B, Hp, Wp, C = x.shape
        H, W = Hp, Wp
        pad_b = pad_r = 0
        if H < self.window_size or W < self.window_size:  # natten 0.14 zero-pads small maps, then crops
            pad_r = max(0, self.window_size - W)
            pad_b = max(0, self.window_size - H)
            x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
            _, H, W, _ = x.shape
        qkv = linear(self.qkv, x).view(B, H, W, 3, self.num_heads, self.head_dim)
        if self.training and self.attn_drop.p > 0.0:
            q, k, v = qkv.permute(3, 0, 4, 1, 2, 5).unbind(0)
            attn = na_ops.na2d_qk(q * self.scale, k, self.kernel_size, self.dilation, rel_pos_bias=self.rpb)
            attn = self.attn_drop(attn.softmax(dim=-1))
            o = na_ops.na2d_av(attn, v, self.kernel_size, self.dilation).permute(0, 2, 3, 1, 4)
        else:
            o = na_ops.na2d_qkvpacked(qkv, self.kernel_size, self.dilation, rel_pos_bias=self.rpb, scale=self.scale)
        o = o.reshape(B, H, W, C)
        if pad_r or pad_b:
            o = o[:, :Hp, :Wp, :]
        return self.proj_drop(linear(self.proj, o))

    def extra_repr(self):
        return (f"head_dim={self.head_dim}, num_heads={self.num_heads}, kernel_size={self.kernel_size}, "
                f"dilation={self.dilation}, rel_pos_bias={self.rpb is not None}")
