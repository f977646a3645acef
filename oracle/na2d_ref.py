"""TEST INFRASTRUCTURE — Python face of the CPU oracle for neighbourhood attention.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product never does.

Three independent statements of the same arithmetic live here so that they can be
checked against each other (the real ``natten`` is absent offline — PARITY UNPINNED,
see DESIGN.md §3):

1. ``COracle``        — ctypes binding of ``oracle/na2d_oracle.c`` (sub-sequence /
                        clamp formulation, fp32 and fp64, OpenMP).
2. ``neighbour_table``— numpy index tables built from the *closed-form* window
                        rules quoted in SURVEY.md §8(c3) (the form natten's
                        kernels use), and ``na2d_*_gather`` torch ops built on
                        them (differentiable by autograd, so they also check
                        the C oracle's analytic backward).
3. the known-answer tests in ``tests/test_oracle.py`` (K == L ⇒ global attention
   with a Swin-style relative bias; interior pixels ⇒ ``F.unfold`` sliding window;
   dilation ⇒ independent sub-grids).
4. an implementation that is not ours: PyTorch's FlexAttention with the neighbourhood mask
   published for NATTEN in PyTorch's attention-gym examples reproduces the fused forward to
   1e-10 in fp64 (``tests/test_oracle.py``).

Call sites restated: /root/reference/core/modules.py:509 (construction, kernel 3,
12 heads) and :517 (forward).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from functools import lru_cache

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libna2d_oracle.so")


def build_oracle(force: bool = False) -> str:
    """Compile oracle/na2d_oracle.c into oracle/_ref/ (gcc + OpenMP)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


# --------------------------------------------------------------------------- #
# 1. ctypes binding of the C oracle
# --------------------------------------------------------------------------- #
class COracle:
    """fp32 / fp64 C oracle.  All tensors are contiguous CPU tensors."""

    def __init__(self):
        self.lib = ctypes.CDLL(build_oracle())
        assert self.lib.na2d_oracle_abi_version() == 1

    @staticmethod
    def _sfx(t: torch.Tensor) -> str:
        if t.dtype == torch.float32:
            return "f32"
        if t.dtype == torch.float64:
            return "f64"
        raise TypeError(f"oracle supports fp32/fp64, got {t.dtype}")

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(0 if t is None else t.data_ptr())

    @staticmethod
    def _check(*ts):
        for t in ts:
            if t is not None:
                assert t.device.type == "cpu" and t.is_contiguous(), "oracle needs contiguous CPU tensors"

    def _real(self, t, x):
        return ctypes.c_float(x) if t.dtype == torch.float32 else ctypes.c_double(x)

    # unfused ops, layout [B, heads, H, W, D] / [B, heads, H, W, K*K]
    def qk_fwd(self, q, k, rpb, K, d=1):
        self._check(q, k, rpb)
        B, Hd, H, W, D = q.shape
        attn = torch.empty(B, Hd, H, W, K * K, dtype=q.dtype)
        getattr(self.lib, "na2d_qk_fwd_" + self._sfx(q))(
            self._p(q), self._p(k), self._p(rpb), self._p(attn), B, Hd, H, W, D, K, d)
        return attn

    def qk_bwd(self, q, k, dattn, K, d=1, need_rpb=True):
        self._check(q, k, dattn)
        B, Hd, H, W, D = q.shape
        dq, dk = torch.empty_like(q), torch.empty_like(k)
        drpb = torch.empty(Hd, 2 * K - 1, 2 * K - 1, dtype=q.dtype) if need_rpb else None
        getattr(self.lib, "na2d_qk_bwd_" + self._sfx(q))(
            self._p(q), self._p(k), self._p(dattn), self._p(dq), self._p(dk), self._p(drpb),
            B, Hd, H, W, D, K, d)
        return dq, dk, drpb

    def av_fwd(self, attn, v, K, d=1):
        self._check(attn, v)
        B, Hd, H, W, D = v.shape
        out = torch.empty_like(v)
        getattr(self.lib, "na2d_av_fwd_" + self._sfx(v))(
            self._p(attn), self._p(v), self._p(out), B, Hd, H, W, D, K, d)
        return out

    def av_bwd(self, attn, v, dout, K, d=1):
        self._check(attn, v, dout)
        B, Hd, H, W, D = v.shape
        dattn, dv = torch.empty_like(attn), torch.empty_like(v)
        getattr(self.lib, "na2d_av_bwd_" + self._sfx(v))(
            self._p(attn), self._p(v), self._p(dout), self._p(dattn), self._p(dv),
            B, Hd, H, W, D, K, d)
        return dattn, dv

    # fused op, layout [B, H, W, heads, D]
    def fused_fwd(self, q, k, v, rpb, K, d=1, scale=None, want_lse=False):
        self._check(q, k, v, rpb)
        B, H, W, Hd, D = q.shape
        scale = D ** -0.5 if scale is None else scale
        out = torch.empty_like(q)
        lse = torch.empty(B, H, W, Hd, dtype=q.dtype) if want_lse else None
        getattr(self.lib, "na2d_fused_fwd_" + self._sfx(q))(
            self._p(q), self._p(k), self._p(v), self._p(rpb), self._p(out), self._p(lse),
            self._real(q, scale), B, H, W, Hd, D, K, d)
        return (out, lse) if want_lse else out

    def fused_bwd(self, q, k, v, rpb, dout, K, d=1, scale=None):
        self._check(q, k, v, rpb, dout)
        B, H, W, Hd, D = q.shape
        scale = D ** -0.5 if scale is None else scale
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        drpb = torch.empty(Hd, 2 * K - 1, 2 * K - 1, dtype=q.dtype) if rpb is not None else None
        getattr(self.lib, "na2d_fused_bwd_" + self._sfx(q))(
            self._p(q), self._p(k), self._p(v), self._p(rpb), self._p(dout),
            self._p(dq), self._p(dk), self._p(dv), self._p(drpb),
            self._real(q, scale), B, H, W, Hd, D, K, d)
        return dq, dk, dv, drpb


@lru_cache(maxsize=1)
def c_oracle() -> COracle:
    return COracle()


# --------------------------------------------------------------------------- #
# 2. closed-form index tables + torch gather ops
# --------------------------------------------------------------------------- #
def _window_start_closed_form(i: int, L: int, K: int, d: int) -> int:
    """First attended key along one axis — closed form quoted in SURVEY.md §8(c3)."""
    ns = K // 2
    if d <= 1:
        return max(i - ns, 0) + ((L - i - ns - 1) if (i + ns >= L) else 0)
    ni = i - ns * d
    if ni < 0:
        return i % d
    if i + ns * d >= L:
        imodd = i % d
        a = (L // d) * d
        b = L - a
        if imodd < b:
            return L - b + imodd - 2 * ns * d
        return a + imodd - K * d
    return ni


def _pb_start_closed_form(i: int, L: int, K: int, d: int) -> int:
    """rpb index of the first attended key — closed form quoted in SURVEY.md §8(c3)."""
    ns = K // 2
    if d <= 1:
        return ns + ((ns - i) if i < ns else 0) + ((L - i - 1 - ns) if (i + ns >= L) else 0)
    if i - ns * d < 0:
        return K - 1 - (i // d)
    if i + ns * d >= L:
        return (L - i - 1) // d
    return ns


@lru_cache(maxsize=256)
def neighbour_table(L: int, K: int, d: int):
    """(keys[L,K], pbs[L,K]) int64 numpy arrays for one axis."""
    if K % 2 != 1 or K < 3:
        raise ValueError("kernel_size must be odd and > 1")
    if L < K * d:
        raise ValueError(f"axis length {L} < kernel_size*dilation {K * d}")
    keys = np.empty((L, K), dtype=np.int64)
    pbs = np.empty((L, K), dtype=np.int64)
    for i in range(L):
        s = _window_start_closed_form(i, L, K, d)
        p = _pb_start_closed_form(i, L, K, d)
        keys[i] = s + d * np.arange(K)
        pbs[i] = p + np.arange(K)
    return keys, pbs


def _flat_tables(H, W, K, d, device):
    ki, pi = neighbour_table(H, K, d)
    kj, pj = neighbour_table(W, K, d)
    # neighbour n = mi*K + mj
    key_flat = (ki[:, None, :, None] * W + kj[None, :, None, :]).reshape(H, W, K * K)
    R = 2 * K - 1
    pb_flat = (pi[:, None, :, None] * R + pj[None, :, None, :]).reshape(H, W, K * K)
    return (torch.from_numpy(key_flat).to(device), torch.from_numpy(pb_flat).to(device))


def na2d_qk_gather(q, k, rpb, K, d=1):
    """[B,heads,H,W,D] x2 (+ rpb[heads,2K-1,2K-1]) -> attn [B,heads,H,W,K*K]; autograd-differentiable."""
    B, Hd, H, W, D = q.shape
    key_flat, pb_flat = _flat_tables(H, W, K, d, q.device)
    kf = k.reshape(B, Hd, H * W, D)
    kn = kf[:, :, key_flat.reshape(-1)].reshape(B, Hd, H, W, K * K, D)
    attn = torch.einsum("bhijd,bhijnd->bhijn", q, kn)
    if rpb is not None:
        attn = attn + rpb.reshape(Hd, -1)[:, pb_flat.reshape(-1)].reshape(1, Hd, H, W, K * K)
    return attn


def na2d_av_gather(attn, v, K, d=1):
    """attn [B,heads,H,W,K*K], v [B,heads,H,W,D] -> out [B,heads,H,W,D]; autograd-differentiable."""
    B, Hd, H, W, D = v.shape
    key_flat, _ = _flat_tables(H, W, K, d, v.device)
    vf = v.reshape(B, Hd, H * W, D)
    vn = vf[:, :, key_flat.reshape(-1)].reshape(B, Hd, H, W, K * K, D)
    return torch.einsum("bhijn,bhijnd->bhijd", attn, vn)


def na2d_gather(q, k, v, K, d=1, rpb=None, scale=None):
    """Fused-API layout [B,H,W,heads,D]; autograd-differentiable."""
    D = q.shape[-1]
    scale = D ** -0.5 if scale is None else scale
    qh, kh, vh = (t.permute(0, 3, 1, 2, 4) for t in (q, k, v))
    attn = na2d_qk_gather(qh * scale, kh, rpb, K, d).softmax(-1)
    return na2d_av_gather(attn, vh, K, d).permute(0, 2, 3, 1, 4)


# --------------------------------------------------------------------------- #
# torch.autograd wrappers over the C oracle (the "reference natten CPU ops" arm
# of the CPU baseline; API spelling of natten 0.14: natten2dqkrpb / natten2dav)
# --------------------------------------------------------------------------- #
class _QKRPB(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, rpb, K, d):
        q, k = q.contiguous(), k.contiguous()
        rpb_c = rpb.contiguous() if rpb is not None else None
        ctx.save_for_backward(q, k)
        ctx.K, ctx.d, ctx.has_rpb = K, d, rpb is not None
        return c_oracle().qk_fwd(q, k, rpb_c, K, d)

    @staticmethod
    def backward(ctx, dattn):
        q, k = ctx.saved_tensors
        dq, dk, drpb = c_oracle().qk_bwd(q, k, dattn.contiguous(), ctx.K, ctx.d, ctx.has_rpb)
        return dq, dk, drpb, None, None


class _AV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attn, v, K, d):
        attn, v = attn.contiguous(), v.contiguous()
        ctx.save_for_backward(attn, v)
        ctx.K, ctx.d = K, d
        return c_oracle().av_fwd(attn, v, K, d)

    @staticmethod
    def backward(ctx, dout):
        attn, v = ctx.saved_tensors
        dattn, dv = c_oracle().av_bwd(attn, v, dout.contiguous(), ctx.K, ctx.d)
        return dattn, dv, None, None


def natten2dqkrpb(q, k, rpb, kernel_size, dilation=1):
    return _QKRPB.apply(q, k, rpb, kernel_size, dilation)


def natten2dav(attn, v, kernel_size, dilation=1):
    return _AV.apply(attn, v, kernel_size, dilation)


class OracleNeighborhoodAttention2D(torch.nn.Module):
    """CPU restatement of natten 0.14's ``NeighborhoodAttention2D`` module as used at
    /root/reference/core/modules.py:509,517 (SURVEY.md §3.2 gives the op sequence):
    qkv Linear -> [3,B,heads,H,W,hd] -> q*scale -> qk+rpb -> softmax -> av -> merge -> proj.
    Parameter names match natten's (qkv, rpb, proj) so state_dicts interchange."""

    def __init__(self, dim, num_heads, kernel_size, dilation=1, bias=True, qkv_bias=True,
                 qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.scale = qk_scale or self.head_dim ** -0.5
        self.kernel_size, self.dilation = kernel_size, dilation or 1
        self.qkv = torch.nn.Linear(dim, dim * 3, bias=qkv_bias)
        if bias:
            self.rpb = torch.nn.Parameter(torch.zeros(num_heads, 2 * kernel_size - 1, 2 * kernel_size - 1))
            torch.nn.init.trunc_normal_(self.rpb, std=0.02, mean=0.0, a=-2.0, b=2.0)
        else:
            self.register_parameter("rpb", None)
        self.attn_drop = torch.nn.Dropout(attn_drop)
        self.proj = torch.nn.Linear(dim, dim)
        self.proj_drop = torch.nn.Dropout(proj_drop)

    def forward(self, x):
        B, H, W, C = x.shape
        qkv = self.qkv(x).reshape(B, H, W, 3, self.num_heads, self.head_dim).permute(3, 0, 4, 1, 2, 5)
        q, k, v = qkv[0], qkv[1], qkv[2]
        q = q * self.scale
        attn = natten2dqkrpb(q, k, self.rpb, self.kernel_size, self.dilation)
        attn = self.attn_drop(attn.softmax(dim=-1))
        x = natten2dav(attn, v, self.kernel_size, self.dilation)
        x = x.permute(0, 2, 3, 1, 4).reshape(B, H, W, C)
        return self.proj_drop(self.proj(x))
