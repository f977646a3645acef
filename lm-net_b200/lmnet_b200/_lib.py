"""ctypes binding of ``liblmnet_b200.so`` (the C ABI declared in ``include/lmnet_b200.h``).

The shared library is built in-tree by ``lm-net_b200/csrc/Makefile`` (see ``__graft_entry__.build``).
There is no fallback of any kind: a missing library raises at import of the first op, and every
entry point raises ``RuntimeError`` on a non-zero status.  PyTorch supplies device memory and the
current CUDA stream; nothing here touches tensor contents on the host.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# LMNET_B200_LIB points at an alternative build of the same ABI (kernel A/B experiments); the default is the in-tree library.
LIB_PATH = os.environ.get("LMNET_B200_LIB") or os.path.join(_HERE, "liblmnet_b200.so")

F32, BF16, F16 = 0, 1, 2
_DTYPES = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}


class View5(Structure):
    _fields_ = [("ptr", c_void_p), ("sb", c_int64), ("sh", c_int64), ("sw", c_int64), ("sn", c_int64)]


class NADims(Structure):
    _fields_ = [("B", c_int32), ("H", c_int32), ("W", c_int32), ("heads", c_int32), ("D", c_int32),
                ("kernel_size", c_int32), ("dilation", c_int32)]


class DwParams(Structure):
    _fields_ = [("w", c_void_p * 4), ("gamma", c_void_p * 4), ("beta", c_void_p * 4),
                ("running_mean", c_void_p * 4), ("running_var", c_void_p * 4)]


class DwGrads(Structure):
    _fields_ = [("dw", c_void_p * 4), ("dgamma", c_void_p * 4), ("dbeta", c_void_p * 4)]


class DwDims(Structure):
    _fields_ = [("B", c_int32), ("E", c_int32), ("H", c_int32), ("W", c_int32)]


class BnDims(Structure):
    _fields_ = [("B", c_int32), ("C", c_int32), ("HW", c_int64)]


ACT_CODES = {"none": 0, "hardswish": 1, "gelu": 2, "relu": 3}


class UpsampleDims(Structure):
    _fields_ = [("planes", c_int64), ("H", c_int32), ("W", c_int32)]


class UpsampleClDims(Structure):
    _fields_ = [("B", c_int64), ("H", c_int32), ("W", c_int32), ("C", c_int32)]


class WgradDims(Structure):
    _fields_ = [("B", c_int32), ("M", c_int32), ("N1", c_int32), ("N2", c_int32), ("P", c_int64)]


class PgemmDims(Structure):
    _fields_ = [("B", c_int32), ("P", c_int64), ("N", c_int32), ("K1", c_int32), ("K2", c_int32)]


class PgemmWeights(Structure):
    _fields_ = [("w1", c_void_p), ("w1_sn", c_int64), ("w1_sk", c_int64), ("gate", c_void_p), ("gate_on_n", c_int32),
                ("w2", c_void_p), ("w2_sn", c_int64), ("w2_sk", c_int64)]


class PoolDims(Structure):
    _fields_ = [("B", c_int32), ("Ho", c_int32), ("Wo", c_int32), ("C", c_int32), ("factor", c_int32)]


class Conv3x3Dims(Structure):
    _fields_ = [("B", c_int32), ("H", c_int32), ("W", c_int32), ("Cin", c_int32), ("Cout", c_int32), ("stride", c_int32)]


_lib = None


def lib():
    """Load the CUDA library once; raise loudly if it was never built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"lmnet_b200: {LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C lm-net_b200/csrc`. There is no CPU or PyTorch fallback for these operators.")
    L = ctypes.CDLL(LIB_PATH)
    L.lmnet_abi_version.restype = c_int
    if L.lmnet_abi_version() != 1:
        raise ImportError("lmnet_b200: ABI version mismatch between liblmnet_b200.so and the Python host")
    L.lmnet_status_string.restype = ctypes.c_char_p
    L.lmnet_status_string.argtypes = [c_int]
    L.lmnet_launch_count.restype = c_uint64
    pv, pd = POINTER(View5), POINTER(NADims)
    L.lmnet_na2d_fwd.argtypes = [pv, pv, pv, c_void_p, pv, c_void_p, pd, c_float, c_int, c_void_p]
    L.lmnet_na2d_bwd_workspace_bytes.restype = c_size_t
    L.lmnet_na2d_bwd_workspace_bytes.argtypes = [pd]
    L.lmnet_na2d_bwd.argtypes = [pv, pv, pv, c_void_p, pv, pv, pv, pv, c_void_p, c_void_p, c_size_t, pd, c_float,
                                 c_int, c_void_p]
    L.lmnet_na2d_qk_fwd.argtypes = [pv, pv, c_void_p, c_void_p, pd, c_int, c_void_p]
    L.lmnet_na2d_qk_bwd_workspace_bytes.restype = c_size_t
    L.lmnet_na2d_qk_bwd_workspace_bytes.argtypes = [pd]
    L.lmnet_na2d_qk_bwd.argtypes = [pv, pv, c_void_p, pv, pv, c_void_p, c_void_p, c_size_t, pd, c_int, c_void_p]
    L.lmnet_na2d_av_fwd.argtypes = [c_void_p, pv, pv, pd, c_int, c_void_p]
    L.lmnet_na2d_av_bwd.argtypes = [c_void_p, pv, pv, c_void_p, pv, pd, c_int, c_void_p]
    pp, pg, pdd = POINTER(DwParams), POINTER(DwGrads), POINTER(DwDims)
    L.lmnet_reparam_dw_workspace_bytes.restype = c_size_t
    L.lmnet_reparam_dw_workspace_bytes.argtypes = [pdd, c_int]
    L.lmnet_reparam_dw_train_fwd.argtypes = [c_void_p, pp, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                             c_float, POINTER(c_void_p), c_void_p, c_size_t, pdd, c_int, c_void_p]
    L.lmnet_reparam_dw_train_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, pp, c_void_p, c_void_p, c_void_p,
                                             pg, c_void_p, c_size_t, pdd, c_int, c_void_p]
    L.lmnet_reparam_dw_eval_fwd.argtypes = [c_void_p, pp, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_size_t,
                                            pdd, c_int, c_void_p]
    L.lmnet_reparam_dw_train_fwd_gram.argtypes = [c_void_p, pp, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                                  c_float, POINTER(c_void_p), c_void_p, POINTER(c_int), c_void_p, c_size_t,
                                                  pdd, c_int, c_void_p]
    L.lmnet_reparam_dw_train_bwd_gram.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, pp, c_void_p, c_void_p, c_void_p,
                                                  c_void_p, pg, c_void_p, c_size_t, pdd, c_int, c_void_p]
    L.lmnet_seg_loss_workspace_bytes.restype = c_size_t
    L.lmnet_seg_loss_workspace_bytes.argtypes = [c_int]
    L.lmnet_seg_loss_supported.argtypes = [c_int, c_int]
    L.lmnet_seg_loss_stats_floats.argtypes = [c_int]
    L.lmnet_seg_loss_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_size_t, c_int64, c_int,
                                     c_int64, c_int, c_void_p]
    L.lmnet_seg_loss_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                     c_int64, c_int, c_void_p]
    pbd = POINTER(BnDims)
    L.lmnet_bn_act_workspace_bytes.restype = c_size_t
    L.lmnet_bn_act_workspace_bytes.argtypes = [pbd]
    L.lmnet_bn_act_fwd.argtypes = [c_void_p] * 9 + [c_float, c_float, c_int, c_int, c_void_p, c_size_t, pbd, c_int, c_void_p]
    L.lmnet_bn_act_bwd.argtypes = [c_void_p] * 9 + [c_int, c_void_p, c_size_t, pbd, c_int, c_void_p]
    L.lmnet_bn_act_cl_supported.argtypes = [pbd, c_int]
    L.lmnet_bn_act_cl_fwd.argtypes = L.lmnet_bn_act_fwd.argtypes
    L.lmnet_bn_act_cl_bwd.argtypes = L.lmnet_bn_act_bwd.argtypes
    L.lmnet_wgrad_1x1_cl_supported.argtypes = [POINTER(WgradDims), c_int, c_int, c_int]
    L.lmnet_wgrad_1x1_cl_workspace_bytes.restype = c_size_t
    L.lmnet_wgrad_1x1_cl_workspace_bytes.argtypes = [POINTER(WgradDims), c_int, c_int]
    L.lmnet_wgrad_1x1_cl.argtypes = [c_void_p] * 5 + [c_void_p, c_size_t, POINTER(WgradDims), c_int, c_int, c_int, c_void_p]
    L.lmnet_wgrad_1x1_cl_sum.argtypes = L.lmnet_wgrad_1x1_cl.argtypes
    L.lmnet_pointwise_grads.argtypes = [c_void_p] * 4 + [c_int64, c_int64] + [c_void_p] * 4 + [c_int, c_int, c_int, c_int, c_void_p]
    L.lmnet_layer_norm_supported.argtypes = [c_int]
    L.lmnet_layer_norm_workspace_bytes.restype = c_size_t
    L.lmnet_layer_norm_workspace_bytes.argtypes = [c_int64, c_int]
    L.lmnet_layer_norm_fwd.argtypes = [c_void_p] * 6 + [c_int64, c_int, c_float, c_int, c_void_p]
    L.lmnet_layer_norm_bwd.argtypes = [c_void_p] * 8 + [c_void_p, c_size_t, c_int64, c_int, c_int, c_void_p]
    pwd = POINTER(WgradDims)
    L.lmnet_wgrad_1x1_supported.argtypes = [pwd, c_int]
    L.lmnet_wgrad_1x1_workspace_bytes.restype = c_size_t
    L.lmnet_wgrad_1x1_workspace_bytes.argtypes = [pwd]
    L.lmnet_wgrad_1x1.argtypes = [c_void_p] * 5 + [c_void_p, c_size_t, pwd, c_int, c_void_p]
    ppg = POINTER(PgemmDims)
    L.lmnet_pixel_gemm_supported.argtypes = [ppg, c_int, c_int, c_int, c_int]
    L.lmnet_pixel_gemm_stats_ctas.argtypes = [ppg, c_int, c_int]
    L.lmnet_pixel_gemm.argtypes = [c_void_p, c_int, c_void_p, POINTER(PgemmWeights), c_void_p, c_void_p, c_int,
                                   c_void_p, ppg, c_int, c_void_p]
    L.lmnet_bn_act_fwd_stats.argtypes = [c_void_p, c_void_p, c_int] + [c_void_p] * 8 + [c_float, c_float, c_int, c_void_p,
                                                                                      c_size_t, pbd, c_int, c_void_p]
    pcv = POINTER(Conv3x3Dims)
    L.lmnet_conv3x3_fwd_supported.argtypes = [pcv, c_int]
    L.lmnet_conv3x3_fwd.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, pcv, c_int, c_void_p]
    L.lmnet_conv3x3_wgrad_supported.argtypes = [pcv, c_int]
    L.lmnet_conv3x3_wgrad_workspace_bytes.restype = c_size_t
    L.lmnet_conv3x3_wgrad_workspace_bytes.argtypes = [pcv]
    L.lmnet_conv3x3_wgrad.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, pcv, c_int, c_void_p]
    L.lmnet_conv3x3_fwd_strided.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, pcv, c_int, c_void_p]
    L.lmnet_conv3x3_wgrad_strided.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, pcv, c_int,
                                              c_void_p]
    L.lmnet_avgpool_cl_fwd.argtypes = [c_void_p, c_void_p, POINTER(PoolDims), c_int, c_void_p]
    L.lmnet_avgpool_cl_bwd.argtypes = [c_void_p, c_void_p, POINTER(PoolDims), c_int, c_void_p]
    L.lmnet_se_gate_supported.argtypes = [c_int, c_int, c_int]
    L.lmnet_se_gate_fwd.argtypes = [c_void_p] * 8 + [c_int, c_int, c_int, c_void_p]
    L.lmnet_se_gate_bwd.argtypes = [c_void_p] * 11 + [c_int, c_int, c_int, c_void_p]
    pud = POINTER(UpsampleDims)
    L.lmnet_upsample2x_fwd.argtypes = [c_void_p, c_void_p, pud, c_int, c_void_p]
    L.lmnet_upsample2x_bwd.argtypes = [c_void_p, c_void_p, pud, c_int, c_void_p]
    pucd = POINTER(UpsampleClDims)
    L.lmnet_upsample2x_cl_fwd.argtypes = [c_void_p, c_void_p, pucd, c_int, c_void_p]
    L.lmnet_upsample2x_cl_bwd.argtypes = [c_void_p, c_void_p, pucd, c_int, c_void_p]
    L.lmnet_profile_enable.argtypes = [c_int]
    L.lmnet_profile_kernel_name.restype = ctypes.c_char_p
    L.lmnet_profile_kernel_name.argtypes = [c_int]
    L.lmnet_profile_collect.argtypes = [POINTER(ctypes.c_double), POINTER(c_uint64), POINTER(ctypes.c_double), c_int]
    _lib = L
    return L


def profile_enable(on: bool) -> None:
    check(lib().lmnet_profile_enable(1 if on else 0), "profile_enable")


def profile_collect() -> dict:
    """{kernel name: {"ms": total ms, "launches": n, "alg_bytes": total algorithmic bytes}} since enable()."""
    n = lib().lmnet_profile_num_kernels()
    ms, cnt, by = (ctypes.c_double * n)(), (c_uint64 * n)(), (ctypes.c_double * n)()
    check(lib().lmnet_profile_collect(ms, cnt, by, n), "profile_collect")
    out = {}
    for i in range(n):
        if cnt[i]:
            out[lib().lmnet_profile_kernel_name(i).decode()] = {"ms": ms[i], "launches": int(cnt[i]), "alg_bytes": by[i]}
    return out


def launch_count() -> int:
    return int(lib().lmnet_launch_count())


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().lmnet_status_string(rc).decode()
        raise RuntimeError(f"lmnet_b200.{what} failed: {msg} (status {rc})")


def dtype_code(t: torch.Tensor) -> int:
    """dtype enum of the launch's main tensor.  Every launch passes through here, so this is also where the device
    contract is enforced: the library launches on the CURRENT device's current stream, and a tensor that lives on a
    different GPU would be touched from the wrong context (illegal address or unsynchronised execution)."""
    if t.is_cuda and t.device.index != torch.cuda.current_device():
        raise RuntimeError(
            f"lmnet_b200: tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
            "call torch.cuda.set_device(...) or wrap the call in `with torch.cuda.device(tensor.device):`")
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"lmnet_b200 supports float32 / bfloat16 / float16 activations, got {t.dtype}") from None


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("lmnet_b200 operators run on CUDA (sm_100a) tensors only; there is no CPU path")


def stream_ptr() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def view5(t: torch.Tensor, order: str) -> View5:
    """Describe a 5-D tensor with unit stride on its last dim as a logical [B,H,W,heads,D] view.

    order = "bhwnd" for the fused layout [B,H,W,heads,D]; "bnhwd" for natten's unfused layout
    [B,heads,H,W,D].  Any strides are accepted on the first four dims."""
    if t.dim() != 5 or t.stride(4) != 1:
        raise ValueError("expected a 5-D tensor with a contiguous last (head_dim) dimension")
    s = t.stride()
    if order == "bhwnd":
        return View5(t.data_ptr(), s[0], s[1], s[2], s[3])
    if order == "bnhwd":
        return View5(t.data_ptr(), s[0], s[2], s[3], s[1])
    raise ValueError(order)


def na_dims(B, H, W, heads, D, kernel_size, dilation) -> NADims:
    return NADims(B, H, W, heads, D, kernel_size, dilation)


def _void4(tensors):
    arr = (c_void_p * 4)()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def dw_params(w, gamma=None, beta=None, running_mean=None, running_var=None) -> DwParams:
    none4 = (None,) * 4
    p = DwParams()
    p.w = _void4(w)
    p.gamma = _void4(gamma or none4)
    p.beta = _void4(beta or none4)
    p.running_mean = _void4(running_mean or none4)
    p.running_var = _void4(running_var or none4)
    return p


def dw_grads(dw, dgamma, dbeta) -> DwGrads:
    g = DwGrads()
    g.dw = _void4(dw)
    g.dgamma = _void4(dgamma)
    g.dbeta = _void4(dbeta)
    return g


__all__ = ["lib", "check", "dtype_code", "require_cuda", "stream_ptr", "ptr", "view5", "na_dims", "dw_params",
           "dw_grads", "View5", "NADims", "DwParams", "DwGrads", "DwDims", "BnDims", "ACT_CODES", "WgradDims", "PgemmDims", "PgemmWeights", "PoolDims", "Conv3x3Dims", "UpsampleDims", "UpsampleClDims", "byref", "launch_count", "LIB_PATH"]
