"""The training loss of the reference loop as one fused op: weighted, label-smoothed cross entropy + weighted Dice.

Reference semantics (/root/reference/utils/train_eval_utils.py:141-142 with the criteria built in
/root/reference/train.py:157-158 and /root/reference/utils/loss.py:170-206):

    loss = CrossEntropyLoss(weight=w, label_smoothing=eps)(output, labels)
         + DiceLoss(n_classes)(output, labels.unsqueeze(1).float(), weight=dw)          # softmax=True

`seg_loss` computes exactly this sum (fp32 arithmetic on the fp32 / bf16 / fp16 logits, as autocast does) with two
kernel launches forward and one backward through the C ABI (csrc/seg_loss.cu); the stock op sequence is ~45 launches.
"""
from __future__ import annotations

import torch

from . import _lib as L


class _SegLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ce_w, dice_w, eps):
        L.require_cuda(logits)
        logits = logits.contiguous()
        labels = labels.contiguous()
        B, C = logits.shape[:2]
        HW = logits[0, 0].numel()
        lib = L.lib()
        stats = torch.empty(lib.lmnet_seg_loss_stats_floats(C), dtype=torch.float32, device=logits.device)
        ws = torch.empty(int(lib.lmnet_seg_loss_workspace_bytes(C)), dtype=torch.uint8, device=logits.device)
        rc = lib.lmnet_seg_loss_fwd(L.ptr(logits), L.ptr(labels), L.ptr(ce_w), L.ptr(dice_w), float(eps), L.ptr(stats),
                                    L.ptr(ws), ws.numel(), B, C, HW, L.dtype_code(logits), L.stream_ptr())
        L.check(rc, "seg_loss_fwd")
        ctx.save_for_backward(logits, labels, ce_w, dice_w, stats)
        ctx.eps = float(eps)
        return stats[0].clone()

    @staticmethod
    def backward(ctx, dloss):
        logits, labels, ce_w, dice_w, stats = ctx.saved_tensors
        B, C = logits.shape[:2]
        HW = logits[0, 0].numel()
        dloss = dloss.detach().float().contiguous()
        dlogits = torch.empty_like(logits)
        rc = L.lib().lmnet_seg_loss_bwd(L.ptr(logits), L.ptr(labels), L.ptr(ce_w), L.ptr(dice_w), ctx.eps, L.ptr(stats),
                                        L.ptr(dloss), L.ptr(dlogits), B, C, HW, L.dtype_code(logits), L.stream_ptr())
        L.check(rc, "seg_loss_bwd")
        return dlogits, None, None, None, None


def seg_loss_supported(logits: torch.Tensor, labels: torch.Tensor) -> bool:
    return (logits.is_cuda and logits.dim() >= 3 and labels.dtype == torch.int64 and labels.shape == logits.shape[:1] + logits.shape[2:]
            and logits.dtype in (torch.float32, torch.bfloat16, torch.float16)
            and bool(L.lib().lmnet_seg_loss_supported(int(logits.shape[1]), L.dtype_code(logits))))


_WEIGHT_CACHE = {}


def _device_weights(w, dev):
    """fp32 weights on `dev`.  Python sequences are uploaded once per (values, device): a host-to-device copy from
    pageable memory is not allowed while a CUDA graph is being captured."""
    if isinstance(w, torch.Tensor):
        return w.detach().to(device=dev, dtype=torch.float32).contiguous()
    key = (tuple(float(x) for x in w), dev)
    t = _WEIGHT_CACHE.get(key)
    if t is None:
        t = _WEIGHT_CACHE[key] = torch.tensor(key[0], dtype=torch.float32, device=dev)
    return t


def seg_loss(logits, labels, ce_weight, label_smoothing, dice_weight):
    """CrossEntropyLoss(weight=ce_weight, label_smoothing)(logits, labels) + DiceLoss(C)(logits, labels[:, None].float(),
    weight=dice_weight).  logits [B,C,...] on a CUDA device, labels int64 [B,...] with values in [0, C)."""
    dev = logits.device
    return _SegLoss.apply(logits, labels, _device_weights(ce_weight, dev), _device_weights(dice_weight, dev),
                          float(label_smoothing))
