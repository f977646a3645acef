"""Training-loop pieces that mirror the reference's driver for the benchmark harness.

* ``DiceLoss``         — same arithmetic as utils/loss.py:170-206 (softmax, one-hot, per-class soft Dice
                         with smooth=1e-5, class weights, mean over classes) minus its ``.item()`` host syncs.
* ``build_training``   — optimiser / criteria exactly as train.py:156-160 (AdamW lr 1e-3 wd 1e-4,
                         CE(weight [1,4], label_smoothing 1e-3), DiceLoss(2)).
* ``train_one_epoch``  — same signature and step order as utils/train_eval_utils.py:120-166
                         (H2D copy, autocast forward, CE + Dice(weight [1,4]), zero_grad, backward, step,
                         loss.item(), argmax -> CPU metric update), with bf16 autocast instead of the
                         reference's fp16 + GradScaler (BASELINE.json configs[1]).
* ``synthetic_batches``— Kvasir-SEG-shaped synthetic data: N(0,1) RGB images and one
                         filled ellipse per mask covering 5-40 % of the area.
* ``ConfusionMetrics`` — duck-typed stand-in for the torchmetrics collection (reset/update/compute/to).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn


class DiceLoss(nn.Module):
    def __init__(self, n_classes: int):
        super().__init__()
        self.n_classes = n_classes

    def forward(self, inputs, target, weight=None, softmax=True):
        if softmax:
            inputs = torch.softmax(inputs, dim=1)
        onehot = torch.cat([(target == c) for c in range(self.n_classes)], dim=1).float()
        if inputs.shape != onehot.shape:
            raise ValueError("predict & target shape do not match")
        weight = [1.0] * self.n_classes if weight is None else weight
        loss = 0.0
        for c in range(self.n_classes):
            score, tgt = inputs[:, c], onehot[:, c]
            inter = (score * tgt).sum()
            dice = (2 * inter + 1e-5) / ((score * score).sum() + (tgt * tgt).sum() + 1e-5)
            loss = loss + (1 - dice) * weight[c]
        return loss / self.n_classes


class WeightedSmoothedCE(nn.Module):
    """nn.CrossEntropyLoss(weight=w, label_smoothing=eps) (train.py:157 of the reference) restated with
    graph-capturable tensor ops: ATen's weighted + smoothed kernel path breaks CUDA-graph capture.
        loss = [(1-eps) * sum_i w[y_i] * nll_i + eps/C * sum_i sum_c w_c * (-logp_ic)] / sum_i w[y_i]"""

    def __init__(self, weight: torch.Tensor, label_smoothing: float = 0.0):
        super().__init__()
        self.register_buffer("weight", weight.float())
        self.eps = float(label_smoothing)

    def forward(self, logits, target):
        logp = F.log_softmax(logits.float(), dim=1)                       # [B, C, ...]
        C = logits.shape[1]
        w = self.weight.view(1, C, *([1] * (logits.dim() - 2)))
        wy = self.weight[target]                                          # [B, ...]
        nll = -(logp.gather(1, target.unsqueeze(1)).squeeze(1) * wy).sum()
        smooth = -(logp * w).sum()
        return ((1.0 - self.eps) * nll + (self.eps / C) * smooth) / wy.sum()


def build_training(model: nn.Module, device, lr=1e-3, weight_decay=1e-4, smoothing=1e-3, fused=None,
                   capturable=False):
    dev = torch.device(device)
    fused = (dev.type == "cuda") if fused is None else fused
    optimizer = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=weight_decay, fused=fused,
                                  capturable=capturable and dev.type == "cuda")
    if capturable and dev.type == "cuda":
        criterion = WeightedSmoothedCE(torch.tensor([1.0, 4.0], device=dev), smoothing).to(dev)
    else:
        criterion = nn.CrossEntropyLoss(weight=torch.tensor([1.0, 4.0], device=dev), label_smoothing=smoothing)
    criterion_dice = DiceLoss(2).to(dev)
    return optimizer, criterion, criterion_dice


_DICE_WEIGHT = (1.0, 4.0)          # utils/train_eval_utils.py:142


def _fused_loss_args(output, labels, criterion, criterion_dice):
    """(ce_weight, label_smoothing) when the two criteria are the reference's pair and the fused kernel applies."""
    if not (output.is_cuda and isinstance(criterion_dice, DiceLoss) and criterion_dice.n_classes == output.shape[1]):
        return None
    if isinstance(criterion, WeightedSmoothedCE):
        w, eps = criterion.weight, criterion.eps
    elif isinstance(criterion, nn.CrossEntropyLoss) and criterion.reduction == "mean" and criterion.weight is not None:
        w, eps = criterion.weight, criterion.label_smoothing
    else:
        return None
    from .segloss import seg_loss_supported
    return (w, eps) if seg_loss_supported(output, labels) else None


def loss_fn(output, labels, criterion, criterion_dice):
    """criterion(output, labels) + criterion_dice(output, labels.unsqueeze(1).float(), weight=[1, 4])
    (utils/train_eval_utils.py:141-142); on a CUDA device one fused op (csrc/seg_loss.cu)."""
    fused = _fused_loss_args(output, labels, criterion, criterion_dice)
    if fused is not None:
        from .segloss import seg_loss
        return seg_loss(output, labels, fused[0], fused[1], _DICE_WEIGHT)
    return criterion(output, labels) + criterion_dice(output, labels.unsqueeze(1).float(), weight=list(_DICE_WEIGHT))


def train_step(model, optimizer, images, labels, criterion, criterion_dice, amp_dtype=torch.bfloat16):
    """forward + loss + backward + optimiser step on device-resident tensors; returns the loss tensor."""
    dev_type = images.device.type
    with torch.autocast(dev_type, dtype=amp_dtype, enabled=amp_dtype is not None):
        output = model(images)
        loss = loss_fn(output, labels, criterion, criterion_dice)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    optimizer.step()
    return loss, output


class GraphedTrainStep:
    """The whole training step (forward, loss, backward, gradient all-reduce, AdamW) captured once into a CUDA
    graph and replayed.

    One LM-Net step is ~2 000 kernel launches; enqueueing them from Python costs ~40 ms of host time, which
    caps the step no matter how fast the kernels are (tools/cpu_bound_probe.py).  Replaying a captured graph
    removes that cost: inputs are copied into static device buffers, `replay()` re-issues every kernel
    (ours through the C ABI, cuDNN/cuBLAS, NCCL) with the recorded arguments.
    Requirements met by this code base: no host synchronisation inside the step, static shapes, the library
    takes the *current* (capturing) stream, the optimiser is built with capturable=True.

    Data parallel (torch.distributed initialised, world > 1; mirrors what utils/distributed_utils.py:7-28 sets up):
    the graph holds forward + backward + one multi-tensor copy of the freshly assigned gradients into ONE flat buffer
    (letting autograd accumulate into pre-set views instead costs a read-modify-write kernel per parameter: 514 launches,
    0.9 ms); the flat buffer is averaged by a single NCCL all-reduce (15.9 MB over NVLink) and the optimiser steps on views
    of it, both launched right after the replay.  `capture_collective=True` records the all-reduce and the optimiser step
    inside the graph as well; measured on 2 x B200 this is slower (31.0 vs 30.7 ms/step) and NCCL's communicator teardown
    hangs while graphs with captured collectives exist, so it is off by default.  Parameters and buffers are broadcast from rank 0 once
    at construction (what DDP does); BatchNorm statistics stay per process, as in the reference.

    Construction has no side effects on the training trajectory: the warm-up iterations that CUDA-graph capture
    needs run on a snapshot — parameters, buffers (BatchNorm running statistics, num_batches_tracked) and the
    optimiser state are restored before the capture.

    Usage: step = GraphedTrainStep(...); loss, output = step(images, labels)  # tensors are static buffers.
    A batch whose shape differs from the example batch (a short last batch) runs eagerly.  If capture is not
    possible the step runs eagerly and says so: a RuntimeWarning carries the reason (also kept in
    `.fallback_reason`); with LMNET_REQUIRE_GRAPH=1 in the environment the constructor raises instead."""

    def __init__(self, model, optimizer, criterion, criterion_dice, example_images, example_labels,
                 amp_dtype=torch.bfloat16, warmup=3, capture_collective=False):
        import os
        import warnings

        import torch.distributed as dist

        self.model, self.optimizer = model, optimizer
        self.criterion, self.criterion_dice, self.amp_dtype = criterion, criterion_dice, amp_dtype
        self.graph, self.fallback_reason, self.library_launches_per_step = None, None, 0
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.images = torch.empty_like(example_images)
        self.labels = torch.empty_like(example_labels)
        self.images.copy_(example_images)
        self.labels.copy_(example_labels)
        self.flat = None
        self.collective_in_graph = False
        if self.world > 1:
            if isinstance(model, torch.nn.parallel.DistributedDataParallel):
                raise ValueError("pass the un-wrapped module: GraphedTrainStep does its own gradient all-reduce")
            with torch.no_grad():
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t, 0)
            params = [p for p in model.parameters() if p.requires_grad]
            self.params = params
            self.flat = torch.zeros(sum(p.numel() for p in params), dtype=params[0].dtype, device=params[0].device)
            self.views, off = [], 0
            for p in params:                       # the optimiser reads every gradient as a view of the flat buffer
                self.views.append(self.flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        if self.images.device.type != "cuda":
            self.fallback_reason = "not a CUDA device: eager execution"
            return
        attempts = [True, False] if (self.world > 1 and capture_collective) else [False]
        for with_collective in attempts:
            try:
                self._capture(warmup, with_collective)
                self.fallback_reason = None
                break
            except Exception as e:  # noqa: BLE001 - any capture failure means: next variant, then eager
                self.graph = None
                self.fallback_reason = f"{type(e).__name__}: {e}"
                torch.cuda.synchronize()
        if self.graph is None:
            msg = f"GraphedTrainStep: CUDA-graph capture failed, the step runs EAGERLY (~2x slower): {self.fallback_reason}"
            if os.environ.get("LMNET_REQUIRE_GRAPH") == "1":
                raise RuntimeError(msg)
            warnings.warn(msg, RuntimeWarning, stacklevel=2)

    # -- pieces -------------------------------------------------------------------------------------
    def _forward_backward(self, images, labels):
        with torch.autocast(images.device.type, dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
            output = self.model(images)
            loss = loss_fn(output, labels, self.criterion, self.criterion_dice)
        if self.flat is not None:
            # autograd ASSIGNS fresh gradient tensors when .grad is None; accumulating into pre-set views instead costs
            # one read-modify-write kernel per parameter (514 tiny launches, 0.9 ms of a 31 ms step at 2 GPUs)
            for p in self.params:
                p.grad = None
        else:
            self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if self.flat is not None:
            have = [(v, p.grad) for v, p in zip(self.views, self.params)]
            missing = [v for v, g in have if g is None]
            if missing:
                torch._foreach_zero_(missing)
            torch._foreach_copy_([v for v, g in have if g is not None], [g for v, g in have if g is not None])
            for v, p in zip(self.views, self.params):   # multi-tensor copy into the flat buffer, then hand the views over
                p.grad = v
        return loss, output

    def _reduce(self):
        import torch.distributed as dist

        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:                                      # gloo (CPU tests) has no AVG
            dist.all_reduce(self.flat)
            self.flat /= self.world

    def _eager(self, images, labels):
        loss, output = self._forward_backward(images, labels)
        if self.flat is not None:
            self._reduce()
        self.optimizer.step()
        return loss, output

    def _snapshot(self):
        import copy

        model_state = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        return model_state, copy.deepcopy(self.optimizer.state_dict())

    def _restore(self, snap):
        model_state, opt_state = snap
        with torch.no_grad():
            for k, v in self.model.state_dict().items():
                v.copy_(model_state[k])
            if opt_state["state"]:
                self.optimizer.load_state_dict(opt_state)
            else:
                # a fresh optimiser: its state was created by the warm-up.  Adam's initial state is all zeros, so
                # zeroing in place (same tensors the capture will record) is exactly "never stepped".
                for st in self.optimizer.state.values():
                    for t in st.values():
                        if torch.is_tensor(t):
                            t.zero_()

    def _capture(self, warmup, with_collective):
        from . import _lib

        snap = self._snapshot()
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._eager(self.images, self.labels)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
        finally:
            self._restore(snap)
        graph = torch.cuda.CUDAGraph()
        if self.flat is None:
            self.optimizer.zero_grad(set_to_none=True)
        before = _lib.launch_count()
        with torch.cuda.graph(graph):
            self.loss, self.output = self._forward_backward(self.images, self.labels)
            if self.flat is None:
                self.optimizer.step()
            elif with_collective:
                self._reduce()
                self.optimizer.step()
        self.library_launches_per_step = _lib.launch_count() - before    # lmnet_b200 kernels inside one replay
        self.graph = graph
        self.collective_in_graph = bool(with_collective and self.flat is not None)
        self._restore(snap)            # capture does not execute, but the optimiser may touch host-side counters

    def __call__(self, images, labels):
        if images.shape != self.images.shape or labels.shape != self.labels.shape:
            # e.g. a short last batch: the captured graph has static shapes, run this one eagerly
            self.loss, self.output = self._eager(images.to(self.images.device), labels.to(self.labels.device))
            return self.loss, self.output
        self.images.copy_(images, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        if self.graph is None:
            self.loss, self.output = self._eager(self.images, self.labels)
        else:
            self.graph.replay()
            if self.flat is not None and not self.collective_in_graph:
                self._reduce()
                self.optimizer.step()
        return self.loss, self.output


class ConfusionMetrics:
    """Accumulates a confusion matrix; compute() returns accuracy / Dice / IoU of class 1.

    Duck-types the torchmetrics collection the reference passes around (reset / update / compute / to).
    update() accepts tensors on any device; with CUDA tensors the 2x2 histogram is built on the GPU and
    stays there until compute(), so the training loop needs no per-step D2H of the [B,H,W] mask."""

    def __init__(self, num_classes=2):
        self.n = num_classes
        self.cm = None
        self.reset()

    def to(self, _device):
        return self

    def reset(self):
        self.cm = None

    def update(self, pred, labels):
        idx = labels.reshape(-1) * self.n + pred.reshape(-1)
        if idx.is_cuda and self.n * self.n <= 16:
            # torch.bincount synchronises the host on CUDA (it reads max(idx) to size its output), which would
            # serialise host and device every step; a compare-and-sum histogram needs no host round trip
            bins = torch.arange(self.n * self.n, device=idx.device, dtype=idx.dtype)
            cm = (idx.unsqueeze(1) == bins).sum(0)
        else:
            cm = torch.bincount(idx, minlength=self.n * self.n)
        self.cm = cm if self.cm is None else self.cm + cm.to(self.cm.device)

    def compute(self):
        cm = torch.zeros(self.n * self.n) if self.cm is None else self.cm.cpu()
        cm = cm.view(self.n, self.n).double()
        tp, fp, fn = cm[1, 1], cm[0, 1], cm[1, 0]
        return {"acc": float(cm.diag().sum() / cm.sum().clamp_min(1)),
                "dice": float(2 * tp / (2 * tp + fp + fn).clamp_min(1)),
                "iou": float(tp / (tp + fp + fn).clamp_min(1))}


_PREFETCH_STATE = {}


class _Prefetcher:
    """Copies the next host batch to the device on a side stream while the current step computes.

    Two persistent device staging buffers (no per-step allocation: fresh tensors + record_stream made the caching
    allocator grow for several steps and cost 1-3 ms/step on slower hosts, tools/e2e_probe.py).  Ordering: the copy of
    batch k+2 into a buffer waits for an event recorded on the consumer's stream after step k (which read that buffer) was
    enqueued; the consumer waits for the copy's event before it touches the buffer."""

    def __init__(self, loader, device):
        self.it, self.device = iter(loader), torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        # the side stream and the two staging buffers persist across epochs (per device): a fresh stream per epoch
        # draws from a different allocator pool, i.e. two cudaMalloc calls of a batch each at the start of every epoch
        # (a 20-40 ms stall in the first step, tools/e2e_probe.py) and memory that is only returned at the next GC
        state = _PREFETCH_STATE.get(self.device) if self.device.type == "cuda" else None
        if state is None and self.device.type == "cuda":
            state = _PREFETCH_STATE[self.device] = {"stream": torch.cuda.Stream(self.device), "bufs": [None, None],
                                                    "free_ev": [None, None]}
        self.stream = state["stream"] if state is not None else None
        self.bufs = state["bufs"] if state is not None else [None, None]
        self.free_ev = state["free_ev"] if state is not None else [None, None]
        self.slot = 0
        self.next, self.ready, self.in_use = None, None, None
        self._load()

    def _load(self):
        try:
            batch = next(self.it)
        except StopIteration:
            self.next = None
            return
        if self.stream is None or all(t.device == self.device for t in batch):     # nothing to stage
            self.next = (tuple(t.to(self.device) for t in batch), None)
            return
        i = self.slot
        self.slot ^= 1
        with torch.cuda.stream(self.stream):
            if self.free_ev[i] is not None:
                self.stream.wait_event(self.free_ev[i])
            buf = self.bufs[i]
            if buf is None or any(b.shape != t.shape or b.dtype != t.dtype for b, t in zip(buf, batch)):
                buf = self.bufs[i] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch)
            for b, t in zip(buf, batch):
                b.copy_(t, non_blocking=True)
            self.ready = torch.cuda.Event()
            self.ready.record(self.stream)
        self.next = (buf, i)

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device) if self.stream is not None else None
        if self.in_use is not None:               # everything that read the previous batch is enqueued by now
            ev = torch.cuda.Event()
            ev.record(cur)
            self.free_ev[self.in_use] = ev
            self.in_use = None
        if self.next is None:
            raise StopIteration
        batch, i = self.next
        if i is not None:
            cur.wait_event(self.ready)
            self.in_use = i
        self._load()
        return batch


def train_one_epoch(model, optimizer, metric_collection=None, num_classes=2, data_loader=None, device=0,
                    criterion=None, scaler=None, criterion_dice=None, amp_dtype=torch.bfloat16,
                    metrics_on_device=True, prefetch=True, defer_loss_read=True, step_fn=None):
    """Reference-shaped epoch loop (utils/train_eval_utils.py:120-166): H2D copy of every batch, autocast
    forward, CE + Dice(weight [1,4]), zero_grad, backward, step, loss.item() every step, argmax ->
    metric update.  `scaler` is accepted for signature compatibility; a non-None value selects the
    autocast branch exactly as in the reference, but with bf16 no loss scaling is needed.

    Two host-side overheads of the reference loop are removed without changing what
    is computed: the next batch is copied on a side stream while the current step runs (`prefetch`), and
    the confusion matrix is accumulated on the GPU instead of shipping the [B,H,W] int64 mask to the CPU
    every step (`metrics_on_device`; set False for the reference's exact D2H behaviour).  `defer_loss_read` reads
    each step's loss one step late (same values, same total)."""
    dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
    if dev.type == "cuda" and (dev.index if dev.index is not None else torch.cuda.current_device()) != torch.cuda.current_device():
        # the C-ABI launches go to the CURRENT device's stream: make the loop's device current for its duration
        with torch.cuda.device(dev):
            return train_one_epoch(model, optimizer, metric_collection, num_classes, data_loader, dev, criterion, scaler,
                                   criterion_dice, amp_dtype, metrics_on_device, prefetch, defer_loss_read, step_fn)
    model.train()
    if metric_collection is not None:
        metric_collection.reset()
    total_loss = 0.0
    use_amp = scaler is not None
    batches = _Prefetcher(data_loader, dev) if prefetch else (
        (i.to(dev, non_blocking=True), l.to(dev, non_blocking=True)) for i, l in data_loader)
    # loss.item() every step, as in the reference — but read one step late through pinned memory, so the
    # host keeps enqueueing the next step instead of draining the GPU pipeline at every iteration
    pipelined = dev.type == "cuda" and defer_loss_read
    slots = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)] if pipelined else None
    pending, step = None, 0
    for images, labels in batches:
        if step_fn is not None:          # e.g. a GraphedTrainStep built from the same model / optimiser / criteria
            loss, output = step_fn(images, labels)
            labels = getattr(step_fn, "labels", labels)
        else:
            loss, output = train_step(model, optimizer, images, labels, criterion, criterion_dice,
                                      amp_dtype if use_amp else None)
        with torch.no_grad():
            if metric_collection is not None:
                pred = output.argmax(1).detach()
                if metrics_on_device:
                    metric_collection.update(pred, labels)
                else:
                    metric_collection.update(pred.cpu(), labels.detach().cpu())
            if pipelined:
                slot = slots[step % 2]
                slot.copy_(loss.detach().float(), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                if pending is not None:
                    pending[1].synchronize()
                    total_loss += float(pending[0])
                pending = (slot, ev)
            else:
                total_loss += loss.item()
        step += 1
    if pending is not None:
        pending[1].synchronize()
        total_loss += float(pending[0])
    return total_loss


def synthetic_batches(n_batches, batch, res, seed=0, pin=True):
    """List of (images fp32 [B,3,R,R], masks int64 [B,R,R]) host tensors, optionally pinned."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(res, dtype=torch.float32), torch.arange(res, dtype=torch.float32), indexing="ij")
    out = []
    for _ in range(n_batches):
        images = torch.randn(batch, 3, res, res, generator=g)
        masks = torch.empty(batch, res, res, dtype=torch.int64)
        for b in range(batch):
            frac = 0.05 + 0.35 * float(torch.rand(1, generator=g))
            aspect = 0.6 + 0.8 * float(torch.rand(1, generator=g))
            ry = math.sqrt(frac * res * res / math.pi * aspect)
            rx = frac * res * res / math.pi / ry
            cy = ry + float(torch.rand(1, generator=g)) * max(1.0, res - 2 * ry)
            cx = rx + float(torch.rand(1, generator=g)) * max(1.0, res - 2 * rx)
            masks[b] = ((((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) <= 1.0).long()
        if pin and torch.cuda.is_available():
            images, masks = images.pin_memory(), masks.pin_memory()
        out.append((images, masks))
    return out
