"""Process-group helpers with the names and behaviour of the reference's
utils/distributed_utils.py:7-70 (env:// initialisation, one process per GPU, NCCL; barrier; scalar
all-reduce), plus the DistributedDataParallel wrap that the reference never reaches (SURVEY.md §0.4,
§8 e1: batch sharding, one gradient all-reduce per step over NVLink, BatchNorm statistics stay
per-process exactly as in the reference, where --syncBN is defined but unused)."""
from __future__ import annotations

import os
from types import SimpleNamespace

import torch
import torch.distributed as dist


def init_distributed_mode(args=None):
    """Same contract as the reference function: fills args.rank / world_size / gpu / distributed from the
    torchrun (or SLURM) environment and initialises the default group.  `args.dist_url` defaults to
    'env://' (the reference's argparse never defines it); `args.dist_backend` defaults to NCCL on CUDA
    machines and gloo otherwise (CPU tests)."""
    args = args if args is not None else SimpleNamespace()
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        args.rank = int(os.environ["RANK"])
        args.world_size = int(os.environ["WORLD_SIZE"])
        args.gpu = int(os.environ.get("LOCAL_RANK", 0))
    elif "SLURM_PROCID" in os.environ:
        args.rank = int(os.environ["SLURM_PROCID"])
        args.world_size = int(os.environ.get("SLURM_NTASKS", 1))
        args.gpu = args.rank % max(1, torch.cuda.device_count())
    else:
        args.rank, args.world_size, args.gpu, args.distributed = 0, 1, 0, False
        return args
    args.distributed = True
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(args.gpu)
    args.dist_backend = getattr(args, "dist_backend", None) or ("nccl" if use_cuda else "gloo")
    args.dist_url = getattr(args, "dist_url", None) or "env://"
    if not dist.is_initialized():
        kw = {"device_id": torch.device("cuda", args.gpu)} if (use_cuda and args.dist_backend == "nccl") else {}
        dist.init_process_group(backend=args.dist_backend, init_method=args.dist_url, world_size=args.world_size,
                                rank=args.rank, **kw)
    dist.barrier()
    return args


def cleanup():
    if is_dist_avail_and_initialized():
        dist.destroy_process_group()


def is_dist_avail_and_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_world_size() -> int:
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank() -> int:
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process() -> bool:
    return get_rank() == 0


def reduce_value(value, average=True):
    """All-reduce a tensor in place (sum, optionally divided by the world size); no-op on one process."""
    world = get_world_size()
    if world < 2:
        return value
    with torch.no_grad():
        dist.all_reduce(value)
        if average:
            value /= world
    return value


def wrap_ddp(model: torch.nn.Module, device=None) -> torch.nn.Module:
    """DistributedDataParallel over the default group.  BatchNorm buffers are NOT broadcast: statistics
    stay per-process, as in the reference (no SyncBN, no buffer sync)."""
    if get_world_size() < 2:
        return model
    from torch.nn.parallel import DistributedDataParallel as DDP

    if device is not None and torch.device(device).type == "cuda":
        return DDP(model, device_ids=[torch.device(device).index], broadcast_buffers=False, gradient_as_bucket_view=True)
    return DDP(model, broadcast_buffers=False)
