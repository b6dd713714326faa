"""Multi-GPU plumbing: one process per GPU, work partitioned by video sequence.

Inference and per-sequence fine-tuning need NO collective (each sequence has its own weights,
reference ``train_online.py:184-186``); ``torch.distributed`` is used only for the barrier and the
max-over-ranks timing reduce, and -- for offline parent training / distillation -- the gradient
all-reduce (``allreduce_gradients``)."""
from __future__ import annotations

import os
from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def init_distributed(backend: str = "nccl") -> Tuple[int, int]:
    """(rank, world_size); initialises the process group from the torchrun environment."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(v: float, device="cuda") -> float:
    if not dist.is_initialized():
        return float(v)
    t = torch.tensor([v], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(v: float, device="cuda") -> float:
    if not dist.is_initialized():
        return float(v)
    t = torch.tensor([v], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def allreduce_flat(flat: torch.Tensor) -> None:
    """Sum ONE flat fp32 gradient buffer over the ranks in place: the data-parallel exchange step of offline
    parent training and distillation (61 MB per optimizer step for the full VGG; NCCL ring / NVLS over NVSwitch)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20) -> None:
    """Sum the fp32 gradients over ranks (data-parallel offline training / distillation):
    flat buckets of ~32 MB so a step's 61 MB payload is two NCCL launches; lr=0 tensors
    (``upscale*``) never carry a gradient and are skipped."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0
    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0
    for g in grads:
        bucket.append(g)
        size += g.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()
