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


class GradBuckets:
    """Flat fp32 gradient buffer of a data-parallel trainer, laid out bucket-major in the order the backward pass
    finishes the layers: bucket k = the convs of stage 4 - k (+ the ``side_prep`` conv hanging off that stage); the last
    bucket (stage 0) also carries the 1x1 heads (``score_dsn``, ``fuse``).  ``p.grad`` of every parameter is a view into
    it, so each bucket is ONE contiguous all-reduce with no packing copies (28.6 / 23.9 / 6.0 / 1.0 / 0.2 MB for the full
    VGG).  The lr = 0 up-sampling weights never carry a gradient and are not part of it."""

    def __init__(self, net, params, device):
        names = net._grad_names()
        order = []
        self.bucket_names = []
        for k in range(5):
            si = 4 - k
            mine = [n for n in names if n.startswith(f"stages.{si}.")]
            if si > 0:
                mine += [n for n in names if n.startswith(f"side_prep.{si - 1}.")]
            if k == 4:
                mine += [n for n in names if n.startswith(("score_dsn.", "fuse."))]
            self.bucket_names.append(mine)
            order += mine
        missing = [n for n in names if n not in order]
        if missing:
            raise RuntimeError(f"GradBuckets: parameters without a bucket: {missing}")
        self.n_buckets = 5
        self.flat = torch.zeros(sum(params[n].numel() for n in order), dtype=torch.float32, device=device)
        self._views = {}
        self.ranges = []
        off = 0
        for mine in self.bucket_names:
            lo = off
            for n in mine:
                p = params[n]
                self._views[n] = self.flat[off:off + p.numel()].view(p.shape)
                off += p.numel()
            self.ranges.append((lo, off))
        self._fold_tables = {}

    def view(self, name: str) -> torch.Tensor:
        return self._views[name]

    def fold_table(self, bucket: int, wgrad_ws, grads):
        """Device table folding this bucket's tensor-core weight-gradient accumulators into its ``.grad`` views."""
        if bucket not in self._fold_tables:
            from . import ops
            ent = [(wgrad_ws[n[:-len(".weight")]], grads[n]) for n in self.bucket_names[bucket]
                   if wgrad_ws is not None and n.endswith(".weight") and n[:-len(".weight")] in wgrad_ws]
            self._fold_tables[bucket] = ops.fold_table(ent, self.flat.device) if ent else None
        return self._fold_tables[bucket]

    def allreduce(self, bucket: int):
        """Start the sum of one bucket over the ranks on the CURRENT stream's dependency chain; returns the async work
        handle (None in a single-process run)."""
        lo, hi = self.ranges[bucket]
        if hi == lo or not (dist.is_initialized() and dist.get_world_size() > 1):
            return None
        return dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True)
