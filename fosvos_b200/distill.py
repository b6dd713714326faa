"""Teacher -> student distillation step (reference ``/root/reference/src/mimic.py:144-218``) on the VGG path.

``calculate_loss`` there runs, per minibatch: student forward, teacher forward (detached), the criterion on
each of the five output maps, ``loss = (1 - epoch / n_epochs) * sum(losses[:-1]) + losses[-1]``,
``loss /= avg_grad_every_n``, backward; every ``avg_grad_every_n`` minibatches ``optimizer.step()`` and
``zero_grad()``.  ``MimicTrainer.step`` is that loop body with the criterion (MSE / L1 / class-balanced CE,
``mimic.py:76-83``) and its gradient fused into one kernel per map and Adam as one launch.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import ops
from .networks import OSVOS_VGG
from .optim import FusedAdam
from .sharding import allreduce_flat


class MimicTrainer:
    def __init__(self, student: OSVOS_VGG, teacher: Optional[OSVOS_VGG], criterion: str = "MSE", learning_rate: float = 1e-3,
                 weight_decay: float = 0.0002, avg_grad_every_n: int = 5, data_parallel: bool = False):
        if criterion not in ("MSE", "L1", "CBCEL"):
            raise Exception('Unknown loss function')                       # mimic.py:84
        dev = next(student.parameters()).device
        L.require_device(dev)
        self.student, self.teacher, self.criterion = student, teacher, criterion
        self.n = int(avg_grad_every_n)
        self.data_parallel = bool(data_parallel)
        params = dict(student.named_parameters())
        names = student._grad_names()
        # one flat gradient buffer: the data-parallel all-reduce is a single NCCL call, no packing copies
        total = sum(params[n].numel() for n in names)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grads: Dict[str, torch.Tensor] = {}
        off = 0
        for n in names:
            p = params[n]
            p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)
            self.grads[n] = p.grad
            off += p.numel()
        # mimic.py:74: Adam over net_student.parameters() (the up-sampling weights never get a gradient here)
        self.optimizer = FusedAdam([params[n] for n in names], lr=learning_rate, weight_decay=weight_decay)
        self.counter = 0

    def _criterion(self, o_s: torch.Tensor, o_t: torch.Tensor, scale: float):
        if self.criterion == "CBCEL":
            loss, stats = ops.bal_loss_fwd(o_s, o_t, True)                 # criterion(o_s, o_t): size_average default True
            return loss, ops.bal_loss_bwd(o_s, o_t, True, stats, None, scale)
        return ops.pixel_loss(o_s, o_t, "mse" if self.criterion == "MSE" else "l1", False, scale)

    @torch.no_grad()
    def step(self, frames: torch.Tensor, ground_truth: Optional[torch.Tensor], epoch: int, n_epochs: int) -> torch.Tensor:
        """One minibatch of ``calculate_loss(mode='train')``.  ``ground_truth=None`` learns from the teacher."""
        outs, _, _, saved = self.student._run_forward(frames, save=True)
        if ground_truth is None:
            targets = self.teacher._run_forward(frames, save=False)[0]
        else:
            targets = [ground_truth] * 5
        w_side = 1.0 - epoch / n_epochs                                     # mimic.py:217
        douts: List[Optional[torch.Tensor]] = [None] * 5
        total = None
        for i in range(5):
            w = 1.0 if i == 4 else w_side
            loss, douts[i] = self._criterion(outs[i], targets[i], w / self.n)
            total = w * loss if total is None else total + w * loss
        self.student._run_backward(saved, douts, self.grads)
        self.counter += 1
        if self.counter % self.n == 0:
            if self.data_parallel:
                allreduce_flat(self.flat_grad)
            self.optimizer.step_and_zero()
            self.counter = 0
        return total
