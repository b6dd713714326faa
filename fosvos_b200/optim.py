"""Fused multi-tensor SGD and the reference's per-group optimizer policy.

``FusedSGD`` is a ``torch.optim.Optimizer`` with the same constructor arguments, param-group
keys and ``state['momentum_buffer']`` layout as ``torch.optim.SGD`` (momentum, dampening 0,
no Nesterov -- the only configuration the reference uses, ``util/network_provider.py:98-159``),
but ``step()`` is ONE kernel launch over every tensor instead of one or more per tensor.

``get_optimizer_online`` / ``get_optimizer_offline`` rebuild exactly the param groups of
``VGGOnlineProvider.get_optimizer`` (``network_provider.py:144-159``) and
``VGGOfflineProvider.get_optimizer`` (``:98-125``).
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch.optim import Optimizer

from . import _lib as L
from . import ops


class FusedSGD(Optimizer):
    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0, weight_decay=0.0, nesterov=False):
        if dampening != 0 or nesterov:
            raise ValueError("FusedSGD implements the reference's configuration only: dampening=0, nesterov=False")
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        super().__init__(params, defaults)
        moms = {g["momentum"] for g in self.param_groups}
        if len(moms) != 1:
            raise ValueError("FusedSGD needs one momentum value for all groups")
        self._momentum = float(next(iter(moms)))
        self._table = None
        self._table_key = None

    def _entries(self):
        ent = []
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if "momentum_buffer" not in st or st["momentum_buffer"] is None:
                    st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                ent.append((p, p.grad, st["momentum_buffer"], float(g["lr"]), float(g["weight_decay"])))
        return ent

    def _ensure_table(self):
        ent = self._entries()
        key = tuple((p.data_ptr(), g.data_ptr(), b.data_ptr(), lr, wd) for p, g, b, lr, wd in ent)
        if key != self._table_key:
            if not ent:
                self._table = None
            else:
                dev = ent[0][0].device
                L.require_device(dev)
                self._table = ops.sgd_table([(p.detach(), g, b, lr, wd) for p, g, b, lr, wd in ent], dev) + (len(ent),)
            self._table_key = key
        return ent

    @torch.no_grad()
    def step(self, closure=None, zero_grad: bool = False):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ent = self._ensure_table()
        if self._table is None:
            return loss
        table, prefix, n_chunks, n_tensors = self._table
        ops.sgd_step(table, prefix, n_tensors, n_chunks, self._momentum, zero_grad)
        for p, _, _, lr, _ in ent:
            if lr != 0.0:
                torch.autograd.graph.increment_version(p)   # the Parameter's own counter: packed copies get rebuilt
        return loss

    def step_and_zero(self):
        """optimizer.step(); optimizer.zero_grad() of train_online.py:99-100 in one launch
        (gradients are cleared in place, not set to None, so their addresses stay stable)."""
        return self.step(zero_grad=True)


def _groups(net, mode: str, learning_rate: float, weight_decay: float) -> List[dict]:
    lr, wd = learning_rate, weight_decay
    groups = [
        {'params': [pr[1] for pr in net.stages.named_parameters() if 'weight' in pr[0]], 'weight_decay': wd},
        {'params': [pr[1] for pr in net.stages.named_parameters() if 'bias' in pr[0]], 'lr': lr * 2},
        {'params': [pr[1] for pr in net.side_prep.named_parameters() if 'weight' in pr[0]], 'weight_decay': wd},
        {'params': [pr[1] for pr in net.side_prep.named_parameters() if 'bias' in pr[0]], 'lr': lr * 2},
    ]
    if mode == "offline":
        groups += [
            {'params': [pr[1] for pr in net.score_dsn.named_parameters() if 'weight' in pr[0]], 'lr': lr / 10, 'weight_decay': wd},
            {'params': [pr[1] for pr in net.score_dsn.named_parameters() if 'bias' in pr[0]], 'lr': 2 * lr / 10},
        ]
    groups += [
        {'params': [pr[1] for pr in net.upscale.named_parameters() if 'weight' in pr[0]], 'lr': 0},
        {'params': [pr[1] for pr in net.upscale_.named_parameters() if 'weight' in pr[0]], 'lr': 0},
        {'params': net.fuse.weight, 'lr': lr / 100, 'weight_decay': wd},
        {'params': net.fuse.bias, 'lr': 2 * lr / 100},
    ]
    return [g for g in groups if not isinstance(g['params'], list) or len(g['params']) > 0]


def get_optimizer_online(net, learning_rate: float = 1e-8, weight_decay: float = 0.0002, momentum: float = 0.9,
                         fused: bool = True):
    """Param groups of VGGOnlineProvider.get_optimizer (network_provider.py:144-159)."""
    cls = FusedSGD if fused else torch.optim.SGD
    return cls(_groups(net, "online", learning_rate, weight_decay), lr=learning_rate, momentum=momentum)


def get_optimizer_offline(net, learning_rate: float = 1e-8, weight_decay: float = 0.0002, momentum: float = 0.9,
                          fused: bool = True):
    """Param groups of VGGOfflineProvider.get_optimizer (network_provider.py:98-125)."""
    cls = FusedSGD if fused else torch.optim.SGD
    return cls(_groups(net, "offline", learning_rate, weight_decay), lr=learning_rate, momentum=momentum)
