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
        self._table_gen = 0          # bumps when the device table is REPLACED (captured graphs holding its address go stale)
        self._exclude = set()        # ids of parameters stepped elsewhere (the trainer's fused conv-weight step)

    def exclude(self, params) -> None:
        """Leave these parameters to another kernel (``OnlineTrainer``'s fused fold + SGD + repack of the conv weights); their
        momentum buffers still live in ``self.state`` so ``state_dict()`` keeps the torch.optim.SGD layout."""
        self._exclude = {id(p) for p in params}
        self._table_key = None

    def group_options(self, p):
        """(lr, weight_decay) of the group holding parameter ``p``."""
        for g in self.param_groups:
            if any(q is p for q in g["params"]):
                return float(g["lr"]), float(g["weight_decay"])
        raise KeyError("parameter is in no param group")

    def momentum_buffer(self, p) -> torch.Tensor:
        st = self.state[p]
        if "momentum_buffer" not in st or st["momentum_buffer"] is None:
            st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        return st["momentum_buffer"]

    def _entries(self):
        ent = []
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None or id(p) in self._exclude:
                    continue
                st = self.state[p]
                if "momentum_buffer" not in st or st["momentum_buffer"] is None:
                    st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                ent.append((p, p.grad, st["momentum_buffer"], float(g["lr"]), float(g["weight_decay"])))
        return ent

    def _ensure_table(self):
        """(Re)build the device table {param, grad, momentum buffer, lr, weight decay} when any of them changed.  A table
        of unchanged geometry is rewritten IN PLACE (same device buffers): a captured CUDA graph of the step reads it at
        replay, so new ``param_groups`` learning rates / weight decays and new momentum buffers (``load_state_dict``) take
        effect without a re-capture; a different tensor count or size replaces the table and bumps ``_table_gen``."""
        moms = {g["momentum"] for g in self.param_groups}
        if len(moms) != 1:
            raise ValueError("FusedSGD needs one momentum value for all groups")
        if float(next(iter(moms))) != self._momentum:
            self._momentum = float(next(iter(moms)))
            self._table_gen += 1                         # the momentum is a kernel argument: captured steps must be re-captured
        ent = self._entries()
        key = tuple((p.data_ptr(), g.data_ptr(), b.data_ptr(), lr, wd) for p, g, b, lr, wd in ent)
        if key != self._table_key:
            if not ent:
                self._table = None
                self._table_gen += 1
            else:
                dev = ent[0][0].device
                L.require_device(dev)
                new = ops.sgd_table([(p.detach(), g, b, lr, wd) for p, g, b, lr, wd in ent], dev) + (len(ent),)
                old = self._table
                if old is not None and old[0].numel() == new[0].numel() and old[2] == new[2] and old[3] == new[3] and \
                        old[0].device == new[0].device and torch.equal(old[1], new[1]):
                    old[0].copy_(new[0])
                else:
                    self._table = new
                    self._table_gen += 1
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


class FusedAdam(Optimizer):
    """torch.optim.Adam (no amsgrad; L2 weight decay) as ONE multi-tensor launch -- the optimizer of the reference's
    distillation and post-pruning fine-tuning loops (``mimic.py:74`` lr 1e-3, ``prune.py`` fine_tune lr 1e-4, wd 2e-4).
    State keys match torch (``exp_avg``, ``exp_avg_sq``); the step counter lives on the device so a captured
    graph can replay the step."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len({(g["betas"], g["eps"]) for g in self.param_groups}) != 1:
            raise ValueError("FusedAdam needs one (betas, eps) for all groups")
        self._table = None
        self._table_key = None
        self._step_dev = None

    def _entries(self):
        ent = []
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                ent.append((p, p.grad, st["exp_avg"], st["exp_avg_sq"], float(g["lr"]), float(g["weight_decay"])))
        return ent

    def _ensure_table(self):
        ent = self._entries()
        key = tuple((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), lr, wd) for p, g, m, v, lr, wd in ent)
        if key != self._table_key:
            if ent:
                dev = ent[0][0].device
                L.require_device(dev)
                self._table = ops.adam_table([(p.detach(), g, m, v, lr, wd) for p, g, m, v, lr, wd in ent], dev) + (len(ent),)
                if self._step_dev is None:
                    self._step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            else:
                self._table = None
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
        g0 = self.param_groups[0]
        ops.adam_step(table, prefix, n_tensors, n_chunks, g0["betas"][0], g0["betas"][1], g0["eps"], self._step_dev, zero_grad)
        for p, *_ in ent:
            torch.autograd.graph.increment_version(p)
        return loss

    def step_and_zero(self):
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
    if mode == "offline":
        # the offline provider also records every group's starting rate (network_provider.py:98-124): schedulers resumed
        # with last_epoch >= 0 read it
        for g in groups:
            g['initial_lr'] = g.get('lr', lr)
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
