"""Channel pruning of the OSVOS VGG-16 network: the reference's ``FilterPruner`` algorithm
(``/root/reference/src/prune.py:69-223``) and ``prune_convolution`` surgery (``:490-514``) on the VGG topology.

The reference implements pruning for ``OSVOS_RESNET`` only; the algorithm is topology-independent
(Taylor criterion per output filter, per-layer L2 normalisation of the ranks, global ``nsmallest``,
filter indices re-based as earlier filters of the same layer disappear), and the VGG topology is the
simple case: conv ``k``'s output filters feed conv ``k+1``'s input channels (a max-pool in between keeps
channels), and the last conv of stages 1..4 also feeds ``side_prep``.

Ranks are NOT gathered with autograd hooks here: the network's backward accumulates
``sum(activation * gradient) / (N H W)`` per filter in one fused pass (``OSVOS_VGG._run_backward``).
"""
from __future__ import annotations

import heapq
import operator
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .networks import OSVOS_VGG

N_MIN_CHANNELS = 4          # reference prune.py:30


def stage_conv_index(net: OSVOS_VGG) -> List[Tuple[int, int]]:
    """Flat conv index -> (stage, module index inside the stage Sequential)."""
    return [(si, mi) for si, st in enumerate(net.stages) for mi, m in enumerate(st) if isinstance(m, nn.Conv2d)]


class FilterPruner:
    """Same bookkeeping and method names as the reference class (``prune.py:69-223``)."""

    def __init__(self, net: OSVOS_VGG):
        self.net = net
        self.filter_ranks: Dict[int, torch.Tensor] = {}
        self.activation_to_layer: Dict[int, int] = {}
        self.skip_layer: List[int] = []
        self.reset()

    def reset(self) -> None:
        self.filter_ranks = {}

    def forward_backward(self, x: torch.Tensor, gts: torch.Tensor, is_offline: bool = False) -> torch.Tensor:
        """One ``train_for_pruning`` minibatch (``prune.py:229-248``): forward, class-balanced loss on the fused
        map (online) or on all five maps (offline), backward -- accumulating the Taylor ranks of every stage
        conv.  Parameter gradients are discarded (the reference zeroes them before every minibatch)."""
        net = self.net
        idx = stage_conv_index(net)
        dev = next(net.parameters()).device
        names = [f"stages.{si}.{mi}" for si, mi in idx]
        self.activation_to_layer = {k: k for k in range(len(idx))}
        self.skip_layer = [k for k, (si, mi) in enumerate(idx) if net.stages[si][mi].out_channels <= N_MIN_CHANNELS]
        taylor = {}
        for k, ((si, mi), name) in enumerate(zip(idx, names)):
            if k not in self.filter_ranks:
                self.filter_ranks[k] = torch.zeros(net.stages[si][mi].out_channels, dtype=torch.float32, device=dev)
            taylor[name] = self.filter_ranks[k]
        with torch.no_grad():
            keep, net._keep_activations = net._keep_activations, True      # the ranks read every conv's output, conv1_2's included
            try:
                outs, _, _, saved = net._run_forward(x, save=True)
            finally:
                net._keep_activations = keep
            douts: List[Optional[torch.Tensor]] = [None] * 5
            total = None
            for i in (range(5) if is_offline else [4]):
                loss, stats = ops.bal_loss_fwd(outs[i], gts, False)
                douts[i] = ops.bal_loss_bwd(outs[i], gts, False, stats, None, 1.0)
                total = loss if total is None else total + loss
            params = dict(net.named_parameters())
            grads = {n: torch.zeros_like(params[n], memory_format=torch.contiguous_format) for n in net._grad_names()}
            net._run_backward(saved, douts, grads, taylor=taylor)
        return total

    def normalize_ranks_per_layer(self) -> None:
        # prune.py:180-188
        for i in self.filter_ranks:
            v = torch.abs(self.filter_ranks[i])
            divisor = float(torch.sqrt(torch.sum(v * v)))
            if divisor >= 1e-5:
                v = v / divisor
            self.filter_ranks[i] = v.cpu()

    def lowest_ranking_filters(self, n_filters_to_prune_per_iter: int):
        # prune.py:190-201
        data = []
        for i in sorted(self.filter_ranks.keys()):
            index_layer = self.activation_to_layer[i]
            if index_layer in self.skip_layer:
                continue
            for j in range(self.filter_ranks[i].size(0)):
                data.append((index_layer, j, float(self.filter_ranks[i][j])))
        return heapq.nsmallest(n_filters_to_prune_per_iter, data, operator.itemgetter(2))

    def get_prunning_plan(self, n_filters_to_prune_per_iter: int) -> List[Tuple[int, int]]:
        # prune.py:203-223: indices shift down as earlier filters of the same layer are removed
        filters_to_prune = self.lowest_ranking_filters(n_filters_to_prune_per_iter)
        per_layer: Dict[int, List[int]] = {}
        for (l, f, _) in filters_to_prune:
            per_layer.setdefault(l, []).append(f)
        for l in per_layer:
            per_layer[l] = sorted(per_layer[l])
            for i in range(len(per_layer[l])):
                per_layer[l][i] = per_layer[l][i] - i
        return [(l, i) for l in per_layer for i in per_layer[l]]


def prune_convolution(conv: nn.Conv2d, filter_index: int, is_reducing_channels_out: bool, keep_bias: bool = True) -> nn.Conv2d:
    """Remove one output filter or one input channel (reference ``prune.py:490-514``).  The reference rebuilds every conv
    it touches with ``bias=False`` -- its ResNet convs are bias-free and followed by BatchNorm -- which on the VGG topology
    would delete the trained bias vectors of the pruned conv, of its consumer and of ``side_prep`` for ONE removed filter.
    So the bias is kept by default here (sliced on out-channel pruning, intact on in-channel pruning);
    ``keep_bias=False`` is the literal reference surgery, used for the BASELINE configs[2] fixture."""
    d_in, d_out = (0, 1) if is_reducing_channels_out else (1, 0)
    has_bias = keep_bias and conv.bias is not None
    new_conv = nn.Conv2d(conv.in_channels - d_in, conv.out_channels - d_out, kernel_size=conv.kernel_size, stride=conv.stride,
                         padding=conv.padding, dilation=conv.dilation, groups=conv.groups, bias=has_bias)
    w = conv.weight.data
    keep = [i for i in range(w.shape[0 if is_reducing_channels_out else 1]) if i != filter_index]
    idx = torch.tensor(keep, dtype=torch.long, device=w.device)
    new_conv.weight.data = w.index_select(0 if is_reducing_channels_out else 1, idx).contiguous()
    if has_bias:
        new_conv.bias.data = conv.bias.data.index_select(0, idx).contiguous() if is_reducing_channels_out else conv.bias.data.clone()
    return new_conv.to(w.device)


def prune_vgg_conv_layer(net: OSVOS_VGG, layer_index: int, filter_index: int, keep_bias: bool = True) -> OSVOS_VGG:
    """Remove output filter ``filter_index`` of stage conv ``layer_index`` and the matching input channel of its
    consumers (next conv; ``side_prep`` when it is the last conv of stages 1..4)."""
    idx = stage_conv_index(net)
    si, mi = idx[layer_index]
    if net.stages[si][mi].out_channels <= N_MIN_CHANNELS:
        return net
    net.stages[si][mi] = prune_convolution(net.stages[si][mi], filter_index, True, keep_bias)
    if layer_index + 1 < len(idx):
        sj, mj = idx[layer_index + 1]
        net.stages[sj][mj] = prune_convolution(net.stages[sj][mj], filter_index, False, keep_bias)
    last_of_stage = layer_index + 1 == len(idx) or idx[layer_index + 1][0] != si
    if last_of_stage and si > 0:
        net.side_prep[si - 1] = prune_convolution(net.side_prep[si - 1], filter_index, False, keep_bias)
    net.invalidate_packed()
    if hasattr(net, "_trainers"):
        net.__dict__["_trainers"].clear()
    net.__dict__.pop("_repack_table", None)
    return net


def prune_step(net: OSVOS_VGG, frames: Sequence[torch.Tensor], gts: Sequence[torch.Tensor], n_filters: int,
               is_offline: bool = False, keep_bias: bool = True) -> List[Tuple[int, int]]:
    """One pruning iteration of the reference driver (``prune.py`` main loop): rank on the given minibatches,
    normalise, plan, apply.  Returns the applied ``(layer, filter)`` plan."""
    pruner = FilterPruner(net)
    for x, g in zip(frames, gts):
        pruner.forward_backward(x, g, is_offline)
    pruner.normalize_ranks_per_layer()
    plan = pruner.get_prunning_plan(n_filters)
    for layer_index, filter_index in plan:
        prune_vgg_conv_layer(net, layer_index, filter_index, keep_bias)
    return plan


def l2_prune_half(net: OSVOS_VGG, keep_fraction: float = 0.5, keep_bias: bool = False) -> OSVOS_VGG:
    """Deterministic stand-in used by BASELINE configs[2]: keep the ceil(fraction * C) filters of
    every stage conv with the largest L2 norm.  ``keep_bias=False`` (the default HERE only) reproduces the reference's
    literal surgery -- every touched conv, ``side_prep`` included, comes back with ``bias=False`` (prune.py:500) -- which is
    what SURVEY 8d specifies for this configuration."""
    idx = stage_conv_index(net)
    for layer_index, (si, mi) in enumerate(idx):
        conv = net.stages[si][mi]
        keep = max(N_MIN_CHANNELS, int(np.ceil(keep_fraction * conv.out_channels)))
        norms = conv.weight.data.flatten(1).norm(dim=1)
        drop = sorted(torch.argsort(norms)[: conv.out_channels - keep].tolist())
        for j, f in enumerate(drop):
            prune_vgg_conv_layer(net, layer_index, f - j, keep_bias)
    return net
