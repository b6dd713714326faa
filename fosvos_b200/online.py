"""One-shot online fine-tuning and per-sequence inference on one GPU.

Mirrors the hot loops of the reference drivers, with the same argument meaning:
  ``finetune``      <- ``train_online._train`` loop body (``src/train_online.py:75-101``)
  ``infer_sequence``<- ``util/experiment_helper.test`` (``src/util/experiment_helper.py:34-64``)
  ``sequences_for_rank`` <- the ``--sequence-group`` round-robin (``src/train_online.py:184-186``)

Differences that do not change results: the frame/mask stay resident on the device (the
reference re-uploads the same first frame every iteration, ``train_online.py:77``), the loss
scalar stays on the device (the reference syncs every iteration, ``:82``), the optimizer step
and ``zero_grad`` are one launch, and the whole micro-iteration can be replayed from a CUDA
graph.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import ops
from .networks import OSVOS_VGG
from .optim import FusedSGD, get_optimizer_online


class _MicroStep:
    """fwd -> balanced loss (fused map, size_average=False) -> /n -> bwd into p.grad (+=)."""

    def __init__(self, net: OSVOS_VGG, frame: torch.Tensor, mask: torch.Tensor, avg_grad_every_n: int,
                 deep_supervision: Optional[float] = None):
        self.net, self.frame, self.mask = net, frame, mask
        self.scale = 1.0 / float(avg_grad_every_n)
        self.deep = deep_supervision      # None: online (fused map only); w: offline weight on the 4 side maps
        self.names = net._grad_names()
        params = dict(net.named_parameters())
        self.grads: Dict[str, torch.Tensor] = {}
        for n in self.names:
            p = params[n]
            if p.grad is None:
                p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
            self.grads[n] = p.grad
        self.loss_sum = torch.zeros((), dtype=torch.float32, device=frame.device)
        self.last_loss = torch.zeros((), dtype=torch.float32, device=frame.device)

    def run(self) -> None:
        outs, _, _, saved = self.net._run_forward(self.frame, save=True)
        douts: List[Optional[torch.Tensor]] = [None] * 5
        loss, stats = ops.bal_loss_fwd(outs[4], self.mask, False)
        douts[4] = ops.bal_loss_bwd(outs[4], self.mask, False, stats, None, self.scale)
        total = loss
        if self.deep is not None:
            for i in range(4):
                li, st = ops.bal_loss_fwd(outs[i], self.mask, False)
                douts[i] = ops.bal_loss_bwd(outs[i], self.mask, False, st, None, self.scale * self.deep)
                total = total + self.deep * li
        self.last_loss.copy_(total)
        self.loss_sum.add_(total)
        self.net._run_backward(saved, douts, self.grads)


def finetune(net: OSVOS_VGG, frame: torch.Tensor, mask: torch.Tensor, n_iters: int, avg_grad_every_n: int = 5,
             optimizer: Optional[FusedSGD] = None, use_graph: bool = False,
             losses_out: Optional[list] = None) -> torch.Tensor:
    """Fine-tune ``net`` on one annotated frame (``train_online.py:70-101`` with a 1-sample loader).

    frame (1,3,H,W) fp32, mask (1,1,H,W) fp32, both on the device.  Returns the device scalar of the
    summed per-iteration losses (no host sync inside the loop unless ``losses_out`` is given)."""
    L.require_device(frame.device)
    if optimizer is None:
        optimizer = get_optimizer_online(net)
    frame = frame.contiguous().float()
    mask = mask.contiguous().float()
    micro = _MicroStep(net, frame, mask, avg_grad_every_n)
    graph = None
    if use_graph:
        # warm up once outside capture (packs weights, sizes the allocator), then undo its gradient
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            calls0 = L.CALLS[0]
            micro.run()
            calls_per_replay = L.CALLS[0] - calls0
            for g in micro.grads.values():
                g.zero_()
            micro.loss_sum.zero_()
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            micro.run()
        for g in micro.grads.values():
            g.zero_()
        micro.loss_sum.zero_()
    counter = 0
    for _ in range(n_iters):
        if graph is not None:
            graph.replay()
            L.CALLS[0] += calls_per_replay
        else:
            micro.run()
        if losses_out is not None:
            losses_out.append(float(micro.last_loss.item()))
        counter += 1
        if counter % avg_grad_every_n == 0:
            optimizer.step_and_zero()
            if graph is not None:
                _repack_in_place(net)
            counter = 0
    return micro.loss_sum


def _repack_in_place(net: OSVOS_VGG) -> None:
    """Refresh the packed weight copies INTO their existing buffers (addresses are baked into a
    captured graph) after an optimizer step."""
    from .networks import _act_dtype
    dt = _act_dtype(net.precision)
    tc = net._impl() == "tc"
    convs = [c for st in net._stage_convs() for c in st] + list(net.side_prep)
    for conv in convs:
        pc = net._packed.get(id(conv))
        if pc is None:
            continue
        b = conv.bias
        if pc.w_fwd is not None:
            ops.pack_weight(conv.weight, L.W_TC_FWD if tc else L.W_SIMT_FWD, dt, out=pc.w_fwd)
            ops.pad_bias(b, conv.out_channels, conv.weight.device, out=pc.bias)
        if pc.w_dgrad is not None:
            ops.pack_weight(conv.weight, L.W_TC_DGRAD if tc else L.W_SIMT_DGRAD, dt, out=pc.w_dgrad)
        pc.key = (conv.weight.data_ptr(), conv.weight._version, None if b is None else (b.data_ptr(), b._version),
                  net.precision, tuple(conv.weight.shape))
    # fuse / score heads feed the side-chain parameter block
    ops.side_params_prepare([m.weight for m in net.upscale], [m.weight for m in net.upscale_],
                            [m.weight for m in net.score_dsn], [m.bias for m in net.score_dsn],
                            net.fuse.weight, net.fuse.bias, out=net._side_params)
    ps = [m.weight for m in net.upscale] + [m.weight for m in net.upscale_] + \
         [m.weight for m in net.score_dsn] + [m.bias for m in net.score_dsn] + [net.fuse.weight, net.fuse.bias]
    net._side_key = tuple((p.data_ptr(), p._version) for p in ps)


@torch.no_grad()
def infer_sequence(net: OSVOS_VGG, frames: torch.Tensor, batch_size: int = 1, want_prob: bool = False):
    """Segment every frame of a sequence (``experiment_helper.test``'s loop): returns uint8 masks
    (F,1,H,W) [and fp32 probabilities] on the device; sigmoid + threshold are fused into the
    side-chain kernel instead of a host-side numpy pass (``experiment_helper.py:55-57``)."""
    L.require_device(frames.device)
    masks, probs = [], []
    for i in range(0, frames.shape[0], batch_size):
        _, prob, mask = net.predict(frames[i:i + batch_size])
        masks.append(mask)
        if want_prob:
            probs.append(prob)
    masks = torch.cat(masks)
    return (masks, torch.cat(probs)) if want_prob else masks


def sequences_for_rank(sequences: Sequence, rank: int, world_size: int) -> list:
    """``[s for i, s in enumerate(sequences) if i % group_size == group]`` -- the reference's manual
    job sharding (``train_online.py:184-186``); one process per GPU, no collective."""
    return [s for i, s in enumerate(sequences) if i % world_size == rank]


def region_iou(pred_masks: torch.Tensor, gt_masks: torch.Tensor) -> torch.Tensor:
    """(F,2) int64 intersection/union counts per frame: the integers behind DAVIS J."""
    return ops.mask_iou(pred_masks.reshape(pred_masks.shape[0], -1), gt_masks.reshape(gt_masks.shape[0], -1))
