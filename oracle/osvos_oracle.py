"""CPU oracle for the OSVOS VGG-16 hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement, on the CPU, of the arithmetic the
reference (klausondrag/FOSVOS) performs on its per-frame hot path.  It exists
so that the CUDA path in ``fosvos_b200`` can be checked bit-for-tolerance on
a box where ``/root/reference`` does not exist.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package never does.

Parity pin
----------
The reference holds no tests or golden vectors for this path,
and its convolution arithmetic lives in a third-party dependency that is not
vendored under ``/root/reference``: PyTorch (``torch==0.4.0`` per the
reference ``README.md:11``; unpinned torchvision).  The pin therefore is the
*live* reference classes executed in the build container under torch 2.11:
``oracle/make_golden.py`` imports ``networks.osvos_vgg.OSVOS_VGG``,
``layers.osvos_layers.*`` and ``util.network_provider.VGGOnlineProvider``
from ``/root/reference/src``, checks every function below against them and
writes the fixtures in ``tests/golden/``; ``tests/test_oracle_golden.py``
re-checks this file against those fixtures everywhere.

Each function cites the reference file:line it follows.  The third-party
convolution / pooling / transposed-convolution arithmetic is expressed through
``torch.nn.functional`` on CPU tensors (the same ATen code the reference
itself dispatches to); ``oracle/osvos_numpy.py`` restates those primitives
once more as plain numpy loops for small cases.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# (cfg, in_channels) of the five VGG stages -- reference osvos_vgg.py:20-25
LAY_LIST = [[64, 64], ["M", 128, 128], ["M", 256, 256, 256], ["M", 512, 512, 512], ["M", 512, 512, 512]]
IN_CHANNELS = [3, 64, 128, 256, 512]


# --------------------------------------------------------------------------
# layers/osvos_layers.py
# --------------------------------------------------------------------------
def upsample_filt(size: int) -> np.ndarray:
    """Bilinear kernel, float64 (reference osvos_layers.py:57-65)."""
    factor = (size + 1) // 2
    center = factor - 1 if size % 2 == 1 else factor - 0.5
    og = np.ogrid[:size, :size]
    return (1 - abs(og[0] - center) / factor) * (1 - abs(og[1] - center) / factor)


def interp_surgery_weight(channels: int, k: int) -> torch.Tensor:
    """Weight a ConvTranspose2d(channels, channels, k) gets from interp_surgery:
    zeros with the bilinear kernel on the channel diagonal
    (reference osvos_layers.py:70-81 after osvos_vgg.py:109-111)."""
    w = torch.zeros(channels, channels, k, k, dtype=torch.float32)
    filt = torch.from_numpy(upsample_filt(k)).float()
    for i in range(channels):
        w[i, i] = filt
    return w


def center_crop_pads(in_h: int, in_w: int, height: int, width: int) -> Tuple[int, int, int, int]:
    """(left, right, top, bottom) pads handed to F.pad (negative = crop).
    left/top = ceil(-(d)/2), right/bottom = floor(-(d)/2), d = in - target
    (reference osvos_layers.py:47-54; float32 arithmetic there, exact here)."""
    ch = -(in_h - height) / 2.0
    cw = -(in_w - width) / 2.0
    return (int(math.ceil(cw)), int(math.floor(cw)), int(math.ceil(ch)), int(math.floor(ch)))


def center_crop(x: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """reference osvos_layers.py:47-54."""
    l, r, t, b = center_crop_pads(x.shape[2], x.shape[3], height, width)
    return F.pad(x, [l, r, t, b])


def class_balanced_cross_entropy_loss(output: torch.Tensor, label: torch.Tensor,
                                      size_average: bool = True) -> torch.Tensor:
    """reference osvos_layers.py:17-44, op for op."""
    labels = torch.ge(label, 0.5).to(output.dtype)
    num_labels_pos = torch.sum(labels)
    num_labels_neg = torch.sum(1.0 - labels)
    num_total = num_labels_pos + num_labels_neg
    output_gt_zero = torch.ge(output, 0).to(output.dtype)
    loss_val = torch.mul(output, (labels - output_gt_zero)) - torch.log(
        1 + torch.exp(output - 2 * torch.mul(output, output_gt_zero)))
    loss_pos = torch.sum(-torch.mul(labels, loss_val))
    loss_neg = torch.sum(-torch.mul(1.0 - labels, loss_val))
    final_loss = num_labels_neg / num_total * loss_pos + num_labels_pos / num_total * loss_neg
    if size_average:
        final_loss = final_loss / (label.size(0) * label.size(1) * label.size(2) * label.size(3))
    return final_loss


def class_balanced_cross_entropy_grad(output: torch.Tensor, label: torch.Tensor,
                                      size_average: bool = True, grad_out: float = 1.0) -> torch.Tensor:
    """Closed form of what autograd yields for the loss above:
    dL/dx = w(y) * (sigmoid(x) - y) * g,  w(1) = neg/total, w(0) = pos/total."""
    y = torch.ge(label, 0.5).to(output.dtype)
    pos = y.sum()
    total = torch.tensor(float(y.numel()), dtype=output.dtype)
    neg = total - pos
    w = y * (neg / total) + (1 - y) * (pos / total)
    g = w * (torch.sigmoid(output) - y) * grad_out
    if size_average:
        g = g / y.numel()
    return g


# --------------------------------------------------------------------------
# networks/osvos_vgg.py
# --------------------------------------------------------------------------
def stage_conv_indices() -> List[List[int]]:
    """Indices of the Conv2d modules inside each ``stages[i]`` Sequential
    (reference osvos_vgg.py:85-95): [0,2] / [1,3] / [1,3,5] x3."""
    out = []
    for cfg in LAY_LIST:
        idx, k = [], 0
        for v in cfg:
            if v == "M":
                k += 1
            else:
                idx.append(k)
                k += 2
        out.append(idx)
    return out


def state_dict_spec(channels: Optional[Sequence[Sequence[int]]] = None,
                    stage_bias: bool = True) -> List[Tuple[str, Tuple[int, ...]]]:
    """Ordered (key, shape) list of ``OSVOS_VGG.state_dict()``: attribute
    assignment order upscale, upscale_, stages, side_prep, score_dsn, fuse
    (reference osvos_vgg.py:50-56).  ``channels`` overrides the per-stage
    output widths (channel-pruned variants, reference prune.py:490-514 rebuilds
    convs with bias=False -> ``stage_bias=False``)."""
    widths = [[v for v in cfg if v != "M"] for cfg in LAY_LIST] if channels is None else [list(c) for c in channels]
    spec: List[Tuple[str, Tuple[int, ...]]] = []
    for i in range(4):
        k = 2 ** (2 + i)
        spec.append((f"upscale.{i}.weight", (16, 16, k, k)))
    for i in range(4):
        k = 2 ** (2 + i)
        spec.append((f"upscale_.{i}.weight", (1, 1, k, k)))
    cin = 3
    for si, idxs in enumerate(stage_conv_indices()):
        for j, mi in enumerate(idxs):
            cout = widths[si][j]
            spec.append((f"stages.{si}.{mi}.weight", (cout, cin, 3, 3)))
            if stage_bias:
                spec.append((f"stages.{si}.{mi}.bias", (cout,)))
            cin = cout
    for i in range(4):
        spec.append((f"side_prep.{i}.weight", (16, widths[i + 1][-1], 3, 3)))
        spec.append((f"side_prep.{i}.bias", (16,)))
    for i in range(4):
        spec.append((f"score_dsn.{i}.weight", (1, 16, 1, 1)))
        spec.append((f"score_dsn.{i}.bias", (1,)))
    spec.append(("fuse.weight", (1, 64, 1, 1)))
    spec.append(("fuse.bias", (1,)))
    return spec


def init_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """``pretrained=0`` initialisation: Conv2d W ~ N(0, 1e-3^2), b = 0; ConvT
    zero + diagonal bilinear (reference osvos_vgg.py:97-111).  (Same
    distribution, not the same random stream as the reference.)"""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape in state_dict_spec():
        if key.startswith("upscale"):
            sd[key] = interp_surgery_weight(shape[0], shape[2])
        elif key.endswith("weight"):
            sd[key] = torch.randn(shape, generator=g) * 0.001
        else:
            sd[key] = torch.zeros(shape)
    return sd


def vgg_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor,
                return_intermediates: bool = False):
    """``OSVOS_VGG.forward`` (reference osvos_vgg.py:61-83) as a function of a
    reference-layout state_dict.  Works for any per-layer widths and for
    bias-less stage convs (pruned variants).  Returns [side0..side3, fused]."""
    dt = x.dtype
    P = lambda k: sd[k].to(dt) if k in sd else None  # noqa: E731
    crop_h, crop_w = int(x.shape[-2]), int(x.shape[-1])
    idxs = stage_conv_indices()
    inter: Dict[str, torch.Tensor] = {}
    side, side_out = [], []
    for si in range(5):
        if si > 0:
            x = F.max_pool2d(x, kernel_size=2, stride=2, ceil_mode=True)          # osvos_vgg.py:90
        for mi in idxs[si]:
            x = F.relu(F.conv2d(x, P(f"stages.{si}.{mi}.weight"), P(f"stages.{si}.{mi}.bias"), padding=1))  # :92-93
            inter[f"stages.{si}.{mi}"] = x
        if si > 0:
            i = si - 1
            side_temp = F.conv2d(x, P(f"side_prep.{i}.weight"), P(f"side_prep.{i}.bias"), padding=1)       # :69
            inter[f"side_prep.{i}"] = side_temp
            s = 2 ** si
            up = F.conv_transpose2d(side_temp, P(f"upscale.{i}.weight"), stride=s)                        # :71
            side.append(center_crop(up, crop_h, crop_w))                                                   # :72-73
            score = F.conv2d(side_temp, P(f"score_dsn.{i}.weight"), P(f"score_dsn.{i}.bias"))              # :75
            up_ = F.conv_transpose2d(score, P(f"upscale_.{i}.weight"), stride=s)                          # :76
            side_out.append(center_crop(up_, crop_h, crop_w))                                              # :77-78
    out = torch.cat(side, dim=1)                                                                           # :80
    out = F.conv2d(out, P("fuse.weight"), P("fuse.bias"))                                                  # :81
    side_out.append(out)
    if return_intermediates:
        return side_out, inter
    return side_out


# --------------------------------------------------------------------------
# util/network_provider.py : optimizer groups; torch.optim.SGD arithmetic
# --------------------------------------------------------------------------
def optimizer_groups(keys: Sequence[str], mode: str = "online", learning_rate: float = 1e-8,
                     weight_decay: float = 0.0002) -> List[dict]:
    """(keys, lr, weight_decay) per param group.
    online : reference network_provider.py:144-159 (score_dsn is in NO group)
    offline: reference network_provider.py:98-125."""
    lr, wd = learning_rate, weight_decay
    sel = lambda prefix, kind: [k for k in keys if k.startswith(prefix + ".") and kind in k]  # noqa: E731
    groups = [
        dict(keys=sel("stages", "weight"), lr=lr, weight_decay=wd),
        dict(keys=sel("stages", "bias"), lr=2 * lr, weight_decay=0.0),
        dict(keys=sel("side_prep", "weight"), lr=lr, weight_decay=wd),
        dict(keys=sel("side_prep", "bias"), lr=2 * lr, weight_decay=0.0),
    ]
    if mode == "offline":
        groups += [
            dict(keys=sel("score_dsn", "weight"), lr=lr / 10, weight_decay=wd),
            dict(keys=sel("score_dsn", "bias"), lr=2 * lr / 10, weight_decay=0.0),
        ]
    groups += [
        dict(keys=sel("upscale", "weight"), lr=0.0, weight_decay=0.0),
        dict(keys=sel("upscale_", "weight"), lr=0.0, weight_decay=0.0),
        dict(keys=["fuse.weight"], lr=lr / 100, weight_decay=wd),
        dict(keys=["fuse.bias"], lr=2 * lr / 100, weight_decay=0.0),
    ]
    return groups


def sgd_step(p: torch.Tensor, g: torch.Tensor, buf: Optional[torch.Tensor], lr: float,
             weight_decay: float, momentum: float = 0.9) -> Tuple[torch.Tensor, torch.Tensor]:
    """One torch.optim.SGD update (dampening 0, no Nesterov), as the
    reference's optimizer performs it: d = g + wd*p; buf = d on the first step
    else mu*buf + d; p -= lr*buf.  Returns (new_p, new_buf)."""
    d = g + weight_decay * p if weight_decay != 0 else g.clone()
    buf = d.clone() if buf is None else momentum * buf + d
    return p - lr * buf, buf


def finetune(sd: Dict[str, torch.Tensor], frame: torch.Tensor, mask: torch.Tensor, n_iters: int,
             avg_grad_every_n: int = 5, learning_rate: float = 1e-8, weight_decay: float = 0.0002,
             momentum: float = 0.9, mode: str = "online", epoch_frac: float = 0.0):
    """The loop body of reference train_online.py:75-101 (``.item()`` in place
    of ``.data[0]``): fwd -> loss on the fused map (size_average=False) ->
    /avg_grad_every_n -> backward -> every n: SGD step + zero_grad.
    ``mode='offline'`` instead follows reference train_offline.py:84-88,102-110
    (five losses, deep-supervision weight (1 - epoch/n_epochs) = 1-epoch_frac).
    ``frame`` / ``mask`` may be lists with one tensor per iteration (cycled): the train-time augmentation of the
    reference loader (random flip and Resize scale per sample, dataloaders/custom_transforms.py:63-109) then
    shows up as frames of different sizes from one iteration to the next.
    Returns (new_state_dict, [loss per iteration before the /n])."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    groups = optimizer_groups(list(params.keys()), mode, learning_rate, weight_decay)
    bufs: Dict[str, Optional[torch.Tensor]] = {k: None for k in params}
    losses, counter = [], 0
    for it in range(n_iters):
        fr = frame[it % len(frame)] if isinstance(frame, (list, tuple)) else frame
        mk = mask[it % len(mask)] if isinstance(mask, (list, tuple)) else mask
        outs = vgg_forward(params, fr)
        if mode == "online":
            loss = class_balanced_cross_entropy_loss(outs[-1], mk, size_average=False)
        else:
            ls = [class_balanced_cross_entropy_loss(o, mk, size_average=False) for o in outs]
            loss = (1 - epoch_frac) * sum(ls[:-1]) + ls[-1]
        losses.append(float(loss.item()))
        (loss / avg_grad_every_n).backward()
        counter += 1
        if counter % avg_grad_every_n == 0:
            with torch.no_grad():
                for grp in groups:
                    for k in grp["keys"]:
                        p = params[k]
                        if p.grad is None:
                            continue
                        new_p, bufs[k] = sgd_step(p.detach(), p.grad, bufs[k], grp["lr"], grp["weight_decay"], momentum)
                        p.copy_(new_p)
                for p in params.values():
                    p.grad = None
            counter = 0
    return {k: v.detach().clone() for k, v in params.items()}, losses


# --------------------------------------------------------------------------
# consumer-side post-processing
# --------------------------------------------------------------------------
def probabilities(logits: torch.Tensor) -> torch.Tensor:
    """1/(1+exp(-x)) as the test loop applies it on the host
    (reference util/experiment_helper.py:57)."""
    return 1.0 / (1.0 + torch.exp(-logits))


def binarise(prob: torch.Tensor) -> torch.Tensor:
    """>=0.5 -> 1 else 0 (reference run_webcam.py:92-93), uint8."""
    return (prob >= 0.5).to(torch.uint8)


def mask_iou_counts(a: torch.Tensor, b: torch.Tensor) -> Tuple[int, int]:
    """(intersection, union) pixel counts of two {0,1} masks: the integers the
    DAVIS J (region IoU) measure is the ratio of."""
    a = a.bool()
    b = b.bool()
    return int((a & b).sum().item()), int((a | b).sum().item())
