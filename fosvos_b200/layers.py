"""Drop-in for the reference ``layers/osvos_layers.py`` (same names and signatures).

``class_balanced_cross_entropy_loss`` runs as two hand-written kernels (forward reduction,
backward) instead of ~30 ATen launches; the small host-side helpers keep the reference's
exact semantics.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch.nn import functional as F

from . import _lib as L
from . import ops


def logit(x):
    # reference osvos_layers.py:9-10
    return np.log(x / (1 - x + 1e-08) + 1e-08)


def sigmoid_np(x):
    # reference osvos_layers.py:13-14
    return 1 / (1 + np.exp(-x))


class _BalancedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, label, size_average):
        L.require_device(output.device)
        out32 = output.detach().to(torch.float32).contiguous()
        lab32 = label.detach().to(torch.float32).contiguous()
        loss, stats = ops.bal_loss_fwd(out32, lab32, size_average)
        ctx.save_for_backward(out32, lab32, stats)
        ctx.size_average = size_average
        ctx.out_shape = output.shape
        ctx.out_dtype = output.dtype
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        out32, lab32, stats = ctx.saved_tensors
        dx = ops.bal_loss_bwd(out32, lab32, ctx.size_average, stats, grad_out)
        return dx.view(ctx.out_shape).to(ctx.out_dtype), None, None


def class_balanced_cross_entropy_loss(output, label, size_average=True):
    """Class-balanced sigmoid cross entropy (reference osvos_layers.py:17-44).

    Args:
    output: Output of the network (logits), (N,C,H,W)
    label: Ground truth label (binarised at >= 0.5), same shape
    Returns:
    0-dim tensor; differentiable w.r.t. ``output``.
    """
    if output.numel() != label.numel():
        raise RuntimeError(f"output {tuple(output.shape)} and label {tuple(label.shape)} must have the same number of elements")
    return _BalancedLoss.apply(output, label, bool(size_average))


class _PixelLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, kind, size_average):
        L.require_device(output.device)
        out32 = output.detach().to(torch.float32).contiguous()
        tgt32 = target.detach().to(torch.float32).contiguous()
        loss, dx = ops.pixel_loss(out32, tgt32, kind, size_average, 1.0, want_grad=True)
        ctx.save_for_backward(dx)
        ctx.out_shape, ctx.out_dtype = output.shape, output.dtype
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (dx,) = ctx.saved_tensors
        return (dx * grad_out).view(ctx.out_shape).to(ctx.out_dtype), None, None, None


def mse_loss(output, target, size_average=False):
    """``nn.MSELoss(size_average=False)`` of the distillation driver (reference mimic.py:76): one fused kernel."""
    return _PixelLoss.apply(output, target, "mse", bool(size_average))


def l1_loss(output, target, size_average=False):
    """``nn.L1Loss(size_average=False)`` (reference mimic.py:79)."""
    return _PixelLoss.apply(output, target, "l1", bool(size_average))


def center_crop(x, height, width):
    """Negative-pad centre crop, reference osvos_layers.py:47-54 (left/top get ceil(-d/2))."""
    crop_h = -(x.size()[2] - height) / 2.0
    crop_w = -(x.size()[3] - width) / 2.0
    return F.pad(x, [int(math.ceil(crop_w)), int(math.floor(crop_w)), int(math.ceil(crop_h)), int(math.floor(crop_h))])


def upsample_filt(size):
    """Bilinear interpolation kernel, float64 (reference osvos_layers.py:57-65)."""
    factor = (size + 1) // 2
    if size % 2 == 1:
        center = factor - 1
    else:
        center = factor - 0.5
    og = np.ogrid[:size, :size]
    return (1 - abs(og[0] - center) / factor) * \
           (1 - abs(og[1] - center) / factor)


def interp_surgery(lay):
    """Fill the channel diagonal of a ConvTranspose2d with the bilinear kernel
    (reference osvos_layers.py:70-81).  Returns ``lay.weight.data``."""
    m, k, h, w = lay.weight.data.size()
    if m != k:
        raise Exception('input + output channels need to be the same')
    if h != w:
        raise Exception('filters need to be square')
    filt = torch.from_numpy(upsample_filt(h)).to(lay.weight.dtype)
    for i in range(m):
        lay.weight.data[i, i, :, :].copy_(filt)
    return lay.weight.data
