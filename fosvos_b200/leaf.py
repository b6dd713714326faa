"""Leaf modules of ``OSVOS_VGG``: ordinary ``nn.Conv2d`` / ``nn.ReLU`` / ``nn.MaxPool2d`` /
``nn.ConvTranspose2d`` subclasses (so ``isinstance`` checks, ``.weight`` / ``.in_channels`` ... and
``state_dict`` keys are the reference's, ``osvos_vgg.py:42-48,85-95``) whose ``forward`` runs THIS repo's
kernels when a module is called on its own.

``OSVOS_VGG.forward`` normally never calls them: the network is one fused pipeline.  They are called
  * by hook-style consumers that walk the module tree themselves, as the reference's ``FilterPruner.forward``
    does (``prune.py:94-103``: ``x = l(x); x.register_hook(...)``), and
  * by ``OSVOS_VGG`` itself in *introspection mode* -- when a forward (pre-)hook is registered on any leaf
    of ``stages`` / ``side_prep`` (or ``net.introspect = True``) -- so that ``register_forward_hook`` and
    tensor hooks on per-conv outputs fire with the kernels' own results.
Each call is a ``torch.autograd.Function`` over NCHW fp32 tensors (the module boundary layout); inside, the
tensor is re-laid out to NHWC and pushed through the same conv / pool kernels the fused pipeline uses, so
this path is slower (two layout changes per module) but runs no cuDNN kernel.  The 1x1 heads and the
transposed convolutions exist only fused into the side-chain kernel: calling those leaves directly raises.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib as L
from . import ops


def _precision_of(m) -> str:
    p = getattr(m, "_fosvos_precision", None) or os.environ.get("FOSVOS_PRECISION", "bf16")
    # the split-operand inference modes have no module-level kernels: a leaf called on its own computes in strict fp32
    return "fp32" if p in ("fp32_tc", "bf16x3") else p


def _act_dtype(precision: str) -> torch.dtype:
    return torch.float32 if precision == "fp32" else torch.bfloat16


def _require_cuda(x: torch.Tensor, what: str) -> None:
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"fosvos_b200 {what}: runs on CUDA sm_100 devices only; there is no CPU fallback")
    L.require_device(x.device)


class _Conv3x3Fn(torch.autograd.Function):
    """y = conv3x3(x, w) + b through ``fosvos_conv3x3_{tc,simt}``; backward = data gradient through the same kernel
    (flipped/transposed weights) + ``fosvos_conv3x3_wgrad_*``."""

    @staticmethod
    def forward(ctx, x, weight, bias, precision):
        dt = _act_dtype(precision)
        tc = precision == "bf16"
        impl = "tc" if tc else "simt"
        cout, cin = int(weight.shape[0]), int(weight.shape[1])
        xh = ops.nchw_to_nhwc(x.detach().float().contiguous(), dt)
        wp = ops.pack_weight(weight.detach(), L.W_TC_FWD if tc else L.W_SIMT_FWD, dt)
        bp = ops.pad_bias(None if bias is None else bias.detach(), cout, x.device)
        yh = ops.conv3x3(xh, wp, bp, ops.pad8(cout), L.CONV_BIAS, impl=impl)
        ctx.save_for_backward(xh, weight)
        ctx.meta = (dt, tc, impl, cout, cin, bias is not None)
        return ops.nhwc_to_nchw(yh, cout)

    @staticmethod
    def backward(ctx, dy):
        xh, weight = ctx.saved_tensors
        dt, tc, impl, cout, cin, has_bias = ctx.meta
        dyh = ops.nchw_to_nhwc(dy.float().contiguous(), dt, cp=ops.pad8(cout))
        dx = None
        if ctx.needs_input_grad[0]:
            wd = ops.pack_weight(weight.detach(), L.W_TC_DGRAD if tc else L.W_SIMT_DGRAD, dt)
            dxh = ops.conv3x3(dyh, wd, None, xh.shape[3], 0, impl=impl)
            dx = ops.nhwc_to_nchw(dxh, cin)
        dw = torch.zeros_like(weight, memory_format=torch.contiguous_format)
        db = torch.zeros(cout, dtype=torch.float32, device=weight.device) if has_bias else None
        ops.conv3x3_wgrad(xh, dyh, dw, db, impl=impl)
        return dx, dw, db, None


class _ReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = ops.relu_fwd(x.detach().float())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return ops.relu_bwd(y, dy.float())


class _PoolFn(torch.autograd.Function):
    """nn.MaxPool2d(2, 2, ceil_mode=True) (osvos_vgg.py:90): fp32 NHWC through ``fosvos_maxpool2x2_*`` (exact)."""

    @staticmethod
    def forward(ctx, x):
        c = int(x.shape[1])
        xh = ops.nchw_to_nhwc(x.detach().float().contiguous(), torch.float32)
        ctx.save_for_backward(xh)
        ctx.c = c
        return ops.nhwc_to_nchw(ops.maxpool2x2(xh), c)

    @staticmethod
    def backward(ctx, dy):
        (xh,) = ctx.saved_tensors
        dyh = ops.nchw_to_nhwc(dy.float().contiguous(), torch.float32, cp=xh.shape[3])
        return ops.nhwc_to_nchw(ops.maxpool2x2_bwd(xh, dyh), ctx.c)


class Conv2d(nn.Conv2d):
    def forward(self, x):
        _require_cuda(x, "Conv2d")
        if tuple(self.kernel_size) != (3, 3) or tuple(self.stride) != (1, 1) or tuple(self.padding) != (1, 1) or \
                tuple(self.dilation) != (1, 1) or self.groups != 1:
            raise RuntimeError("fosvos_b200: this convolution (score_dsn / fuse 1x1) exists only fused into the side-chain kernel; "
                               "call the network, not the leaf (the fused path computes it at low resolution)")
        return _Conv3x3Fn.apply(x, self.weight, self.bias, _precision_of(self))


class ReLU(nn.ReLU):
    def forward(self, x):
        _require_cuda(x, "ReLU")
        return _ReluFn.apply(x)


class MaxPool2d(nn.MaxPool2d):
    def forward(self, x):
        _require_cuda(x, "MaxPool2d")
        if (self.kernel_size, self.stride, self.ceil_mode, self.padding, self.dilation) != (2, 2, True, 0, 1):
            raise RuntimeError("fosvos_b200: only MaxPool2d(kernel_size=2, stride=2, ceil_mode=True) (osvos_vgg.py:90) has a kernel")
        return _PoolFn.apply(x)


class ConvTranspose2d(nn.ConvTranspose2d):
    def forward(self, x, output_size=None):
        raise RuntimeError("fosvos_b200: the up-sampling transposed convolutions exist only fused into the side-chain kernel "
                           "(4 taps per pixel instead of a dense 16x16 ConvT); call the network, not the leaf")


_ADOPT = {nn.Conv2d: Conv2d, nn.ReLU: ReLU, nn.MaxPool2d: MaxPool2d, nn.ConvTranspose2d: ConvTranspose2d}


def adopt(module: nn.Module, precision: str) -> None:
    """Turn the plain torch leaves below ``module`` into this file's classes IN PLACE (class swap: parameters, hooks and
    ``state_dict`` keys are untouched) -- prune-style surgery assigns fresh ``nn.Conv2d`` s (prune.py:490-514) -- and
    hand every leaf the network's precision mode."""
    for m in module.modules():
        cls = _ADOPT.get(type(m))
        if cls is not None:
            m.__class__ = cls
        if isinstance(m, (Conv2d, ReLU, MaxPool2d, ConvTranspose2d)):
            m.__dict__["_fosvos_precision"] = precision
