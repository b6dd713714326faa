"""Plain-numpy restatement of the third-party primitives on the OSVOS hot path.
TEST INFRASTRUCTURE ONLY (see oracle/osvos_oracle.py for the rules).

The reference's convolution / pooling / transposed-convolution arithmetic is
not in ``/root/reference``: it is PyTorch's (``torch==0.4.0``, reference
``README.md:11``; call sites ``src/networks/osvos_vgg.py:42,45,47,48,56,90,
92,93``).  This file writes those published definitions down once more, with
no torch in them, in float64, for small cases -- a second opinion that the
torch-functional oracle and the CUDA kernels are checked against.

Definitions (PyTorch docs, torch.nn.Conv2d / MaxPool2d / ConvTranspose2d):
  conv2d         out[n,o,y,x] = b[o] + sum_{c,r,s} w[o,c,r,s] * in[n,c,y+r-p,x+s-p]   (zero outside)
  max_pool2d     k=2, s=2, ceil_mode=True: out size ceil(h/2); windows clipped at the border
  conv_transpose out[n,o,iy*s+r,ix*s+t] += in[n,c,iy,ix] * w[c,o,r,t]; out size (h-1)*s+k
"""
from __future__ import annotations

import numpy as np


def conv2d(x: np.ndarray, w: np.ndarray, b=None, padding: int = 0) -> np.ndarray:
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    n, c, h, wd = x.shape
    o, c2, kh, kw = w.shape
    assert c == c2
    xp = np.zeros((n, c, h + 2 * padding, wd + 2 * padding))
    xp[:, :, padding:padding + h, padding:padding + wd] = x
    oh, ow = h + 2 * padding - kh + 1, wd + 2 * padding - kw + 1
    out = np.zeros((n, o, oh, ow))
    for r in range(kh):
        for s in range(kw):
            out += np.einsum("nchw,oc->nohw", xp[:, :, r:r + oh, s:s + ow], w[:, :, r, s])
    if b is not None:
        out += np.asarray(b, np.float64)[None, :, None, None]
    return out


def relu(x: np.ndarray) -> np.ndarray:
    return np.maximum(x, 0.0)


def max_pool2x2_ceil(x: np.ndarray) -> np.ndarray:
    n, c, h, w = x.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    out = np.full((n, c, oh, ow), -np.inf)
    for dy in range(2):
        for dx in range(2):
            sub = x[:, :, dy::2, dx::2]
            out[:, :, :sub.shape[2], :sub.shape[3]] = np.maximum(out[:, :, :sub.shape[2], :sub.shape[3]], sub)
    return out


def conv_transpose2d(x: np.ndarray, w: np.ndarray, stride: int) -> np.ndarray:
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    n, c, h, wd = x.shape
    c2, o, k, k2 = w.shape
    assert c == c2 and k == k2
    out = np.zeros((n, o, (h - 1) * stride + k, (wd - 1) * stride + k))
    for r in range(k):
        for t in range(k):
            out[:, :, r:r + (h - 1) * stride + 1:stride, t:t + (wd - 1) * stride + 1:stride] += \
                np.einsum("nchw,co->nohw", x, w[:, :, r, t])
    return out


def center_crop(x: np.ndarray, height: int, width: int) -> np.ndarray:
    """Negative-pad crop of reference layers/osvos_layers.py:47-54:
    left/top pad = ceil(-d/2), right/bottom = floor(-d/2)."""
    import math
    dh, dw = x.shape[2] - height, x.shape[3] - width
    top, bottom = -math.ceil(-dh / 2), -math.floor(-dh / 2)
    left, right = -math.ceil(-dw / 2), -math.floor(-dw / 2)
    assert min(top, bottom, left, right) >= 0, "oracle numpy crop only crops"
    return x[:, :, top:x.shape[2] - bottom, left:x.shape[3] - right]


def vgg_forward(sd: dict, x: np.ndarray):
    """reference networks/osvos_vgg.py:61-83 in float64 numpy."""
    from .osvos_oracle import stage_conv_indices
    g = lambda k: None if k not in sd else np.asarray(sd[k], np.float64)  # noqa: E731
    H, W = x.shape[2], x.shape[3]
    side, side_out = [], []
    x = np.asarray(x, np.float64)
    for si, idxs in enumerate(stage_conv_indices()):
        if si > 0:
            x = max_pool2x2_ceil(x)
        for mi in idxs:
            x = relu(conv2d(x, g(f"stages.{si}.{mi}.weight"), g(f"stages.{si}.{mi}.bias"), 1))
        if si > 0:
            i = si - 1
            sp = conv2d(x, g(f"side_prep.{i}.weight"), g(f"side_prep.{i}.bias"), 1)
            side.append(center_crop(conv_transpose2d(sp, g(f"upscale.{i}.weight"), 2 ** si), H, W))
            sc = conv2d(sp, g(f"score_dsn.{i}.weight"), g(f"score_dsn.{i}.bias"), 0)
            side_out.append(center_crop(conv_transpose2d(sc, g(f"upscale_.{i}.weight"), 2 ** si), H, W))
    out = conv2d(np.concatenate(side, 1), g("fuse.weight"), g("fuse.bias"), 0)
    return side_out + [out]
