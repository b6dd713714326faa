"""Seeded synthetic DAVIS-shaped inputs and calibrated random weights.

Host-side data generation only (CPU torch, no kernels): the reference reads
DAVIS JPEG/PNG files (``src/dataloaders/davis_2016.py:115-134``), there is no
dataset in this environment, so frames, first-frame masks and a parent-network
``state_dict`` are synthesised deterministically.  SURVEY.md §8(d) fixes the
recipe so that both sides of every parity test see identical bytes.

The reference's default init (N(0,1e-3^2), ``osvos_vgg.py:99-102``) yields
fused logits of ~1e-10, useless for parity; ``make_state_dict`` therefore
draws He-normal weights and ``calibrate`` rescales the 1x1 heads so logits have
a useful spread.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# BGR mean the reference subtracts (davis_2016.py:24,127-128)
MEAN_BGR = (104.00699, 116.66877, 122.67892)

_LAY = [[64, 64], [128, 128], [256, 256, 256], [512, 512, 512], [512, 512, 512]]
_CONV_IDX = [[0, 2], [1, 3], [1, 3, 5], [1, 3, 5], [1, 3, 5]]


def ellipse_mask(seq: int, frame: int, H: int, W: int) -> torch.Tensor:
    """{0,1} float mask (H,W): an ellipse covering ~10-25 % of the pixels whose
    centre drifts with the frame index."""
    g = torch.Generator().manual_seed(4321 + seq)
    frac = 0.10 + 0.15 * torch.rand(1, generator=g).item()
    aspect = 0.6 + 0.8 * torch.rand(1, generator=g).item()
    phase = 2 * math.pi * torch.rand(1, generator=g).item()
    area = frac * H * W
    ry = math.sqrt(area / (math.pi * aspect))
    rx = aspect * ry
    cy = H / 2 + 0.15 * H * math.sin(phase + 0.08 * frame)
    cx = W / 2 + 0.20 * W * math.cos(phase + 0.05 * frame)
    yy = torch.arange(H, dtype=torch.float32).view(H, 1)
    xx = torch.arange(W, dtype=torch.float32).view(1, W)
    return ((((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) <= 1.0).float()


def make_frame(seq: int, frame: int, H: int = 480, W: int = 854,
               noise: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """One synthetic frame -> (image fp32 (1,3,H,W) mean-subtracted BGR, mask fp32
    (1,1,H,W) in {0,1}).  ``noise=True`` gives the worst-case randn*60 variant."""
    g = torch.Generator().manual_seed(1234 + seq * 1000 + frame)
    m = ellipse_mask(seq, frame, H, W)
    if noise:
        x = torch.randn(1, 3, H, W, generator=g) * 60.0
        return x, m.view(1, 1, H, W)
    u8 = torch.randint(0, 256, (1, 3, H, W), generator=g).float()
    # 8x8 box low-pass (same-size), then a brighter ellipse
    blur = F.avg_pool2d(F.pad(u8, [3, 4, 3, 4], mode="replicate"), 8, stride=1)
    img = 0.5 * blur + 0.5 * blur.mean() * 0.6
    img = torch.clamp(img + 90.0 * m.view(1, 1, H, W), 0, 255).round()
    mean = torch.tensor(MEAN_BGR, dtype=torch.float32).view(1, 3, 1, 1)
    return img - mean, m.view(1, 1, H, W)


def make_sequence(seq: int, n_frames: int, H: int = 480, W: int = 854):
    """(frames (n,3,H,W), masks (n,1,H,W)) of one synthetic sequence."""
    fr = [make_frame(seq, f, H, W) for f in range(n_frames)]
    return torch.cat([a for a, _ in fr]), torch.cat([b for _, b in fr])


def upsample_filt_t(k: int) -> torch.Tensor:
    factor = (k + 1) // 2
    center = factor - 1 if k % 2 == 1 else factor - 0.5
    r = torch.arange(k, dtype=torch.float64)
    f = 1 - (r - center).abs() / factor
    return (f.view(k, 1) * f.view(1, k)).float()


def make_state_dict(seed: int = 0, kind: str = "random",
                    channels: Optional[Sequence[Sequence[int]]] = None,
                    stage_bias: bool = True) -> Dict[str, torch.Tensor]:
    """Reference-layout ``state_dict`` (same keys/order/shapes as
    ``OSVOS_VGG.state_dict()``, reference osvos_vgg.py:50-56) with usable
    random weights.  kind='random': He-normal 3x3 convs, biases U(-0.1,0.1);
    kind='structured': all-positive filters of sum 1.5 (a brightness
    detector, so logits are bimodal like a trained net's);
    kind='parent': the same filters with gain 0.3 in the first conv and 1.0
    elsewhere, so activations stay O(10) at every depth as in a trained
    parent network -- with 'structured' they grow 2-3x per stage to ~1e4 and
    the reference's lr = 1e-8 on the sum-reduced loss (train_online.py:81,
    network_provider.py:144-159) makes the online fine-tune diverge (in the
    oracle exactly as on the GPU); with 'parent' the loss decreases
    monotonically.  The fine-tune benchmarks and their parity tests use it."""
    g = torch.Generator().manual_seed(seed)
    widths = _LAY if channels is None else [list(c) for c in channels]
    sd: Dict[str, torch.Tensor] = {}
    for i in range(4):
        k = 2 ** (2 + i)
        w = torch.zeros(16, 16, k, k)
        f = upsample_filt_t(k)
        for c in range(16):
            w[c, c] = f
        sd[f"upscale.{i}.weight"] = w
    for i in range(4):
        k = 2 ** (2 + i)
        sd[f"upscale_.{i}.weight"] = upsample_filt_t(k).view(1, 1, k, k).clone()

    def conv_w(cout, cin):
        w = torch.randn(cout, cin, 3, 3, generator=g) * math.sqrt(2.0 / (9 * cin))
        if kind in ("structured", "parent"):
            w = w.abs()
            gain = 1.5 if kind == "structured" else (0.3 if cin == 3 else 1.0)
            w = w / w.sum(dim=(1, 2, 3), keepdim=True) * gain
        return w

    def conv_b(cout):
        b = torch.rand(cout, generator=g) * 0.2 - 0.1
        return b * (20.0 if kind == "random" else 1.0)

    cin = 3
    for si in range(5):
        for j, mi in enumerate(_CONV_IDX[si]):
            cout = widths[si][j]
            sd[f"stages.{si}.{mi}.weight"] = conv_w(cout, cin)
            if stage_bias:
                sd[f"stages.{si}.{mi}.bias"] = conv_b(cout)
            cin = cout
    for i in range(4):
        sd[f"side_prep.{i}.weight"] = conv_w(16, widths[i + 1][-1])
        sd[f"side_prep.{i}.bias"] = conv_b(16)
    for i in range(4):
        w = torch.randn(1, 16, 1, 1, generator=g) * 0.25
        sd[f"score_dsn.{i}.weight"] = w.abs() if kind in ("structured", "parent") else w
        sd[f"score_dsn.{i}.bias"] = torch.zeros(1)
    w = torch.randn(1, 64, 1, 1, generator=g) * 0.125
    sd["fuse.weight"] = w.abs() if kind in ("structured", "parent") else w
    sd["fuse.bias"] = torch.zeros(1)
    return sd


def calibrate(sd: Dict[str, torch.Tensor], forward_fn: Callable[[Dict[str, torch.Tensor], torch.Tensor], List[torch.Tensor]],
              x: torch.Tensor, target_std: float = 3.0, mask: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Rescale the 1x1 heads so that each of the five logit maps has standard
    deviation ``target_std`` on ``x``.  With ``mask`` (structured weights) the
    bias is also set so the decision boundary falls between the mean logit of
    foreground and background."""
    sd = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        outs = [o.float().cpu() for o in forward_fn(sd, x)]
    for i in range(4):
        s = outs[i].std().item()
        if s > 0:
            sd[f"score_dsn.{i}.weight"] *= target_std / s
    s = outs[4].std().item()
    if s > 0:
        sd["fuse.weight"] *= target_std / s
    if mask is not None:
        with torch.no_grad():
            outs = [o.float().cpu() for o in forward_fn(sd, x)]
        m = mask.bool().cpu()
        for i in range(5):
            o = outs[i]
            mid = 0.5 * (o[m].mean().item() + o[~m].mean().item())
            key = f"score_dsn.{i}.bias" if i < 4 else "fuse.bias"
            sd[key] = sd[key] - mid
    else:
        with torch.no_grad():
            outs = [o.float().cpu() for o in forward_fn(sd, x)]
        for i in range(5):
            key = f"score_dsn.{i}.bias" if i < 4 else "fuse.bias"
            sd[key] = sd[key] - outs[i].mean().item()
    return sd


def pruned_channels(keep: float = 0.5, n_min: int = 4) -> List[List[int]]:
    """Per-stage widths of a channel-pruned VGG keeping ceil(keep*C) filters of
    every stage conv, never fewer than ``n_min`` (reference prune.py:30)."""
    return [[max(n_min, int(math.ceil(keep * c))) for c in st] for st in _LAY]


def prune_state_dict(sd: Dict[str, torch.Tensor], keep: float = 0.5, n_min: int = 4) -> Dict[str, torch.Tensor]:
    """Deterministic stand-in for the reference's Taylor-ranked pruning
    (prune.py:190-223,490-514): drop the lowest-L2-norm filters of every stage
    conv, slice the consumers on Cin, and drop the stage-conv biases (the
    reference rebuilds pruned convs with bias=False, prune.py:500)."""
    out: Dict[str, torch.Tensor] = {k: v.clone() for k, v in sd.items() if not (k.startswith("stages.") and k.endswith("bias"))}
    keep_in = torch.arange(3)
    for si in range(5):
        for mi in _CONV_IDX[si]:
            w = out[f"stages.{si}.{mi}.weight"][:, keep_in]
            n_keep = max(n_min, int(math.ceil(keep * w.shape[0])))
            norms = w.flatten(1).norm(dim=1)
            keep_out = torch.sort(torch.topk(norms, n_keep).indices).values
            out[f"stages.{si}.{mi}.weight"] = w[keep_out].contiguous()
            keep_in = keep_out
        if si > 0:
            out[f"side_prep.{si - 1}.weight"] = out[f"side_prep.{si - 1}.weight"][:, keep_in].contiguous()
    return out
