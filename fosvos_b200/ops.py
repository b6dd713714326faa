"""Tensor-level wrappers around the C ABI (one function per entry point).

Every function takes/returns CUDA tensors, enqueues on the current stream and never
synchronises.  Activations are NHWC (``(N,H,W,Cp)``, ``Cp % 8 == 0``) in fp32 or bf16.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L


def pad8(c: int) -> int:
    return (c + 7) // 8 * 8


def nchw_to_nhwc(x: torch.Tensor, dtype: torch.dtype, cp: Optional[int] = None) -> torch.Tensor:
    L.require_device(x.device)
    assert x.dtype == torch.float32 and x.dim() == 4
    x = x.contiguous()
    n, c, h, w = x.shape
    cp = pad8(c) if cp is None else cp
    y = torch.empty((n, h, w, cp), dtype=dtype, device=x.device)
    L.check(L.lib().fosvos_nchw_to_nhwc(x.data_ptr(), y.data_ptr(), n, c, h, w, cp, L.dtype_code(dtype), L.stream()), "nchw_to_nhwc")
    return y


def nhwc_to_nchw(x: torch.Tensor, c: Optional[int] = None) -> torch.Tensor:
    L.require_device(x.device)
    n, h, w, cp = x.shape
    c = cp if c is None else c
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    L.check(L.lib().fosvos_nhwc_to_nchw(x.data_ptr(), y.data_ptr(), n, c, h, w, cp, L.dtype_code(x.dtype), L.stream()), "nhwc_to_nchw")
    return y


def pack_weight(w: torch.Tensor, layout: int, dtype: torch.dtype, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """OIHW fp32 (Cout,Cin,3,3) -> packed kernel layout (see fosvos_wlayout)."""
    L.require_device(w.device)
    assert w.dtype == torch.float32 and w.dim() == 4 and w.shape[2:] == (3, 3)
    w = w.detach().contiguous()
    cout, cin = w.shape[0], w.shape[1]
    coutp, cinp = pad8(cout), pad8(cin)
    n = L.lib().fosvos_packed_weight_elems(coutp, cinp, layout)
    if out is None:
        out = torch.empty(n, dtype=dtype, device=w.device)
    assert out.numel() == n and out.dtype == dtype
    L.check(L.lib().fosvos_pack_conv3x3_weight(w.data_ptr(), out.data_ptr(), cout, cin, coutp, cinp, layout,
                                               L.dtype_code(dtype), L.stream()), "pack_conv3x3_weight")
    return out


def pad_bias(b: Optional[torch.Tensor], c: int, device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    cp = pad8(c)
    if out is None:
        out = torch.empty(cp, dtype=torch.float32, device=device)
    L.check(L.lib().fosvos_pad_bias(L.ptr(None if b is None else b.detach()), out.data_ptr(), c, cp, L.stream()), "pad_bias")
    return out


def conv3x3(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], cout_p: int, flags: int,
            mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, impl: str = "tc") -> torch.Tensor:
    """3x3/s1/p1 convolution over NHWC activations.  impl: 'tc' (tcgen05, bf16) or 'simt'."""
    L.require_device(x.device)
    n, h, w, cin_p = x.shape
    if out is None:
        assert not (flags & L.CONV_ACCUMULATE)
        out = torch.empty((n, h, w, cout_p), dtype=x.dtype, device=x.device)
    assert out.shape == (n, h, w, cout_p) and out.dtype == x.dtype and x.is_contiguous() and out.is_contiguous()
    if mask is not None:
        assert mask.shape == out.shape and mask.dtype == x.dtype and mask.is_contiguous()
    if impl == "tc":
        assert x.dtype == torch.bfloat16, "the tcgen05 path computes in bf16"
        rc = L.lib().fosvos_conv3x3_tc(x.data_ptr(), w_packed.data_ptr(), L.ptr(bias), L.ptr(mask), out.data_ptr(),
                                       n, h, w, cin_p, cout_p, flags, L.stream())
    elif impl == "simt":
        rc = L.lib().fosvos_conv3x3_simt(x.data_ptr(), w_packed.data_ptr(), L.ptr(bias), L.ptr(mask), out.data_ptr(),
                                         n, h, w, cin_p, cout_p, flags, L.dtype_code(x.dtype), L.stream())
    else:
        raise ValueError(impl)
    L.check(rc, f"conv3x3_{impl}")
    return out


def conv3x3_pool(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], cout_p: int, flags: int):
    """tcgen05 conv + ReLU with the following 2x2/2 ceil-mode max pool written from the same epilogue
    -> (y (N,H,W,C), pooled (N,ceil(H/2),ceil(W/2),C))."""
    L.require_device(x.device)
    n, h, w, cin_p = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    y = torch.empty((n, h, w, cout_p), dtype=x.dtype, device=x.device)
    yp = torch.empty((n, (h + 1) // 2, (w + 1) // 2, cout_p), dtype=x.dtype, device=x.device)
    L.check(L.lib().fosvos_conv3x3_tc_pool(x.data_ptr(), w_packed.data_ptr(), L.ptr(bias), y.data_ptr(), yp.data_ptr(),
                                           n, h, w, cin_p, cout_p, flags, L.stream()), "conv3x3_tc_pool")
    return y, yp


def conv3x3_pool_arg(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], cout_p: int, flags: int,
                     want_y: bool = True):
    """Training form of the fused pool: (y or None, pooled, arg) -- ``arg`` (N, PH, PW, cout_p / 32, 2) int32 holds the window
    index of each pooled element's (first) maximum as two bit planes per 32 channels, which ``maxpool2x2_bwd_arg`` consumes
    instead of re-reading ``y``.  ``want_y=False``: the full-resolution map is not written at all."""
    L.require_device(x.device)
    n, h, w, cin_p = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and cout_p % 32 == 0
    ph, pw = (h + 1) // 2, (w + 1) // 2
    y = torch.empty((n, h, w, cout_p), dtype=x.dtype, device=x.device) if want_y else None
    yp = torch.empty((n, ph, pw, cout_p), dtype=x.dtype, device=x.device)
    arg = torch.empty((n, ph, pw, cout_p // 32, 2), dtype=torch.int32, device=x.device)
    L.check(L.lib().fosvos_conv3x3_tc_pool_arg(x.data_ptr(), w_packed.data_ptr(), L.ptr(bias), L.ptr(y), yp.data_ptr(), arg.data_ptr(),
                                               n, h, w, cin_p, cout_p, flags, L.stream()), "conv3x3_tc_pool_arg")
    return y, yp, arg


def maxpool2x2_bwd_arg(arg: torch.Tensor, dy: torch.Tensor, h: int, w: int, add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Gradient of the pool w.r.t. its (n, h, w, c) input from the recorded window indices; with ``add`` the result is
    written INTO it (fan-in, as ``maxpool2x2_bwd``)."""
    L.require_device(dy.device)
    n, ph, pw, c = dy.shape
    assert (ph, pw) == ((h + 1) // 2, (w + 1) // 2) and c % 32 == 0 and dy.is_contiguous()
    assert arg.dtype == torch.int32 and arg.shape == (n, ph, pw, c // 32, 2) and arg.is_contiguous()
    if add is not None:
        assert add.shape == (n, h, w, c) and add.dtype == dy.dtype and add.is_contiguous()
    dx = torch.empty((n, h, w, c), dtype=dy.dtype, device=dy.device) if add is None else add
    L.check(L.lib().fosvos_maxpool2x2_bwd_arg(arg.data_ptr(), dy.data_ptr(), L.ptr(add), dx.data_ptr(), n, h, w, c,
                                              L.dtype_code(dy.dtype), L.stream()), "maxpool2x2_bwd_arg")
    return dx


def conv3x3_pool_only(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], cout_p: int, flags: int) -> torch.Tensor:
    """Same launch with the full-resolution output suppressed: only the pooled map leaves the kernel (inference, when
    nothing but the pool consumes the conv -- conv1_2, osvos_vgg.py:63,68)."""
    L.require_device(x.device)
    n, h, w, cin_p = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    yp = torch.empty((n, (h + 1) // 2, (w + 1) // 2, cout_p), dtype=x.dtype, device=x.device)
    L.check(L.lib().fosvos_conv3x3_tc_pool(x.data_ptr(), w_packed.data_ptr(), L.ptr(bias), None, yp.data_ptr(),
                                           n, h, w, cin_p, cout_p, flags, L.stream()), "conv3x3_tc_pool")
    return yp


def side_tc_supported(cin_p: int) -> bool:
    return bool(L.lib().fosvos_conv3x3_side_tc_supported(int(cin_p)))


def conv3x3_side(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
                 zs: Optional[torch.Tensor] = None, heads: Optional[torch.Tensor] = None, want_y: bool = True) -> Optional[torch.Tensor]:
    """side_prep convolution (C -> 16, bias, no ReLU; osvos_vgg.py:42,69) through the row-stacked tcgen05 kernel.
    ``zs`` (N*h*w, 2) fp32: the two 1x1 heads of the side chain are written there from the fp32 accumulators, with
    ``heads`` the stage's 36 head parameters (a view into the side-chain parameter block).  Returns y (N,h,w,16) bf16
    (or None with ``want_y=False``)."""
    L.require_device(x.device)
    n, h, w, cin_p = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    if want_y and out is None:
        out = torch.empty((n, h, w, 16), dtype=x.dtype, device=x.device)
    if out is not None:
        assert out.shape == (n, h, w, 16) and out.dtype == x.dtype and out.is_contiguous()
    if zs is not None:
        assert zs.dtype == torch.float32 and zs.numel() == 2 * n * h * w and zs.is_contiguous()
        assert heads is not None and heads.dtype == torch.float32 and heads.numel() >= 36
    L.check(L.lib().fosvos_conv3x3_side_tc(x.data_ptr(), w_packed.data_ptr(), L.ptr(bias), L.ptr(out), L.ptr(zs), L.ptr(heads),
                                           n, h, w, cin_p, L.stream()), "conv3x3_side_tc")
    return out


# ---- fp32 through the bf16 tensor cores (split operands; csrc/split.cu, conv_tc.cu SPLIT) ---------------------------------
def seg64(c: int) -> int:
    return (c + 63) // 64 * 64


def split_frames(x: torch.Tensor, terms: int) -> torch.Tensor:
    """NCHW fp32 frames -> split map (N,H,W, terms * seg64(C)) bf16: the bf16 terms of every value side by side."""
    L.require_device(x.device)
    assert x.dtype == torch.float32 and x.dim() == 4
    x = x.contiguous()
    n, c, h, w = x.shape
    seg = seg64(c)
    y = torch.empty((n, h, w, terms * seg), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().fosvos_split_nchw(x.data_ptr(), y.data_ptr(), n, c, h, w, seg, terms, L.stream()), "split_nchw")
    return y


def pack_weight_split(w: torch.Tensor, terms: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """OIHW fp32 (Cout,Cin,3,3) -> the split-operand packed layout (one GEMM-K segment of seg64(Cin) per kept term product)."""
    L.require_device(w.device)
    assert w.dtype == torch.float32 and w.dim() == 4 and w.shape[2:] == (3, 3)
    w = w.detach().contiguous()
    cout, cin = int(w.shape[0]), int(w.shape[1])
    coutp, seg = pad8(cout), seg64(cin)
    n = L.lib().fosvos_packed_weight_split_elems(coutp, seg, terms)
    if out is None:
        out = torch.empty(n, dtype=torch.bfloat16, device=w.device)
    assert out.numel() == n and out.dtype == torch.bfloat16
    L.check(L.lib().fosvos_pack_conv3x3_weight_split(w.data_ptr(), out.data_ptr(), cout, cin, coutp, seg, terms, L.stream()),
            "pack_conv3x3_weight_split")
    return out


def conv3x3_split(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], cout_p: int, terms: int, flags: int) -> torch.Tensor:
    """3x3 conv of a split map on the tensor cores with fp32-equivalent arithmetic.  cout_p > 32: returns the split map of the
    result (N,H,W, terms * seg64(cout_p)); cout_p <= 32 (side_prep): returns plain fp32 NHWC (N,H,W,cout_p)."""
    L.require_device(x.device)
    n, h, w, ct = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and ct % terms == 0
    seg = ct // terms
    if cout_p > 32:
        y = torch.empty((n, h, w, terms * seg64(cout_p)), dtype=torch.bfloat16, device=x.device)
        yp, yf = y.data_ptr(), None
    else:
        y = torch.empty((n, h, w, cout_p), dtype=torch.float32, device=x.device)
        yp, yf = None, y.data_ptr()
    L.check(L.lib().fosvos_conv3x3_tc_split(x.data_ptr(), w_packed.data_ptr(), L.ptr(bias), yp, yf, n, h, w, seg, terms, cout_p, flags,
                                            L.stream()), "conv3x3_tc_split")
    return y


def maxpool2x2_split(x: torch.Tensor, terms: int) -> torch.Tensor:
    L.require_device(x.device)
    n, h, w, ct = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and ct % terms == 0
    y = torch.empty((n, (h + 1) // 2, (w + 1) // 2, ct), dtype=x.dtype, device=x.device)
    L.check(L.lib().fosvos_maxpool2x2_split(x.data_ptr(), y.data_ptr(), n, h, w, ct // terms, terms, L.stream()), "maxpool2x2_split")
    return y


def wgrad_workspace(cin_p: int, cout_p: int, device) -> torch.Tensor:
    """Zeroed [tap][M][N] fp32 accumulator of the tensor-core weight gradient (kept live across micro-iterations)."""
    return torch.zeros(L.lib().fosvos_conv3x3_wgrad_tc_workspace_bytes(cin_p, cout_p) // 4, dtype=torch.float32, device=device)


def conv3x3_wgrad_accumulate(x: torch.Tensor, dz: torch.Tensor, ws: torch.Tensor, db: Optional[torch.Tensor], cout: int) -> None:
    """ws += gradient of this micro-iteration (tensor cores, bf16 operands); db += sum dz."""
    L.require_device(x.device)
    n, h, w, cin_p = x.shape
    cout_p = dz.shape[3]
    assert dz.shape[:3] == x.shape[:3] and x.dtype == dz.dtype == torch.bfloat16 and x.is_contiguous() and dz.is_contiguous()
    assert ws.dtype == torch.float32 and ws.numel() == 9 * cin_p * cout_p
    L.check(L.lib().fosvos_conv3x3_wgrad_tc_accumulate(x.data_ptr(), dz.data_ptr(), L.ptr(db), ws.data_ptr(), n, h, w, cin_p, cout_p,
                                                       cout, L.stream()), "conv3x3_wgrad_tc_accumulate")


def conv3x3_wgrad_finish(ws: torch.Tensor, dw: torch.Tensor, cin_p: int, cout_p: int, zero_workspace: bool = True) -> None:
    """dw (OIHW fp32 .grad) += ws ; ws = 0."""
    L.require_device(ws.device)
    assert dw.dtype == torch.float32 and dw.is_contiguous() and ws.numel() == 9 * cin_p * cout_p
    L.check(L.lib().fosvos_conv3x3_wgrad_tc_finish(ws.data_ptr(), dw.data_ptr(), cin_p, cout_p, dw.shape[1], dw.shape[0],
                                                   int(zero_workspace), L.stream()), "conv3x3_wgrad_tc_finish")


def conv3x3_wgrad(x: torch.Tensor, dz: torch.Tensor, dw: torch.Tensor, db: Optional[torch.Tensor], impl: str = "simt") -> None:
    """dw (OIHW fp32) += x (*) dz ; db += sum dz.  Accumulates in place.
    impl 'tc': tcgen05 kernel (bf16 operands); 'simt': direct fp32-FMA kernel."""
    L.require_device(x.device)
    n, h, w, cin_p = x.shape
    cout_p = dz.shape[3]
    cout, cin = dw.shape[0], dw.shape[1]
    assert dz.shape[:3] == x.shape[:3] and dw.dtype == torch.float32 and dw.is_contiguous()
    assert x.dtype == dz.dtype and x.is_contiguous() and dz.is_contiguous()
    if impl == "tc":
        assert x.dtype == torch.bfloat16
        ws = torch.empty(L.lib().fosvos_conv3x3_wgrad_tc_workspace_bytes(cin_p, cout_p), dtype=torch.uint8, device=x.device)
        L.check(L.lib().fosvos_conv3x3_wgrad_tc(x.data_ptr(), dz.data_ptr(), dw.data_ptr(), L.ptr(db), ws.data_ptr(), n, h, w,
                                                cin_p, cout_p, cin, cout, L.stream()), "conv3x3_wgrad_tc")
        return
    L.check(L.lib().fosvos_conv3x3_wgrad_simt(x.data_ptr(), dz.data_ptr(), dw.data_ptr(), L.ptr(db), n, h, w, cin_p, cout_p,
                                              cin, cout, L.dtype_code(x.dtype), L.stream()), "conv3x3_wgrad_simt")


def maxpool2x2(x: torch.Tensor) -> torch.Tensor:
    L.require_device(x.device)
    n, h, w, c = x.shape
    y = torch.empty((n, (h + 1) // 2, (w + 1) // 2, c), dtype=x.dtype, device=x.device)
    L.check(L.lib().fosvos_maxpool2x2_fwd(x.data_ptr(), y.data_ptr(), n, h, w, c, L.dtype_code(x.dtype), L.stream()), "maxpool2x2_fwd")
    return y


def maxpool2x2_bwd(x: torch.Tensor, dy: torch.Tensor, add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Gradient of the pool w.r.t. its input; with ``add`` (same shape as x) the result is written INTO it:
    add += pool gradient (the fan-in of the stage output's two consumers)."""
    L.require_device(x.device)
    n, h, w, c = x.shape
    assert dy.shape == (n, (h + 1) // 2, (w + 1) // 2, c) and dy.dtype == x.dtype
    if add is not None:
        assert add.shape == x.shape and add.dtype == x.dtype and add.is_contiguous()
    dx = torch.empty_like(x) if add is None else add
    L.check(L.lib().fosvos_maxpool2x2_bwd_add(x.data_ptr(), dy.data_ptr(), L.ptr(add), dx.data_ptr(), n, h, w, c,
                                              L.dtype_code(x.dtype), L.stream()), "maxpool2x2_bwd")
    return dx


# ---- side-output chain -----------------------------------------------------------------------
def side_params_prepare(upscale_w: Sequence[torch.Tensor], upscale1_w: Sequence[torch.Tensor], score_w: Sequence[torch.Tensor],
                        score_b: Sequence[torch.Tensor], fuse_w: torch.Tensor, fuse_b: torch.Tensor,
                        out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = fuse_w.device
    L.require_device(dev)
    if out is None:
        out = torch.empty(L.lib().fosvos_side_params_bytes() // 4, dtype=torch.float32, device=dev)
    for i in range(4):
        k = 4 << i
        assert tuple(upscale_w[i].shape) == (16, 16, k, k) and tuple(upscale1_w[i].shape) == (1, 1, k, k)
    keep = [t.detach().contiguous() for t in list(upscale_w) + list(upscale1_w) + list(score_w) + list(score_b)]
    fw, fb = fuse_w.detach().contiguous(), fuse_b.detach().contiguous()
    L.check(L.lib().fosvos_side_prepare(L.ptr_array(keep[0:4]), L.ptr_array(keep[4:8]), L.ptr_array(keep[8:12]),
                                        L.ptr_array(keep[12:16]), fw.data_ptr(), fb.data_ptr(), out.data_ptr(), L.stream()),
            "side_prepare")
    return out


def side_check_diagonal(upscale_w: Sequence[torch.Tensor]) -> torch.Tensor:
    """int32 device scalar: number of elements violating the diagonal/shared-kernel structure."""
    keep = [t.detach().contiguous() for t in upscale_w]
    flag = torch.empty(1, dtype=torch.int32, device=keep[0].device)
    L.check(L.lib().fosvos_side_check_diagonal(L.ptr_array(keep), flag.data_ptr(), L.stream()), "side_check_diagonal")
    return flag


def side_separable(params: torch.Tensor) -> bool:
    """True when both shared up-sampling kernels of a prepared parameter block factor exactly (host sync)."""
    return float(params[L.lib().fosvos_side_params_separable_flag()].item()) == 0.0


def side_fwd(sp: Sequence[torch.Tensor], params: torch.Tensor, H: int, W: int, general=False,
             want_prob: bool = False, want_mask: bool = False):
    """-> ([side0..3, fused] each (N,1,H,W) fp32, prob or None, mask or None)
    general: False/0 = shared-kernel fast path, True/1 = any up-sampling weights, 2 = separable fast path
    (only when ``side_separable(params)``)."""
    dev = params.device
    L.require_device(dev)
    n = sp[0].shape[0]
    hs = [int(t.shape[1]) for t in sp]
    ws = [int(t.shape[2]) for t in sp]
    for t in sp:
        assert t.shape[3] == 16 and t.is_contiguous() and t.dtype == sp[0].dtype
    outs = [torch.empty((n, 1, H, W), dtype=torch.float32, device=dev) for _ in range(5)]
    prob = torch.empty((n, 1, H, W), dtype=torch.float32, device=dev) if want_prob else None
    mask = torch.empty((n, 1, H, W), dtype=torch.uint8, device=dev) if want_mask else None
    ha, wa = L.int_array(hs), L.int_array(ws)
    wsb = L.lib().fosvos_side_workspace_bytes(ha, wa, n)
    workspace = None if int(general) == 1 else torch.empty(wsb, dtype=torch.uint8, device=dev)
    L.check(L.lib().fosvos_side_fwd(L.ptr_array(sp), ha, wa, params.data_ptr(), L.ptr_array(outs), L.ptr(prob), L.ptr(mask),
                                    L.ptr(workspace), int(general), n, H, W, L.dtype_code(sp[0].dtype), L.stream()), "side_fwd")
    return outs, prob, mask


def side_heads_views(params: torch.Tensor) -> List[torch.Tensor]:
    """Per stage the 36 head parameters {score_dsn.w[16], score_dsn.b, 3 unused, fuse.w[16 i ..]} inside a prepared block."""
    return [params[L.lib().fosvos_side_params_heads_offset(i):L.lib().fosvos_side_params_heads_offset(i) + 36] for i in range(4)]


def side_zs_workspace(n: int, hs: Sequence[int], ws: Sequence[int], device) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """The low-res head maps of a batch, stage after stage ((N,h_i,w_i) float2 each), and the per-stage views."""
    sizes = [n * h * w for h, w in zip(hs, ws)]
    flat = torch.empty((sum(sizes), 2), dtype=torch.float32, device=device)
    views, off = [], 0
    for sz in sizes:
        views.append(flat[off:off + sz])
        off += sz
    return flat, views


def side_fwd_heads_done(zs: torch.Tensor, hs: Sequence[int], ws: Sequence[int], params: torch.Tensor, n: int, H: int, W: int,
                        general: int = 2, want_prob: bool = False, want_mask: bool = False):
    """``side_fwd`` fast paths (general 0 / 2) from head maps the side_prep convolutions already wrote (``conv3x3_side``)."""
    dev = params.device
    L.require_device(dev)
    assert int(general) in (0, 2) and zs.dtype == torch.float32 and zs.numel() == 2 * sum(n * h * w for h, w in zip(hs, ws))
    outs = [torch.empty((n, 1, H, W), dtype=torch.float32, device=dev) for _ in range(5)]
    prob = torch.empty((n, 1, H, W), dtype=torch.float32, device=dev) if want_prob else None
    mask = torch.empty((n, 1, H, W), dtype=torch.uint8, device=dev) if want_mask else None
    L.check(L.lib().fosvos_side_fwd_heads_done(L.int_array(list(hs)), L.int_array(list(ws)), params.data_ptr(), L.ptr_array(outs),
                                               L.ptr(prob), L.ptr(mask), zs.data_ptr(), int(general), n, H, W, L.stream()), "side_fwd_heads_done")
    return outs, prob, mask


def side_bwd(sp: Sequence[torch.Tensor], params: torch.Tensor, dout: Sequence[Optional[torch.Tensor]], H: int, W: int,
             d_fuse_w: Optional[torch.Tensor], d_fuse_b: Optional[torch.Tensor],
             d_score_w: Optional[Sequence[Optional[torch.Tensor]]], d_score_b: Optional[Sequence[Optional[torch.Tensor]]]) -> List[torch.Tensor]:
    """Returns dsp[0..3]; accumulates into the given parameter-gradient tensors."""
    L.require_device(params.device)
    n = sp[0].shape[0]
    hs = [int(t.shape[1]) for t in sp]
    ws = [int(t.shape[2]) for t in sp]
    assert dout[4] is not None
    douts = [None if d is None else d.contiguous() for d in dout]
    for d in douts:
        assert d is None or (d.dtype == torch.float32 and d.numel() == n * H * W)
    dsp = [torch.empty_like(t) for t in sp]
    L.check(L.lib().fosvos_side_bwd(L.ptr_array(sp), L.int_array(hs), L.int_array(ws), params.data_ptr(), L.ptr_array(douts),
                                    L.ptr_array(dsp), L.ptr(d_fuse_w), L.ptr(d_fuse_b),
                                    L.ptr_array(d_score_w) if d_score_w is not None else None,
                                    L.ptr_array(d_score_b) if d_score_b is not None else None,
                                    n, H, W, L.dtype_code(sp[0].dtype), L.stream()), "side_bwd")
    return dsp


# ---- loss -----------------------------------------------------------------------------------
def bal_loss_fwd(output: torch.Tensor, label: torch.Tensor, size_average: bool,
                 stats: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (loss 0-dim fp32, stats float64: [0] positives, [1] negatives, [2] sum loss_pos, [3] sum loss_neg, ...)"""
    L.require_device(output.device)
    assert output.dtype == torch.float32 and label.dtype == torch.float32 and output.numel() == label.numel()
    output, label = output.contiguous(), label.contiguous()
    if stats is None:
        stats = torch.empty(L.lib().fosvos_bal_loss_stats_bytes() // 8, dtype=torch.float64, device=output.device)
    assert stats.dtype == torch.float64 and stats.numel() * 8 >= L.lib().fosvos_bal_loss_stats_bytes()
    loss = torch.empty((), dtype=torch.float32, device=output.device)
    L.check(L.lib().fosvos_bal_loss_fwd(output.data_ptr(), label.data_ptr(), output.numel(), int(size_average),
                                        stats.data_ptr(), loss.data_ptr(), L.stream()), "bal_loss_fwd")
    return loss, stats


def bal_loss_fwd_bwd(output: torch.Tensor, label: torch.Tensor, size_average: bool, stats: torch.Tensor,
                     grad_out: Optional[torch.Tensor] = None, grad_scale: float = 1.0, out: Optional[torch.Tensor] = None):
    """Loss and its gradient in one pass; ``stats`` from an earlier ``bal_loss_fwd`` on the SAME label (its counts are
    reused, its sums are overwritten).  -> (loss 0-dim fp32, dx)"""
    L.require_device(output.device)
    assert output.dtype == torch.float32 and label.dtype == torch.float32 and output.numel() == label.numel()
    output, label = output.contiguous(), label.contiguous()
    dx = torch.empty_like(output) if out is None else out
    loss = torch.empty((), dtype=torch.float32, device=output.device)
    L.check(L.lib().fosvos_bal_loss_fwd_bwd(output.data_ptr(), label.data_ptr(), output.numel(), int(size_average), stats.data_ptr(),
                                            loss.data_ptr(), L.ptr(grad_out), float(grad_scale), dx.data_ptr(), L.stream()),
            "bal_loss_fwd_bwd")
    return loss, dx


def bal_loss_fwd_bwd_frames(output: torch.Tensor, label: torch.Tensor, size_average: bool, stats: torch.Tensor,
                            grad_out: Optional[torch.Tensor] = None, grad_scale: float = 1.0,
                            out: Optional[torch.Tensor] = None, losses: Optional[torch.Tensor] = None):
    """Per-frame loss and gradient of a batch (N,1,H,W) in ONE launch; ``stats`` (N, S) float64 holds every frame's
    label statistics (from per-frame ``bal_loss_fwd`` calls).  -> (losses (N,) fp32, dx)"""
    L.require_device(output.device)
    assert output.dtype == torch.float32 and label.dtype == torch.float32 and output.shape == label.shape
    output, label = output.contiguous(), label.contiguous()
    n = output.shape[0]
    assert stats.dtype == torch.float64 and stats.dim() == 2 and stats.shape[0] == n and stats.is_contiguous()
    dx = torch.empty_like(output) if out is None else out
    if losses is None:
        losses = torch.empty(n, dtype=torch.float32, device=output.device)
    L.check(L.lib().fosvos_bal_loss_fwd_bwd_frames(output.data_ptr(), label.data_ptr(), output.numel() // n, n, int(size_average),
                                                   stats.data_ptr(), stats.shape[1], losses.data_ptr(), L.ptr(grad_out), float(grad_scale),
                                                   dx.data_ptr(), L.stream()), "bal_loss_fwd_bwd_frames")
    return losses, dx


def bal_loss_bwd(output: torch.Tensor, label: torch.Tensor, size_average: bool, stats: torch.Tensor,
                 grad_out: Optional[torch.Tensor], grad_scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    output, label = output.contiguous(), label.contiguous()
    dx = torch.empty_like(output) if out is None else out
    if grad_out is not None:
        grad_out = grad_out.to(torch.float32).contiguous()
    L.check(L.lib().fosvos_bal_loss_bwd(output.data_ptr(), label.data_ptr(), output.numel(), int(size_average),
                                        stats.data_ptr(), L.ptr(grad_out), float(grad_scale), dx.data_ptr(), L.stream()),
            "bal_loss_bwd")
    return dx


def loss_accumulate(part: torch.Tensor, total: torch.Tensor, weight: Optional[torch.Tensor] = None, init: bool = False) -> None:
    """total[i] = (0 if init else total[i]) + weight * part[i]  (weight: 0-dim device tensor or None = 1)."""
    L.require_device(total.device)
    assert part.dtype == total.dtype == torch.float32 and part.numel() == total.numel()
    L.check(L.lib().fosvos_loss_accumulate(part.data_ptr(), L.ptr(weight), total.data_ptr(), total.numel(), int(init), L.stream()), "loss_accumulate")


def loss_window_finish(total: torch.Tensor, loss_sum: torch.Tensor, last_loss: torch.Tensor) -> None:
    """loss_sum += sum(total); last_loss = total[-1]."""
    L.require_device(total.device)
    L.check(L.lib().fosvos_loss_window_finish(total.data_ptr(), total.numel(), loss_sum.data_ptr(), last_loss.data_ptr(), L.stream()),
            "loss_window_finish")


# ---- optimizer-step companions (one launch for all layers) -------------------------------------
FOLD_ENTRY = np.dtype([("ws", np.uint64), ("dw", np.uint64), ("Cout", np.int32), ("Cin", np.int32), ("CoutP", np.int32),
                       ("CinP", np.int32), ("x_is_a", np.int32), ("pad_", np.int32)])
REPACK_ENTRY = np.dtype([("w", np.uint64), ("bias", np.uint64), ("out_fwd", np.uint64), ("out_dgrad", np.uint64),
                         ("bias_out", np.uint64), ("Cout", np.int32), ("Cin", np.int32), ("pad_ci", np.int32), ("pad_co", np.int32)])


def _tile_table(tab: np.ndarray, counts: Sequence[int], device):
    prefix = np.zeros(len(counts) + 1, dtype=np.int32)
    prefix[1:] = np.cumsum(np.asarray(counts, dtype=np.int64))
    return (torch.from_numpy(tab.view(np.uint8).copy()).to(device), torch.from_numpy(prefix).to(device), len(counts), int(prefix[-1]))


def fold_table(entries: Sequence[Tuple[torch.Tensor, torch.Tensor]], device):
    """entries: (accumulator ws, OIHW fp32 gradient dw) per conv -> device table for `wgrad_fold_all`."""
    assert FOLD_ENTRY.itemsize == 40
    tab = np.zeros(len(entries), dtype=FOLD_ENTRY)
    counts = []
    for i, (ws, dw) in enumerate(entries):
        cout, cin = int(dw.shape[0]), int(dw.shape[1])
        coutp, cinp = pad8(cout), pad8(cin)
        assert ws.dtype == dw.dtype == torch.float32 and dw.is_contiguous() and ws.numel() == 9 * coutp * cinp
        tab[i] = (ws.data_ptr(), dw.data_ptr(), cout, cin, coutp, cinp, L.lib().fosvos_conv3x3_wgrad_tc_orientation(cinp, coutp), 0)
        counts.append(L.lib().fosvos_fold_tile_count(cout, cin))
    return _tile_table(tab, counts, device)


def wgrad_fold_all(table) -> None:
    t, prefix, n, tiles = table
    L.check(L.lib().fosvos_wgrad_fold_all(t.data_ptr(), n, prefix.data_ptr(), tiles, L.stream()), "wgrad_fold_all")


def repack_table(entries, device):
    """entries: (weight OIHW fp32, bias or None, packed fwd or None, packed dgrad or None, padded bias or None)."""
    assert REPACK_ENTRY.itemsize == 56
    tab = np.zeros(len(entries), dtype=REPACK_ENTRY)
    counts = []
    for i, (w, b, fwd, dgr, bo) in enumerate(entries):
        cout, cin = int(w.shape[0]), int(w.shape[1])
        coutp, cinp = pad8(cout), pad8(cin)
        pad_ci, pad_co = (cinp + 63) // 64 * 64, (coutp + 63) // 64 * 64
        assert w.dtype == torch.float32 and w.is_contiguous()
        assert fwd is None or (fwd.dtype == torch.bfloat16 and fwd.numel() == 9 * coutp * pad_ci)
        assert dgr is None or (dgr.dtype == torch.bfloat16 and dgr.numel() == 9 * cinp * pad_co)
        tab[i] = (w.data_ptr(), 0 if b is None else b.data_ptr(), 0 if fwd is None else fwd.data_ptr(),
                  0 if dgr is None else dgr.data_ptr(), 0 if bo is None else bo.data_ptr(), cout, cin, pad_ci, pad_co)
        counts.append(L.lib().fosvos_repack_tile_count(cout, cin))
    return _tile_table(tab, counts, device)


CONVSTEP_ENTRY = np.dtype([("ws", np.uint64), ("dw", np.uint64), ("w", np.uint64), ("buf", np.uint64), ("bias", np.uint64),
                           ("out_fwd", np.uint64), ("out_dgrad", np.uint64), ("bias_out", np.uint64),
                           ("Cout", np.int32), ("Cin", np.int32), ("CoutP", np.int32), ("CinP", np.int32),
                           ("x_is_a", np.int32), ("pad_ci", np.int32), ("pad_co", np.int32), ("pad_", np.int32),
                           ("lr", np.float32), ("wd", np.float32)])


def convstep_table(entries, device, out=None):
    """entries: (ws, dw or None, weight, momentum buffer, bias or None, packed fwd, packed dgrad or None, padded bias, lr, wd)
    per conv -> device table for ``conv_step_all``.  ``out``: an existing table of the same geometry to rewrite in place."""
    assert CONVSTEP_ENTRY.itemsize == 104
    tab = np.zeros(len(entries), dtype=CONVSTEP_ENTRY)
    counts = []
    for i, (ws, dw, w, buf, b, fwd, dgr, bo, lr, wd) in enumerate(entries):
        cout, cin = int(w.shape[0]), int(w.shape[1])
        coutp, cinp = pad8(cout), pad8(cin)
        pad_ci, pad_co = (cinp + 63) // 64 * 64, (coutp + 63) // 64 * 64
        assert ws.dtype == w.dtype == buf.dtype == torch.float32 and w.is_contiguous() and buf.is_contiguous() and ws.numel() == 9 * coutp * cinp
        assert fwd is None or (fwd.dtype == torch.bfloat16 and fwd.numel() == 9 * coutp * pad_ci)
        assert dgr is None or (dgr.dtype == torch.bfloat16 and dgr.numel() == 9 * cinp * pad_co)
        tab[i] = (ws.data_ptr(), 0 if dw is None else dw.data_ptr(), w.data_ptr(), buf.data_ptr(), 0 if b is None else b.data_ptr(),
                  0 if fwd is None else fwd.data_ptr(), 0 if dgr is None else dgr.data_ptr(), 0 if bo is None else bo.data_ptr(),
                  cout, cin, coutp, cinp, L.lib().fosvos_conv3x3_wgrad_tc_orientation(cinp, coutp), pad_ci, pad_co, 0, lr, wd)
        counts.append(L.lib().fosvos_repack_tile_count(cout, cin))
    if out is not None:
        t, prefix, n, tiles = out
        new = torch.from_numpy(tab.view(np.uint8).copy())
        if t.numel() == new.numel() and n == len(counts) and tiles == int(sum(counts)):
            t.copy_(new)
            return out
    return _tile_table(tab, counts, device)


def conv_step_all(table, momentum: float) -> None:
    t, prefix, n, tiles = table
    L.check(L.lib().fosvos_conv_step_all(t.data_ptr(), n, prefix.data_ptr(), tiles, float(momentum), L.stream()), "conv_step_all")


def repack_all(table) -> None:
    t, prefix, n, tiles = table
    L.check(L.lib().fosvos_repack_all(t.data_ptr(), n, prefix.data_ptr(), tiles, L.stream()), "repack_all")


# ---- optimizer ------------------------------------------------------------------------------
SGD_ENTRY = np.dtype([("p", np.uint64), ("g", np.uint64), ("buf", np.uint64), ("n", np.int64), ("lr", np.float32),
                      ("wd", np.float32)])


def sgd_table(entries: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, float, float]], device) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """entries: (param, grad, momentum_buffer, lr, weight_decay) -> (table, chunk_prefix, n_chunks) on device."""
    chunk = L.lib().fosvos_sgd_chunk_elems()
    tab = np.zeros(len(entries), dtype=SGD_ENTRY)
    prefix = np.zeros(len(entries) + 1, dtype=np.int64)
    for i, (p, g, b, lr, wd) in enumerate(entries):
        assert p.dtype == g.dtype == b.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous() and b.is_contiguous()
        tab[i] = (p.data_ptr(), g.data_ptr(), b.data_ptr(), p.numel(), lr, wd)
        prefix[i + 1] = prefix[i] + (p.numel() + chunk - 1) // chunk
    assert SGD_ENTRY.itemsize == 40
    t = torch.from_numpy(tab.view(np.uint8).copy()).to(device)
    pf = torch.from_numpy(prefix).to(device)
    return t, pf, int(prefix[-1])


def sgd_step(table: torch.Tensor, prefix: torch.Tensor, n_tensors: int, n_chunks: int, momentum: float, zero_grad: bool) -> None:
    L.check(L.lib().fosvos_sgd_step(table.data_ptr(), n_tensors, prefix.data_ptr(), n_chunks, float(momentum), int(zero_grad),
                                    L.stream()), "sgd_step")


# ---- Adam / distillation losses / pruning criterion / ingest ----------------------------
ADAM_ENTRY = np.dtype([("p", np.uint64), ("g", np.uint64), ("m", np.uint64), ("v", np.uint64), ("n", np.int64),
                       ("lr", np.float32), ("wd", np.float32)])


def adam_table(entries, device):
    """entries: (param, grad, exp_avg, exp_avg_sq, lr, weight_decay) -> (table, chunk_prefix, n_chunks) on device."""
    chunk = L.lib().fosvos_adam_chunk_elems()
    assert ADAM_ENTRY.itemsize == 48
    tab = np.zeros(len(entries), dtype=ADAM_ENTRY)
    prefix = np.zeros(len(entries) + 1, dtype=np.int64)
    for i, (p, g, m, v, lr, wd) in enumerate(entries):
        assert p.dtype == g.dtype == m.dtype == v.dtype == torch.float32
        assert p.is_contiguous() and g.is_contiguous() and m.is_contiguous() and v.is_contiguous()
        tab[i] = (p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, wd)
        prefix[i + 1] = prefix[i] + (p.numel() + chunk - 1) // chunk
    return torch.from_numpy(tab.view(np.uint8).copy()).to(device), torch.from_numpy(prefix).to(device), int(prefix[-1])


def adam_step(table: torch.Tensor, prefix: torch.Tensor, n_tensors: int, n_chunks: int, beta1: float, beta2: float, eps: float,
              state: torch.Tensor, zero_grad: bool) -> None:
    assert state.dtype == torch.int64 and state.numel() >= 1
    L.check(L.lib().fosvos_adam_step(table.data_ptr(), n_tensors, prefix.data_ptr(), n_chunks, float(beta1), float(beta2), float(eps),
                                     state.data_ptr(), int(zero_grad), L.stream()), "adam_step")


def pixel_loss(output: torch.Tensor, target: torch.Tensor, kind: str, size_average: bool, grad_scale: float = 1.0,
               want_grad: bool = True):
    """nn.MSELoss ('mse') / nn.L1Loss ('l1') -> (loss 0-dim fp32, d loss / d output * grad_scale or None)"""
    L.require_device(output.device)
    assert output.dtype == torch.float32 and target.dtype == torch.float32 and output.numel() == target.numel()
    output, target = output.contiguous(), target.contiguous()
    loss = torch.empty((), dtype=torch.float32, device=output.device)
    dx = torch.empty_like(output) if want_grad else None
    L.check(L.lib().fosvos_pixel_loss(output.data_ptr(), target.data_ptr(), output.numel(), {"mse": 0, "l1": 1}[kind],
                                      int(size_average), float(grad_scale), loss.data_ptr(), L.ptr(dx), L.stream()), "pixel_loss")
    return loss, dx


def relu_fwd(x: torch.Tensor) -> torch.Tensor:
    L.require_device(x.device)
    assert x.dtype == torch.float32
    x = x.contiguous()
    y = torch.empty_like(x)
    L.check(L.lib().fosvos_relu_fwd(x.data_ptr(), y.data_ptr(), x.numel(), L.stream()), "relu_fwd")
    return y


def relu_bwd(y: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    L.require_device(y.device)
    assert y.dtype == torch.float32 and dy.dtype == torch.float32 and y.numel() == dy.numel()
    y, dy = y.contiguous(), dy.contiguous()
    dx = torch.empty_like(y)
    L.check(L.lib().fosvos_relu_bwd(y.data_ptr(), dy.data_ptr(), dx.data_ptr(), y.numel(), L.stream()), "relu_bwd")
    return dx


def taylor_rank(act: torch.Tensor, grad: torch.Tensor, rank: torch.Tensor) -> None:
    """rank (C fp32) += sum over pixels of act * grad / (N H W);  act, grad NHWC of the same dtype."""
    L.require_device(act.device)
    n, h, w, cp = act.shape
    assert grad.shape == act.shape and grad.dtype == act.dtype and act.is_contiguous() and grad.is_contiguous()
    assert rank.dtype == torch.float32 and rank.numel() <= cp
    L.check(L.lib().fosvos_taylor_rank(act.data_ptr(), grad.data_ptr(), rank.data_ptr(), n, h, w, cp, rank.numel(),
                                       L.dtype_code(act.dtype), L.stream()), "taylor_rank")


def ingest_u8(img: torch.Tensor, mean, dtype: torch.dtype) -> torch.Tensor:
    """uint8 (N,H,W,3) device frames (cv2 channel order) -> NHWC8 activations, mean-subtracted."""
    import ctypes
    L.require_device(img.device)
    assert img.dtype == torch.uint8 and img.dim() == 4 and img.shape[3] == 3
    img = img.contiguous()
    n, h, w, _ = img.shape
    y = torch.empty((n, h, w, 8), dtype=dtype, device=img.device)
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    L.check(L.lib().fosvos_ingest_u8(img.data_ptr(), y.data_ptr(), n, h, w, m, L.dtype_code(dtype), L.stream()), "ingest_u8")
    return y


# ---- mask egress ------------------------------------------------------------------------------
def sigmoid_threshold(logits: torch.Tensor, want_prob: bool = True, want_mask: bool = True):
    L.require_device(logits.device)
    logits = logits.contiguous()
    prob = torch.empty_like(logits) if want_prob else None
    mask = torch.empty(logits.shape, dtype=torch.uint8, device=logits.device) if want_mask else None
    L.check(L.lib().fosvos_sigmoid_threshold(logits.data_ptr(), L.ptr(prob), L.ptr(mask), logits.numel(), L.stream()), "sigmoid_threshold")
    return prob, mask


def mask_iou(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a, b: uint8 {0,1} masks (F, ...) -> int64 (F, 2): intersection and union pixel counts per frame."""
    L.require_device(a.device)
    assert a.dtype == torch.uint8 and b.dtype == torch.uint8 and a.shape == b.shape
    a, b = a.contiguous(), b.contiguous()
    f = a.shape[0]
    counts = torch.empty((f, 2), dtype=torch.int64, device=a.device)
    L.check(L.lib().fosvos_mask_iou(a.data_ptr(), b.data_ptr(), a.numel() // f, f, counts.data_ptr(), L.stream()), "mask_iou")
    return counts


# every launching wrapper runs with its tensors' device current (see _lib.on_tensor_device)
for _name in ("nchw_to_nhwc", "nhwc_to_nchw", "pack_weight", "pad_bias", "split_frames", "pack_weight_split", "conv3x3_split", "maxpool2x2_split", "conv3x3", "conv3x3_pool", "conv3x3_pool_only", "conv3x3_pool_arg", "maxpool2x2_bwd_arg", "conv3x3_side",
              "conv3x3_wgrad_accumulate", "conv3x3_wgrad_finish", "conv3x3_wgrad", "maxpool2x2", "maxpool2x2_bwd", "side_params_prepare",
              "side_check_diagonal", "side_fwd", "side_fwd_heads_done", "side_bwd", "bal_loss_fwd", "bal_loss_fwd_bwd",
              "bal_loss_fwd_bwd_frames", "bal_loss_bwd", "loss_accumulate", "loss_window_finish", "wgrad_fold_all", "repack_all", "sgd_step", "adam_step", "pixel_loss", "taylor_rank", "relu_fwd", "relu_bwd",
              "ingest_u8", "sigmoid_threshold", "mask_iou"):
    globals()[_name] = L.on_tensor_device(globals()[_name])
del _name
