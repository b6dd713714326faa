"""GPU: every kernel behind the C ABI against the CPU oracle primitives on seeded inputs."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from fosvos_b200 import _lib as L, ops  # noqa: E402
from oracle import osvos_oracle as O  # noqa: E402

DEV = "cuda"


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _bf16r(t):
    return t.to(torch.bfloat16).float()


def _nhwc(x_nchw, dtype):
    """CPU NCHW fp32 -> device NHWC (padded to 8 channels)."""
    return ops.nchw_to_nhwc(x_nchw.to(DEV), dtype)


def _nchw(x_nhwc, c):
    return ops.nhwc_to_nchw(x_nhwc, c).cpu()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 3, 5, 7), (2, 16, 9, 4), (1, 20, 3, 3)])
def test_layout_roundtrip(dtype, shape):
    x = torch.randn(shape, generator=_gen(0))
    y = _nhwc(x, dtype)
    assert y.shape == (shape[0], shape[2], shape[3], ops.pad8(shape[1]))
    ref = x if dtype == torch.float32 else _bf16r(x)
    assert torch.equal(y[..., :shape[1]].float().cpu().permute(0, 3, 1, 2), ref)
    assert (y[..., shape[1]:] == 0).all()
    assert torch.equal(_nchw(y, shape[1]), ref)


@pytest.mark.parametrize("cout,cin", [(64, 3), (16, 128), (20, 12)])
def test_pack_weight_layouts(cout, cin):
    w = torch.randn(cout, cin, 3, 3, generator=_gen(1))
    wd = w.to(DEV)
    cop, cip = ops.pad8(cout), ops.pad8(cin)
    wp = torch.zeros(cop, cip, 3, 3)
    wp[:cout, :cin] = w
    taps = wp.reshape(cop, cip, 9)
    simt_f = ops.pack_weight(wd, L.W_SIMT_FWD, torch.float32).cpu().view(9, cip, cop)
    assert torch.equal(simt_f, taps.permute(2, 1, 0))
    simt_d = ops.pack_weight(wd, L.W_SIMT_DGRAD, torch.float32).cpu().view(9, cop, cip)
    assert torch.equal(simt_d, taps.flip(2).permute(2, 0, 1))
    pad = (cip + 63) // 64 * 64
    tc_f = ops.pack_weight(wd, L.W_TC_FWD, torch.bfloat16).float().cpu().view(cop, 9, pad)
    assert torch.equal(tc_f[:, :, :cip], _bf16r(taps.permute(0, 2, 1))) and (tc_f[:, :, cip:] == 0).all()
    padc = (cop + 63) // 64 * 64
    tc_d = ops.pack_weight(wd, L.W_TC_DGRAD, torch.bfloat16).float().cpu().view(cip, 9, padc)
    assert torch.equal(tc_d[:, :, :cop], _bf16r(taps.flip(2).permute(1, 2, 0))) and (tc_d[:, :, cop:] == 0).all()


CONV_CASES = [
    # N, H, W, Cin, Cout
    (1, 9, 11, 3, 64),
    (2, 16, 16, 64, 64),
    (1, 17, 23, 64, 128),
    (1, 8, 40, 128, 16),
    (1, 30, 54, 24, 40),
    (1, 5, 3, 256, 256),
    # narrow-N tiles take two 64-channel K blocks per pipeline stage: odd block counts end on a short stage
    (1, 12, 20, 64, 16),
    (1, 10, 18, 192, 32),
    (2, 14, 9, 32, 32),
    # 16 input channels and a wide output: MODE_K16 (one K = 16 MMA per tap, SWIZZLE_32B operands) on the tensor-core path
    (1, 17, 23, 16, 128),
    (2, 9, 40, 16, 256),
    (1, 30, 54, 16, 512),
    (1, 12, 11, 16, 40),
]


def _conv_ref(x, w, b, relu):
    y = F.conv2d(x.double(), w.double(), None if b is None else b.double(), padding=1)
    return (F.relu(y) if relu else y).float()


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("mode", ["fp32_simt", "bf16_simt", "bf16_tc"])
def test_conv3x3_forward(case, mode):
    n, h, w_, cin, cout = case
    g = _gen(hash(case) % 1000)
    x = torch.randn(n, cin, h, w_, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g)
    dt = torch.float32 if mode == "fp32_simt" else torch.bfloat16
    impl = "tc" if mode == "bf16_tc" else "simt"
    xd = _nhwc(x, dt)
    wp = ops.pack_weight(w.to(DEV), L.W_TC_FWD if impl == "tc" else L.W_SIMT_FWD, dt)
    bp = ops.pad_bias(b.to(DEV), cout, DEV)
    y = ops.conv3x3(xd, wp, bp, ops.pad8(cout), L.CONV_BIAS | L.CONV_RELU, impl=impl)
    torch.cuda.synchronize()
    got = _nchw(y, cout)
    if dt == torch.float32:
        ref = _conv_ref(x, w, b, True)
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), float((got - ref).abs().max())
    else:
        ref = _conv_ref(_bf16r(x), _bf16r(w), b, True)     # same rounded operands, exact accumulation
        # one bf16 rounding of the result (2^-9 relative) + fp32 accumulation order
        assert torch.allclose(got, ref, rtol=2 ** -7, atol=2e-2), float((got - ref).abs().max())
    assert (y[..., cout:] == 0).all()


@pytest.mark.parametrize("case", [(1, 33, 45, 64, 64), (2, 16, 24, 64, 128), (1, 31, 70, 128, 256), (1, 9, 17, 40, 72)])
def test_conv3x3_fused_maxpool(case):
    """The producing conv's epilogue also writes MaxPool2d(2, 2, ceil_mode=True) of its ReLU output (odd sizes: the
    last window is one pixel wide / high): bit-identical to the separate pool kernel."""
    n, h, w_, cin, cout = case
    g = _gen(51)
    x = _nhwc(torch.randn(n, cin, h, w_, generator=g), torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, generator=g).to(DEV) * 0.05
    b = torch.randn(cout, generator=g).to(DEV)
    wp = ops.pack_weight(w, L.W_TC_FWD, torch.bfloat16)
    bp = ops.pad_bias(b, cout, DEV)
    cp = ops.pad8(cout)
    y_ref = ops.conv3x3(x, wp, bp, cp, L.CONV_BIAS | L.CONV_RELU)
    p_ref = ops.maxpool2x2(y_ref)
    y, yp = ops.conv3x3_pool(x, wp, bp, cp, L.CONV_BIAS | L.CONV_RELU)
    assert torch.equal(y, y_ref)
    assert torch.equal(yp, p_ref)


def _unsplit(y, terms, c):
    """split map (N,H,W, terms*seg) -> fp32 NCHW of the first c channels: the sum of the bf16 terms"""
    seg = y.shape[3] // terms
    return sum(y[..., t * seg:t * seg + c].float() for t in range(terms)).permute(0, 3, 1, 2).cpu()


@pytest.mark.parametrize("terms", [3, 2])
@pytest.mark.parametrize("case", [(1, 9, 11, 3, 64), (2, 12, 10, 64, 128), (1, 17, 23, 128, 256), (1, 8, 8, 72, 40), (1, 7, 9, 128, 16), (2, 5, 30, 512, 16)])
def test_conv3x3_split_operands(case, terms):
    """fp32 through the bf16 tensor cores (csrc/split.cu, conv_tc.cu SPLIT): three bf16 terms per operand and six term
    products reproduce the fp32 convolution (the reference's strict mode) to fp32 rounding; two terms / three products
    hold ~2^-16.  Cout > 32 returns a split map, Cout <= 32 (side_prep) plain fp32."""
    n, h, w_, cin, cout = case
    g = _gen(hash(case) % 1000 + 3)
    x = torch.randn(n, cin, h, w_, generator=g) * 30.0
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g)
    relu = cout > 32
    ref = _conv_ref(x, w, b, relu)
    xs = ops.split_frames(x.to(DEV), terms)
    assert xs.shape == (n, h, w_, terms * ops.seg64(cin))
    assert torch.allclose(_unsplit(xs, terms, cin), x, rtol=2.0 ** -23 if terms == 3 else 2.0 ** -15, atol=0)
    wp = ops.pack_weight_split(w.to(DEV), terms)
    bp = ops.pad_bias(b.to(DEV), cout, DEV)
    y = ops.conv3x3_split(xs, wp, bp, ops.pad8(cout), terms, L.CONV_BIAS | (L.CONV_RELU if relu else 0))
    torch.cuda.synchronize()
    got = _unsplit(y, terms, cout) if cout > 32 else y[..., :cout].permute(0, 3, 1, 2).cpu()
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max())
    # three terms: what is left is the tensor core's chopped fp32 accumulation (~2^-25 per MMA of the longest accumulator
    # chain: the kernel spreads a tile over four accumulators and adds them in the epilogue)
    print(f"split conv {case} terms {terms}: max err {err:.3e} = {err / scale:.2e} of the map scale")
    assert err <= (8e-6 if terms == 3 else 1e-4) * scale, (err, scale)
    if cout > 32:
        assert (y.view(n, h, w_, terms, -1)[..., cout:] == 0).all()       # padded lanes of every term stay zero


@pytest.mark.parametrize("terms", [3, 2])
@pytest.mark.parametrize("shape", [(2, 64, 9, 7), (1, 40, 30, 54), (1, 130, 5, 1)])
def test_maxpool2x2_split(shape, terms):
    n, c, h, w_ = shape
    x = torch.randn(shape, generator=_gen(61)) * 10
    xs = ops.split_frames(x.to(DEV), terms)
    ys = ops.maxpool2x2_split(xs, terms)
    x_rep = _unsplit(xs, terms, c)                                         # what the split map holds (exact for three terms)
    ref = F.max_pool2d(x_rep, 2, 2, ceil_mode=True)
    got = _unsplit(ys, terms, c)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, rtol=0 if terms == 3 else 2.0 ** -15, atol=0), float((got - ref).abs().max())


@pytest.mark.parametrize("case", [(1, 33, 45, 64, 64), (2, 16, 24, 64, 128), (1, 9, 17, 40, 72)])
def test_conv3x3_pool_only(case):
    """Inference form of the fused pool: the full-resolution output is not written at all (y = NULL); the pooled map is
    bit-identical to the one written next to the full output."""
    n, h, w_, cin, cout = case
    g = _gen(52)
    x = _nhwc(torch.randn(n, cin, h, w_, generator=g), torch.bfloat16)
    wp = ops.pack_weight(torch.randn(cout, cin, 3, 3, generator=g).to(DEV) * 0.05, L.W_TC_FWD, torch.bfloat16)
    bp = ops.pad_bias(torch.randn(cout, generator=g).to(DEV), cout, DEV)
    cp = ops.pad8(cout)
    _, p_ref = ops.conv3x3_pool(x, wp, bp, cp, L.CONV_BIAS | L.CONV_RELU)
    yp = ops.conv3x3_pool_only(x, wp, bp, cp, L.CONV_BIAS | L.CONV_RELU)
    assert torch.equal(yp, p_ref)


STACK_CASES = [
    # N, H, W, Cin (Cout = 64): tile = 14 x 8 outputs; widths / heights around the tile edges, odd sizes for the ceil-mode pool
    (1, 33, 45, 64), (2, 16, 28, 64), (1, 9, 13, 64), (1, 8, 15, 128), (3, 17, 29, 128), (1, 1, 1, 64), (1, 2, 14, 64), (1, 25, 57, 64),
]


@pytest.mark.parametrize("case", STACK_CASES)
def test_conv3x3_row_stack_forward_and_pool(case, monkeypatch):
    """Cout = 64 layers run the row-stacked kernel (conv_stack_tc.cu: three taps per N = 192 MMA, partial sums combined by
    shuffles, pool fused with in-warp shuffles): against torch on the same rounded operands, and against the generic
    tensor-core kernel (same bf16 operands, fp32 accumulation in a different order: equal up to one bf16 ulp)."""
    n, h, w_, cin = case
    g = _gen(61)
    x = torch.randn(n, cin, h, w_, generator=g)
    w = torch.randn(64, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(64, generator=g)
    xd = _nhwc(x, torch.bfloat16)
    wp = ops.pack_weight(w.to(DEV), L.W_TC_FWD, torch.bfloat16)
    bp = ops.pad_bias(b.to(DEV), 64, DEV)
    for flags, relu in ((L.CONV_BIAS | L.CONV_RELU, True), (L.CONV_BIAS, False), (0, False)):
        y = ops.conv3x3(xd, wp, bp if flags & L.CONV_BIAS else None, 64, flags)
        ref = _conv_ref(_bf16r(x), _bf16r(w), b if flags & L.CONV_BIAS else None, relu)
        got = _nchw(y, 64)
        assert torch.allclose(got, ref, rtol=2 ** -7, atol=2e-2), float((got - ref).abs().max())
    flags = L.CONV_BIAS | L.CONV_RELU
    y = ops.conv3x3(xd, wp, bp, 64, flags)
    y2, yp = ops.conv3x3_pool(xd, wp, bp, 64, flags)
    yp_only = ops.conv3x3_pool_only(xd, wp, bp, 64, flags)
    assert torch.equal(y2, y)
    assert torch.equal(yp, ops.maxpool2x2(y))
    assert torch.equal(yp_only, yp)
    monkeypatch.setenv("FOSVOS_TC_NO_STACK", "1")
    y_gen = ops.conv3x3(xd, wp, bp, 64, flags)
    monkeypatch.delenv("FOSVOS_TC_NO_STACK")
    d = (y.float() - y_gen.float()).abs()
    assert float((d / y_gen.float().abs().clamp_min(1e-2)).max()) <= 2 ** -7, float(d.max())


@pytest.mark.parametrize("case", STACK_CASES)
def test_conv3x3_row_stack_dgrad(case, monkeypatch):
    """Data-gradient use of the row-stacked kernel (64 channels out): flipped / transposed weights, ReLU mask.
    (By default only dZ with >= 128 channels takes it; the environment switch sends the 64-channel cases there too.)"""
    monkeypatch.setenv("FOSVOS_TC_STACK_ALL", "1")
    n, h, w_, cz = case                    # cz = channels of dZ (the forward conv's Cout); the forward Cin is 64
    g = _gen(62)
    xprev = torch.randn(n, 64, h, w_, generator=g).clamp_min(0)
    wt = torch.randn(cz, 64, 3, 3, generator=g) * 0.05
    dz = torch.randn(n, cz, h, w_, generator=g)
    ref = F.conv_transpose2d(_bf16r(dz).double(), _bf16r(wt).double(), padding=1).float() * (_bf16r(xprev) > 0)
    wp = ops.pack_weight(wt.to(DEV), L.W_TC_DGRAD, torch.bfloat16)
    out = ops.conv3x3(_nhwc(dz, torch.bfloat16), wp, None, 64, L.CONV_MASK, mask=_nhwc(xprev, torch.bfloat16))
    got = _nchw(out, 64)
    assert torch.allclose(got, ref, rtol=2 ** -6, atol=5e-2), float((got - ref).abs().max())
    assert (got[_bf16r(xprev) <= 0] == 0).all()


def _window_argmax(y):
    """(n, h, w, c) non-negative map -> (n, ph, pw, c) index 2 dy + dx of the FIRST maximum of each 2x2 ceil-mode window"""
    n, h, w_, c = y.shape
    ph, pw = (h + 1) // 2, (w_ + 1) // 2
    pad = torch.full((n, 2 * ph, 2 * pw, c), -1.0, device=y.device)
    pad[:, :h, :w_] = y.float()
    cand = torch.stack([pad[:, 0::2, 0::2], pad[:, 0::2, 1::2], pad[:, 1::2, 0::2], pad[:, 1::2, 1::2]], dim=-1)
    return cand.argmax(dim=-1)


def _unpack_arg(arg, c):
    """(n, ph, pw, c / 32, 2) int32 bit planes -> (n, ph, pw, c) indices"""
    bits = torch.arange(32, device=arg.device)
    lo = (arg[..., 0].unsqueeze(-1) >> bits) & 1
    hi = (arg[..., 1].unsqueeze(-1) >> bits) & 1
    return (lo + 2 * hi).reshape(*arg.shape[:3], c)


@pytest.mark.parametrize("case", [(1, 33, 45, 64, 64), (2, 16, 28, 64, 64), (2, 16, 24, 64, 128), (1, 31, 70, 128, 256), (1, 9, 13, 128, 64),
                                  (1, 1, 1, 64, 64), (1, 7, 30, 256, 512)])
def test_conv3x3_pool_arg_and_backward(case):
    """Training form of the fused pool (both tensor-core kernels): same y / pooled map as the plain fused pool, window indices
    = the first maximum in scan order (sparse post-ReLU maps: many ties at zero), and the pool gradient routed by those
    indices == the one found by re-reading the activation, with and without fan-in."""
    n, h, w_, cin, cout = case
    g = _gen(71)
    x = _nhwc(torch.randn(n, cin, h, w_, generator=g), torch.bfloat16)
    wp = ops.pack_weight(torch.randn(cout, cin, 3, 3, generator=g).to(DEV) * 0.05, L.W_TC_FWD, torch.bfloat16)
    bp = ops.pad_bias(torch.randn(cout, generator=g).to(DEV) - 0.5, cout, DEV)          # negative shift: ~70 % zeros after ReLU
    fl = L.CONV_BIAS | L.CONV_RELU
    y_ref, p_ref = ops.conv3x3_pool(x, wp, bp, cout, fl)
    y, yp, arg = ops.conv3x3_pool_arg(x, wp, bp, cout, fl)
    assert torch.equal(y, y_ref) and torch.equal(yp, p_ref)
    y0, yp0, arg0 = ops.conv3x3_pool_arg(x, wp, bp, cout, fl, want_y=False)
    assert y0 is None and torch.equal(yp0, p_ref) and torch.equal(arg0, arg)
    assert torch.equal(_unpack_arg(arg, cout), _window_argmax(y_ref))
    dy = _nhwc(torch.randn(n, cout, (h + 1) // 2, (w_ + 1) // 2, generator=g), torch.bfloat16)
    assert torch.equal(ops.maxpool2x2_bwd_arg(arg, dy, h, w_), ops.maxpool2x2_bwd(y_ref, dy))
    add = _nhwc(torch.randn(n, cout, h, w_, generator=g), torch.bfloat16)
    a1, a2 = add.clone(), add.clone()
    assert torch.equal(ops.maxpool2x2_bwd_arg(arg, dy, h, w_, add=a1), ops.maxpool2x2_bwd(y_ref, dy, add=a2))


SIDE_CASES = [
    # N, H, W, Cin: tile = 30 x 4 outputs; widths around the tile edge, one-pixel maps, ragged channel counts (pruned nets)
    (1, 30, 54, 512), (2, 15, 27, 512), (1, 45, 70, 128), (3, 7, 5, 64), (1, 4, 30, 256), (1, 5, 31, 128), (1, 3, 29, 64),
    (1, 1, 1, 64), (2, 9, 61, 104), (1, 12, 33, 520), (1, 8, 16, 8),
]


@pytest.mark.parametrize("case", SIDE_CASES)
@pytest.mark.parametrize("heads", [False, True])
def test_conv3x3_side_tc(case, heads):
    """side_prep (C -> 16, bias, no ReLU; osvos_vgg.py:42,69) through the row-stacked kernel (conv_side_tc.cu) against exact
    accumulation over the same bf16 operands, and against the generic tcgen05 kernel; with ``heads`` the two 1x1 heads
    (score_dsn, this stage's fuse columns; osvos_vgg.py:75,81) come out of the same epilogue from the fp32 accumulators."""
    n, h, w_, cin = case
    g = _gen(hash(case) % 1000 + 7)
    x = torch.randn(n, cin, h, w_, generator=g)
    w = torch.randn(16, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(16, generator=g)
    xd = _nhwc(x, torch.bfloat16)
    assert ops.side_tc_supported(xd.shape[3])
    wp = ops.pack_weight(w.to(DEV), L.W_TC_FWD, torch.bfloat16)
    bp = ops.pad_bias(b.to(DEV), 16, DEV)
    ref = _conv_ref(_bf16r(x), _bf16r(w), b, False)                       # fp32 result of exact accumulation
    zs = hp = None
    if heads:
        hp = torch.randn(36, generator=g).to(DEV)                          # score.w[16], score.b, 3 unused, fuse.w[16]
        zs = torch.full((n * h * w_, 2), float("nan"), device=DEV)
    y = ops.conv3x3_side(xd, wp, bp, zs=zs, heads=hp)
    torch.cuda.synchronize()
    got = _nchw(y, 16)
    assert torch.allclose(got, ref, rtol=2 ** -7, atol=2e-2), float((got - ref).abs().max())
    y_gen = ops.conv3x3(xd, wp, bp, 16, L.CONV_BIAS)                       # same operands, other summation order
    assert torch.allclose(y.float(), y_gen.float(), rtol=2 ** -7, atol=1e-2)
    if heads:
        hpc = hp.cpu()
        z_ref = (ref * hpc[20:36].view(1, 16, 1, 1)).sum(1)
        s_ref = (ref * hpc[0:16].view(1, 16, 1, 1)).sum(1) + hpc[16]
        got_z = zs.cpu().view(n, h, w_, 2)
        assert torch.allclose(got_z[..., 0], z_ref, rtol=1e-4, atol=1e-3), float((got_z[..., 0] - z_ref).abs().max())
        assert torch.allclose(got_z[..., 1], s_ref, rtol=1e-4, atol=1e-3), float((got_z[..., 1] - s_ref).abs().max())
        # heads only (inference): the 16-channel map itself is not written
        zs2 = torch.full_like(zs, float("nan"))
        assert ops.conv3x3_side(xd, wp, bp, zs=zs2, heads=hp, want_y=False) is None
        assert torch.equal(zs2, zs)


def test_side_fwd_from_fused_heads_matches_heads_kernel():
    """fosvos_side_fwd_heads_done on head maps written by the conv epilogue == fosvos_side_fwd on the 16-channel maps
    (up to the bf16 rounding of those maps, which the fused path skips)."""
    g = _gen(77)
    n, H, W = 2, 45, 70
    hs, ws, h_, w_ = [], [], H, W
    for _ in range(4):
        h_, w_ = (h_ + 1) // 2, (w_ + 1) // 2
        hs.append(h_); ws.append(w_)
    up = [O.interp_surgery_weight(16, 4 << i).to(DEV) for i in range(4)]
    up1 = [O.interp_surgery_weight(1, 4 << i).to(DEV) for i in range(4)]
    sw = [(torch.randn(1, 16, 1, 1, generator=g) * 0.2).to(DEV) for _ in range(4)]
    sb = [torch.randn(1, generator=g).to(DEV) for _ in range(4)]
    fw, fb = (torch.randn(1, 64, 1, 1, generator=g) * 0.1).to(DEV), torch.randn(1, generator=g).to(DEV)
    params = ops.side_params_prepare(up, up1, sw, sb, fw, fb)
    assert ops.side_separable(params)
    heads = ops.side_heads_views(params)
    zs_flat, zs_views = ops.side_zs_workspace(n, hs, ws, DEV)
    sps = []
    for i, cin in enumerate([128, 256, 512, 512]):
        x = _nhwc(torch.randn(n, cin, hs[i], ws[i], generator=g), torch.bfloat16)
        wp = ops.pack_weight((torch.randn(16, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5).to(DEV), L.W_TC_FWD, torch.bfloat16)
        bp = ops.pad_bias(torch.randn(16, generator=g).to(DEV), 16, DEV)
        sps.append(ops.conv3x3_side(x, wp, bp, zs=zs_views[i], heads=heads[i]))
    outs_a, prob_a, mask_a = ops.side_fwd(sps, params, H, W, general=2, want_prob=True, want_mask=True)
    outs_b, prob_b, mask_b = ops.side_fwd_heads_done(zs_flat, hs, ws, params, n, H, W, general=2, want_prob=True, want_mask=True)
    for a, b in zip(outs_a, outs_b):
        assert torch.allclose(a, b, rtol=1e-2, atol=3e-2), float((a - b).abs().max())
    assert torch.equal(mask_b.cpu(), O.binarise(prob_b.cpu()))
    outs_c, _, _ = ops.side_fwd_heads_done(zs_flat, hs, ws, params, n, H, W, general=0)
    for b, c in zip(outs_b, outs_c):
        assert torch.allclose(b, c, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape", [(1, 13, 18, 64, 128), (2, 13, 18, 128, 16), (1, 30, 54, 512, 16), (1, 9, 35, 40, 16)])
@pytest.mark.parametrize("impl,dt", [("simt", torch.float32), ("tc", torch.bfloat16)])
def test_conv3x3_mask_accumulate_dgrad(impl, dt, shape):
    """The data-gradient use of the kernel: flipped/transposed weights, ReLU mask, += fan-in.  cout = 16 is the
    side_prep data gradient: on the tensor-core path one K = 16 MMA per tap (conv_tc.cu MODE_K16, SWIZZLE_32B operands)."""
    g = _gen(5)
    n, h, w_, cin, cout = shape
    x = torch.randn(n, cin, h, w_, generator=g).clamp_min(0)             # post-ReLU activation (has zeros)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.05
    dz = torch.randn(n, cout, h, w_, generator=g)
    prev = torch.randn(n, cin, h, w_, generator=g)
    r = (lambda t: t) if dt == torch.float32 else _bf16r
    ref = F.conv_transpose2d(r(dz).double(), r(wt).double(), padding=1).float() * (r(x) > 0) + r(prev)
    wp = ops.pack_weight(wt.to(DEV), L.W_TC_DGRAD if impl == "tc" else L.W_SIMT_DGRAD, dt)
    out = _nhwc(prev, dt)
    ops.conv3x3(_nhwc(dz, dt), wp, None, cin, L.CONV_MASK | L.CONV_ACCUMULATE, mask=_nhwc(x, dt), out=out, impl=impl)
    got = _nchw(out, cin)
    tol = dict(rtol=1e-4, atol=1e-4) if dt == torch.float32 else dict(rtol=2 ** -6, atol=5e-2)
    assert torch.allclose(got, ref, **tol), float((got - ref).abs().max())


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", [(1, 9, 11, 3, 64), (2, 12, 20, 64, 16), (1, 7, 5, 40, 72)])
def test_conv3x3_wgrad(dt, case):
    n, h, w_, cin, cout = case
    g = _gen(7)
    x = torch.randn(n, cin, h, w_, generator=g)
    dz = torch.randn(n, cout, h, w_, generator=g)
    r = (lambda t: t) if dt == torch.float32 else _bf16r
    wref = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(r(x).double(), wref, padding=1)
    (gw,) = torch.autograd.grad(y, wref, r(dz).double())
    gb = r(dz).double().sum(dim=(0, 2, 3))
    dw = torch.full((cout, cin, 3, 3), 0.5, device=DEV)       # accumulates on top of existing .grad
    db = torch.full((cout,), -1.0, device=DEV)
    ops.conv3x3_wgrad(_nhwc(x, dt), _nhwc(dz, dt), dw, db)
    assert torch.allclose(dw.cpu() - 0.5, gw.float(), rtol=1e-4, atol=2e-3), float((dw.cpu() - 0.5 - gw.float()).abs().max())
    assert torch.allclose(db.cpu() + 1.0, gb.float(), rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize("case", [(1, 16, 24, 64, 64), (1, 17, 23, 64, 128), (2, 9, 30, 128, 16), (1, 33, 40, 256, 256),
                                  (1, 30, 54, 512, 512), (1, 12, 20, 40, 72), (1, 8, 8, 128, 128), (2, 17, 23, 128, 256),
                                  (3, 20, 27, 256, 512), (2, 21, 37, 64, 64), (2, 13, 50, 64, 128), (1, 40, 40, 64, 256),
                                  (1, 20, 21, 256, 24), (1, 12, 13, 64, 16), (2, 30, 54, 512, 16)])
def test_conv3x3_wgrad_tc(case):
    """tcgen05 weight gradient (MN-major operands, halo-box tap reuse, split-K reductions).  Cout % 256 == 0 with
    Cin % 128 == 0 runs the CTA-pair kernel (cta_group::2, M = 256 over the two SMs of a TPC); Cin = 64 runs the row-stacked
    N = 192 MMAs (Cout = 64: all nine taps from one visit of each patch)."""
    n, h, w_, cin, cout = case
    g = _gen(17)
    x = _bf16r(torch.randn(n, cin, h, w_, generator=g))
    dz = _bf16r(torch.randn(n, cout, h, w_, generator=g))
    wref = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    (gw,) = torch.autograd.grad(F.conv2d(x.double(), wref, padding=1), wref, dz.double())
    gb = dz.double().sum(dim=(0, 2, 3))
    dw = torch.full((cout, cin, 3, 3), 0.25, device=DEV)
    db = torch.full((cout,), -2.0, device=DEV)
    ops.conv3x3_wgrad(_nhwc(x, torch.bfloat16), _nhwc(dz, torch.bfloat16), dw, db, impl="tc")
    err = float((dw.cpu() - 0.25 - gw.float()).abs().max())
    assert torch.allclose(dw.cpu() - 0.25, gw.float(), rtol=1e-4, atol=5e-3), err
    assert torch.allclose(db.cpu() + 2.0, gb.float(), rtol=1e-4, atol=5e-3)


@pytest.mark.parametrize("case", [(1, 16, 24, 8, 64), (1, 19, 40, 64, 64), (2, 9, 30, 128, 16), (1, 30, 54, 512, 512), (1, 12, 20, 40, 72)])
@pytest.mark.parametrize("splits", [None, 1, 5])
def test_conv3x3_wgrad_tc_accumulate_finish(case, splits, monkeypatch):
    """Two micro-iterations accumulate in the live [tap][M][N] workspace (bias gradient summed by the idle
    epilogue warps of the same kernel), one `finish` folds them into the OIHW .grad and clears the workspace."""
    n, h, w_, cin, cout = case
    if splits is not None:
        monkeypatch.setenv("FOSVOS_WG_SPLITS", str(splits))
    g = _gen(23)
    cinp, coutp = ops.pad8(cin), ops.pad8(cout)
    ws = ops.wgrad_workspace(cinp, coutp, DEV)
    dw = torch.full((cout, cin, 3, 3), 0.25, device=DEV)
    db = torch.full((cout,), -2.0, device=DEV)
    gw_ref = torch.zeros(cout, cin, 3, 3, dtype=torch.float64)
    gb_ref = torch.zeros(cout, dtype=torch.float64)
    for it in range(2):
        x = _bf16r(torch.randn(n, cin, h, w_, generator=g))
        dz = _bf16r(torch.randn(n, cout, h, w_, generator=g))
        wref = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
        gw_ref += torch.autograd.grad(F.conv2d(x.double(), wref, padding=1), wref, dz.double())[0]
        gb_ref += dz.double().sum(dim=(0, 2, 3))
        ops.conv3x3_wgrad_accumulate(_nhwc(x, torch.bfloat16), _nhwc(dz, torch.bfloat16), ws, db, cout)
    assert torch.equal(dw.cpu(), torch.full((cout, cin, 3, 3), 0.25))         # untouched until the fold
    ops.conv3x3_wgrad_finish(ws, dw, cinp, coutp, zero_workspace=True)
    assert torch.allclose(dw.cpu() - 0.25, gw_ref.float(), rtol=1e-4, atol=5e-3), float((dw.cpu() - 0.25 - gw_ref.float()).abs().max())
    assert torch.allclose(db.cpu() + 2.0, gb_ref.float(), rtol=1e-4, atol=5e-3), float((db.cpu() + 2.0 - gb_ref.float()).abs().max())
    assert float(ws.abs().max()) == 0.0


def test_fused_conv_step_equals_fold_sgd_repack():
    """`conv_step_all` (fold + SGD with momentum + repack of the 3x3 conv weights in one launch, step.cu) against the three
    separate launches it replaces, two steps (the second with a momentum history), several shapes incl. ragged tiles."""
    g = _gen(33)
    shapes = [(64, 3), (16, 128), (72, 40), (128, 64), (64, 64)]
    mu = 0.9
    st = []
    for k, (cout, cin) in enumerate(shapes):
        cinp, coutp = ops.pad8(cin), ops.pad8(cout)
        w = torch.randn(cout, cin, 3, 3, generator=g).to(DEV)
        b = torch.randn(cout, generator=g).to(DEV)
        st.append(dict(cout=cout, cin=cin, cinp=cinp, coutp=coutp, lr=1e-3 * (k + 1), wd=2e-4 if k % 2 == 0 else 0.0,
                       w_a=w.clone(), w_b=w.clone(), buf_a=torch.zeros_like(w), buf_b=torch.zeros_like(w), bias=b,
                       ws_a=ops.wgrad_workspace(cinp, coutp, DEV), ws_b=ops.wgrad_workspace(cinp, coutp, DEV),
                       dw=torch.zeros_like(w), dw_b=torch.zeros_like(w),
                       fwd_a=ops.pack_weight(w, L.W_TC_FWD, torch.bfloat16), dgr_a=ops.pack_weight(w, L.W_TC_DGRAD, torch.bfloat16),
                       fwd_b=ops.pack_weight(w, L.W_TC_FWD, torch.bfloat16), dgr_b=ops.pack_weight(w, L.W_TC_DGRAD, torch.bfloat16),
                       bo_a=ops.pad_bias(None, cout, DEV), bo_b=ops.pad_bias(None, cout, DEV)))
    fold_t = ops.fold_table([(e["ws_a"], e["dw"]) for e in st], DEV)
    sgd_t = ops.sgd_table([(e["w_a"], e["dw"], e["buf_a"], e["lr"], e["wd"]) for e in st], DEV)
    rep_t = ops.repack_table([(e["w_a"], e["bias"], e["fwd_a"], e["dgr_a"], e["bo_a"]) for e in st], DEV)
    # the fused step also takes a pending .grad (cleared like the accumulator): entries 1 and 3 carry one
    conv_t = ops.convstep_table([(e["ws_b"], e["dw_b"] if k in (1, 3) else None, e["w_b"], e["buf_b"], e["bias"], e["fwd_b"], e["dgr_b"], e["bo_b"],
                                  e["lr"], e["wd"]) for k, e in enumerate(st)], DEV)
    for step in range(2):
        for k, e in enumerate(st):
            if k in (1, 3):
                e["dw"].copy_(torch.randn(e["dw"].shape, generator=g) * 0.01)
                e["dw_b"].copy_(e["dw"])
        for e in st:
            x = _nhwc(_bf16r(torch.randn(1, e["cin"], 10, 16, generator=g)), torch.bfloat16)
            dz = _nhwc(_bf16r(torch.randn(1, e["cout"], 10, 16, generator=g)), torch.bfloat16)
            ops.conv3x3_wgrad_accumulate(x, dz, e["ws_a"], None, e["cout"])
            e["ws_b"].copy_(e["ws_a"])
        ops.wgrad_fold_all(fold_t)
        ops.sgd_step(sgd_t[0], sgd_t[1], len(st), sgd_t[2], mu, True)
        ops.repack_all(rep_t)
        ops.conv_step_all(conv_t, mu)
        for e in st:
            assert float(e["ws_b"].abs().max()) == 0.0 and float(e["dw_b"].abs().max()) == 0.0
            assert torch.allclose(e["w_b"], e["w_a"], rtol=1e-6, atol=1e-7), (step, float((e["w_b"] - e["w_a"]).abs().max()))
            assert torch.allclose(e["buf_b"], e["buf_a"], rtol=1e-6, atol=1e-7)
            assert torch.equal(e["bo_b"], e["bo_a"])
            # the packed copies are bf16 roundings of (nearly) identical fp32 weights
            assert float((e["fwd_b"].float() - e["fwd_a"].float()).abs().max()) <= 2 ** -7 * float(e["w_a"].abs().max())
            assert float((e["dgr_b"].float() - e["dgr_a"].float()).abs().max()) <= 2 ** -7 * float(e["w_a"].abs().max())
    # in-place rewrite of the table (new learning rates) keeps its device buffer
    conv_t2 = ops.convstep_table([(e["ws_b"], e["dw_b"] if k in (1, 3) else None, e["w_b"], e["buf_b"], e["bias"], e["fwd_b"], e["dgr_b"], e["bo_b"],
                                   2 * e["lr"], e["wd"]) for k, e in enumerate(st)], DEV, out=conv_t)
    assert conv_t2[0].data_ptr() == conv_t[0].data_ptr()


def test_multi_tensor_fold_and_repack():
    """One-launch `wgrad_fold_all` / `repack_all` over several convs == the per-layer finish / pack calls."""
    g = _gen(31)
    shapes = [(64, 3), (16, 128), (72, 40), (128, 64), (20, 12)]        # (cout, cin): both orientations, ragged tiles
    fold_ent, ref = [], []
    for cout, cin in shapes:
        cinp, coutp = ops.pad8(cin), ops.pad8(cout)
        x = _bf16r(torch.randn(1, cin, 10, 16, generator=g))
        dz = _bf16r(torch.randn(1, cout, 10, 16, generator=g))
        ws = ops.wgrad_workspace(cinp, coutp, DEV)
        ops.conv3x3_wgrad_accumulate(_nhwc(x, torch.bfloat16), _nhwc(dz, torch.bfloat16), ws, None, cout)
        dw = torch.full((cout, cin, 3, 3), 0.5, device=DEV)
        dw_ref = dw.clone()
        ops.conv3x3_wgrad_finish(ws.clone(), dw_ref, cinp, coutp, zero_workspace=False)
        fold_ent.append((ws, dw))
        ref.append(dw_ref)
    ops.wgrad_fold_all(ops.fold_table(fold_ent, DEV))
    for (ws, dw), r in zip(fold_ent, ref):
        assert torch.equal(dw, r)
        assert float(ws.abs().max()) == 0.0
    pack_ent, refs = [], []
    for cout, cin in shapes:
        w = torch.randn(cout, cin, 3, 3, generator=g).to(DEV)
        b = torch.randn(cout, generator=g).to(DEV)
        f_ref = ops.pack_weight(w, L.W_TC_FWD, torch.bfloat16)
        d_ref = ops.pack_weight(w, L.W_TC_DGRAD, torch.bfloat16)
        b_ref = ops.pad_bias(b, cout, DEV)
        # buffers first filled from OTHER values by the full pack (padding zero), then refreshed in place
        w0 = torch.randn(cout, cin, 3, 3, generator=g).to(DEV)
        f = ops.pack_weight(w0, L.W_TC_FWD, torch.bfloat16)
        d = ops.pack_weight(w0, L.W_TC_DGRAD, torch.bfloat16)
        bo = ops.pad_bias(None, cout, DEV)
        pack_ent.append((w, b, f, d, bo))
        refs.append((f_ref, d_ref, b_ref))
    ops.repack_all(ops.repack_table(pack_ent, DEV))
    for (w, b, f, d, bo), (f_ref, d_ref, b_ref) in zip(pack_ent, refs):
        assert torch.equal(f, f_ref) and torch.equal(d, d_ref) and torch.equal(bo, b_ref)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 64, 480 // 8, 854 // 7), (2, 16, 7, 9), (1, 8, 1, 1), (1, 24, 30, 107)])
def test_maxpool_fwd_bwd(dt, shape):
    g = _gen(9)
    x = torch.randn(shape, generator=g).clamp_min(0)          # ties at 0 like post-ReLU maps
    r = (lambda t: t) if dt == torch.float32 else _bf16r
    xr = r(x).requires_grad_(True)
    ref = F.max_pool2d(xr, 2, 2, ceil_mode=True)
    xd = _nhwc(x, dt)
    y = ops.maxpool2x2(xd)
    assert torch.equal(_nchw(y, shape[1]), ref.detach())
    dy = r(torch.randn(ref.shape, generator=g))
    (gref,) = torch.autograd.grad(ref, xr, dy)
    dx = ops.maxpool2x2_bwd(xd, _nhwc(dy, dt))
    assert torch.equal(_nchw(dx, shape[1]), gref)


def _side_inputs(n, H, W, seed, dt):
    g = _gen(seed)
    hs, ws, h, w_ = [], [], H, W
    sp_nchw = []
    for i in range(4):
        h, w_ = (h + 1) // 2, (w_ + 1) // 2
        hs.append(h)
        ws.append(w_)
        sp_nchw.append(torch.randn(n, 16, h, w_, generator=g))
    sd = {}
    for i in range(4):
        k = 4 << i
        sd[f"upscale.{i}.weight"] = O.interp_surgery_weight(16, k)
        sd[f"upscale_.{i}.weight"] = O.interp_surgery_weight(1, k)
        sd[f"score_dsn.{i}.weight"] = torch.randn(1, 16, 1, 1, generator=g) * 0.3
        sd[f"score_dsn.{i}.bias"] = torch.randn(1, generator=g)
    sd["fuse.weight"] = torch.randn(1, 64, 1, 1, generator=g) * 0.2
    sd["fuse.bias"] = torch.randn(1, generator=g)
    return sp_nchw, sd


def _side_ref(sp_nchw, sd, H, W):
    side, side_out = [], []
    for i in range(4):
        s = 2 << i
        side.append(O.center_crop(F.conv_transpose2d(sp_nchw[i], sd[f"upscale.{i}.weight"], stride=s), H, W))
        sc = F.conv2d(sp_nchw[i], sd[f"score_dsn.{i}.weight"], sd[f"score_dsn.{i}.bias"])
        side_out.append(O.center_crop(F.conv_transpose2d(sc, sd[f"upscale_.{i}.weight"], stride=s), H, W))
    fused = F.conv2d(torch.cat(side, 1), sd["fuse.weight"], sd["fuse.bias"])
    return side_out + [fused]


def _side_params(sd):
    d = {k: v.to(DEV) for k, v in sd.items()}
    return ops.side_params_prepare([d[f"upscale.{i}.weight"] for i in range(4)], [d[f"upscale_.{i}.weight"] for i in range(4)],
                                   [d[f"score_dsn.{i}.weight"] for i in range(4)], [d[f"score_dsn.{i}.bias"] for i in range(4)],
                                   d["fuse.weight"], d["fuse.bias"]), d


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("general", [False, True, 2])
@pytest.mark.parametrize("HW", [(48, 72), (45, 70), (33, 17), (70, 1100)])
def test_side_chain_forward(dt, general, HW):
    H, W = HW
    sp, sd = _side_inputs(2, H, W, 11, dt)
    r = (lambda t: t) if dt == torch.float32 else _bf16r
    ref = _side_ref([r(t) for t in sp], sd, H, W)
    params, dsd = _side_params(sd)
    assert int(ops.side_check_diagonal([dsd[f"upscale.{i}.weight"] for i in range(4)]).item()) == 0
    assert ops.side_separable(params)              # the bilinear kernels of interp_surgery factor exactly in fp32
    outs, prob, mask = ops.side_fwd([_nhwc(t, dt) for t in sp], params, H, W, general=general, want_prob=True, want_mask=True)
    for a, b in zip(outs, ref):
        assert torch.allclose(a.cpu(), b, rtol=1e-5, atol=2e-5), float((a.cpu() - b).abs().max())
    p = O.probabilities(outs[4].cpu())
    assert torch.allclose(prob.cpu(), p, atol=1e-6), (float((prob.cpu() - p).abs().max()), int(((prob.cpu() - p).abs() > 1e-6).sum()))
    assert torch.equal(mask.cpu(), O.binarise(prob.cpu()))


@pytest.mark.parametrize("HW", [(70, 300), (37, 259), (130, 31)])
def test_side_chain_fast_path_arbitrary_shared_kernel(HW):
    """The fast path only needs `upscale` to be diagonal with ONE shared k x k kernel; that kernel (and the
    `upscale_` one) may be anything, not just the bilinear one.  Sizes cross the 256-column / 32-row strips."""
    H, W = HW
    sp, sd = _side_inputs(2, H, W, 21, torch.float32)
    g = _gen(22)
    for i in range(4):
        k = 4 << i
        shared = torch.randn(k, k, generator=g) * 0.2
        w = torch.zeros(16, 16, k, k)
        for c in range(16):
            w[c, c] = shared
        sd[f"upscale.{i}.weight"] = w
        sd[f"upscale_.{i}.weight"] = torch.randn(1, 1, k, k, generator=g) * 0.2
    ref = _side_ref(sp, sd, H, W)
    params, dsd = _side_params(sd)
    assert int(ops.side_check_diagonal([dsd[f"upscale.{i}.weight"] for i in range(4)]).item()) == 0
    assert not ops.side_separable(params)          # a random k x k kernel has full rank: stays on the phase-table path
    outs, prob, mask = ops.side_fwd([_nhwc(t, torch.float32) for t in sp], params, H, W, general=False, want_prob=True, want_mask=True)
    for a, b in zip(outs, ref):
        assert torch.allclose(a.cpu(), b, rtol=1e-5, atol=5e-5), float((a.cpu() - b).abs().max())
    assert torch.allclose(prob.cpu(), O.probabilities(outs[4].cpu()), atol=1e-6)
    assert torch.equal(mask.cpu(), O.binarise(prob.cpu()))


def test_side_chain_general_dense_upscale():
    """Non-diagonal `upscale` weights (e.g. after an Adam step, prune.py:256): exact general path."""
    H, W = 40, 56
    sp, sd = _side_inputs(1, H, W, 13, torch.float32)
    g = _gen(14)
    for i in range(4):
        sd[f"upscale.{i}.weight"] = sd[f"upscale.{i}.weight"] + 0.01 * torch.randn(sd[f"upscale.{i}.weight"].shape, generator=g)
        sd[f"upscale_.{i}.weight"] = sd[f"upscale_.{i}.weight"] + 0.01 * torch.randn(sd[f"upscale_.{i}.weight"].shape, generator=g)
    ref = _side_ref(sp, sd, H, W)
    params, dsd = _side_params(sd)
    assert int(ops.side_check_diagonal([dsd[f"upscale.{i}.weight"] for i in range(4)]).item()) > 0
    outs, _, _ = ops.side_fwd([_nhwc(t, torch.float32) for t in sp], params, H, W, general=True)
    for a, b in zip(outs, ref):
        assert torch.allclose(a.cpu(), b, rtol=1e-5, atol=5e-5), float((a.cpu() - b).abs().max())


@pytest.mark.parametrize("deep", [False, True])
def test_side_chain_backward(deep):
    H, W = 45, 70
    n = 2
    sp, sd = _side_inputs(n, H, W, 15, torch.float32)
    g = _gen(16)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items() if not k.startswith("upscale")}
    spl = [t.clone().requires_grad_(True) for t in sp]
    full = dict(sd)
    full.update(leaf)
    outs = _side_ref(spl, full, H, W)
    douts = [torch.randn(n, 1, H, W, generator=g) if (deep or i == 4) else None for i in range(5)]
    loss = sum((o * d).sum() for o, d in zip(outs, douts) if d is not None)
    loss.backward()
    params, _ = _side_params(sd)
    dfw = torch.zeros(1, 64, 1, 1, device=DEV)
    dfb = torch.zeros(1, device=DEV)
    dsw = [torch.zeros(1, 16, 1, 1, device=DEV) for _ in range(4)]
    dsb = [torch.zeros(1, device=DEV) for _ in range(4)]
    dsp = ops.side_bwd([_nhwc(t, torch.float32) for t in sp], params, [None if d is None else d.to(DEV) for d in douts],
                       H, W, dfw, dfb, dsw, dsb)
    for i in range(4):
        assert torch.allclose(_nchw(dsp[i], 16), spl[i].grad, rtol=1e-4, atol=1e-4), i
        if deep:
            assert torch.allclose(dsw[i].cpu(), leaf[f"score_dsn.{i}.weight"].grad, rtol=1e-4, atol=1e-3)
            assert torch.allclose(dsb[i].cpu(), leaf[f"score_dsn.{i}.bias"].grad, rtol=1e-4, atol=1e-3)
        else:
            assert float(dsw[i].abs().max()) == 0.0
    assert torch.allclose(dfw.cpu(), leaf["fuse.weight"].grad, rtol=1e-4, atol=1e-3)
    assert torch.allclose(dfb.cpu(), leaf["fuse.bias"].grad, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("shape", [(1, 1, 4, 4), (2, 1, 37, 53), (1, 1, 480, 854)])
@pytest.mark.parametrize("size_average", [True, False])
def test_balanced_loss(shape, size_average):
    g = _gen(21)
    out = torch.randn(shape, generator=g) * 4
    lab = (torch.rand(shape, generator=g) > 0.7).float()
    o = out.clone().requires_grad_(True)
    ref = O.class_balanced_cross_entropy_loss(o, lab, size_average)
    (gref,) = torch.autograd.grad(ref, o, torch.tensor(0.2))
    loss, stats = ops.bal_loss_fwd(out.to(DEV), lab.to(DEV), size_average)
    assert torch.allclose(loss.cpu(), ref.detach(), rtol=2e-5), (float(loss), float(ref))
    assert int(stats[0].item()) == int(lab.sum().item()) and int(stats[1].item()) == int((1 - lab).sum().item())
    dx = ops.bal_loss_bwd(out.to(DEV), lab.to(DEV), size_average, stats, torch.tensor(0.2, device=DEV))
    assert torch.allclose(dx.cpu(), gref, rtol=1e-4, atol=1e-7 * float(gref.abs().max()) + 1e-12)


@pytest.mark.parametrize("shape", [(1, 1, 4, 4), (2, 1, 37, 53), (1, 1, 480, 854)])
@pytest.mark.parametrize("size_average", [True, False])
def test_balanced_loss_fused_forward_backward(shape, size_average):
    """One-pass loss + gradient with the label counts of an earlier forward == the two-kernel path."""
    g = _gen(61)
    x = torch.randn(shape, generator=g) * 3
    lab = (torch.rand(shape, generator=g) > 0.7).float()
    xr = x.clone().requires_grad_(True)
    ref = O.class_balanced_cross_entropy_loss(xr, lab, size_average=size_average)
    (0.2 * ref).backward()
    xd, ld = x.to(DEV), lab.to(DEV)
    _, stats = ops.bal_loss_fwd(torch.zeros_like(xd), ld, size_average)      # counts only depend on the label
    for _ in range(2):                                                       # the stats buffer is reusable
        loss, dx = ops.bal_loss_fwd_bwd(xd, ld, size_average, stats, None, 0.2)
        assert abs(float(loss) - float(ref)) <= 2e-5 * abs(float(ref)) + 1e-7
        assert torch.allclose(dx.cpu(), xr.grad, rtol=1e-4, atol=1e-7 * float(xr.grad.abs().max()) + 1e-12)


def test_balanced_loss_golden_kat():
    from fosvos_b200 import class_balanced_cross_entropy_loss
    from conftest import GOLDEN
    kat = torch.load(os.path.join(GOLDEN, "kat.pt"), weights_only=False)
    for name in ["loss_4x4_sa1", "loss_4x4_sa0", "loss_2x1x37x53_sa1", "loss_2x1x37x53_sa0", "loss_teacher_logits"]:
        c = kat[name]
        out = c["output"].to(DEV).requires_grad_(True)
        loss = class_balanced_cross_entropy_loss(out, c["label"].to(DEV), size_average=not name.endswith("sa0"))
        assert loss.dim() == 0
        assert torch.allclose(loss.detach().cpu(), c["loss"], rtol=2e-5), name
        loss.backward()
        assert torch.allclose(out.grad.cpu(), c["grad"], rtol=1e-4, atol=1e-7), name


def test_fused_sgd_matches_reference_sgd():
    from fosvos_b200 import FusedSGD
    g = _gen(31)
    shapes = [(64, 3, 3, 3), (64,), (5,), (16, 512, 3, 3), (1, 64, 1, 1), (1,), (4099,)]
    ps = [torch.randn(s, generator=g) for s in shapes]
    cfg = [(1e-2, 2e-4), (2e-2, 0.0), (0.0, 0.0), (1e-2, 2e-4), (1e-4, 2e-4), (2e-4, 0.0), (0.5, 0.1)]
    mine = [torch.nn.Parameter(p.clone().to(DEV)) for p in ps]
    opt = FusedSGD([dict(params=[m], lr=lr, weight_decay=wd) for m, (lr, wd) in zip(mine, cfg)], lr=1e-2, momentum=0.9)
    ref = [p.clone() for p in ps]
    bufs = [None] * len(ps)
    for step in range(3):
        grads = [torch.randn(s, generator=g) for s in shapes]
        for m, gr in zip(mine, grads):
            m.grad = gr.clone().to(DEV) if m.grad is None else m.grad.copy_(gr.to(DEV))
        opt.step_and_zero()
        for i, (lr, wd) in enumerate(cfg):
            ref[i], bufs[i] = O.sgd_step(ref[i], grads[i], bufs[i], lr, wd, 0.9)
        for m, r_ in zip(mine, ref):
            assert torch.allclose(m.detach().cpu(), r_, rtol=1e-6, atol=2e-6)
            assert float(m.grad.abs().max()) == 0.0
    # and against torch.optim.SGD itself
    tp = [torch.nn.Parameter(p.clone()) for p in ps]
    topt = torch.optim.SGD([dict(params=[t], lr=lr, weight_decay=wd) for t, (lr, wd) in zip(tp, cfg)], lr=1e-2, momentum=0.9)
    g2 = _gen(31)
    _ = [torch.randn(s, generator=g2) for s in shapes]
    for step in range(3):
        for t, s in zip(tp, shapes):
            t.grad = torch.randn(s, generator=g2)
        topt.step()
    for m, t in zip(mine, tp):
        assert torch.allclose(m.detach().cpu(), t.detach(), rtol=1e-6, atol=2e-6)


def test_mask_iou_bit_exact():
    g = _gen(41)
    a = (torch.rand(5, 1, 480, 854, generator=g) > 0.6).to(torch.uint8)
    b = (torch.rand(5, 1, 480, 854, generator=g) > 0.5).to(torch.uint8)
    a[3] = 0
    b[3] = 0                                                   # empty masks: 0/0
    counts = ops.mask_iou(a.to(DEV).view(5, -1), b.to(DEV).view(5, -1)).cpu()
    for f in range(5):
        assert tuple(counts[f].tolist()) == O.mask_iou_counts(a[f], b[f])
    # odd length / unaligned tail
    a1, b1 = a.view(5, -1)[:, :1001].contiguous(), b.view(5, -1)[:, :1001].contiguous()
    c1 = ops.mask_iou(a1.to(DEV), b1.to(DEV)).cpu()
    for f in range(5):
        assert tuple(c1[f].tolist()) == O.mask_iou_counts(a1[f], b1[f])


def test_sigmoid_threshold():
    x = torch.randn(3, 1, 17, 19, generator=_gen(43)) * 5
    x[0, 0, 0, 0] = 0.0
    prob, mask = ops.sigmoid_threshold(x.to(DEV))
    assert torch.allclose(prob.cpu(), O.probabilities(x), atol=1e-6)
    assert torch.equal(mask.cpu(), O.binarise(prob.cpu()))
    assert int(mask[0, 0, 0, 0]) == 1
