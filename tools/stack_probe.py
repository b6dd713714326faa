"""Timing of the Cout = 64 layers (row-stacked kernel vs the generic one: FOSVOS_TC_NO_STACK=1) at the shapes of the job."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fosvos_b200 import _lib as L
from fosvos_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def run(tag):
    g = torch.Generator().manual_seed(0)
    w = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).to(dev)
    wf = ops.pack_weight(w, L.W_TC_FWD, torch.bfloat16)
    wd = ops.pack_weight(w, L.W_TC_DGRAD, torch.bfloat16)
    w2 = (torch.randn(128, 64, 3, 3, generator=g) * 0.05).to(dev)
    w2d = ops.pack_weight(w2, L.W_TC_DGRAD, torch.bfloat16)
    bp = ops.pad_bias(torch.randn(64, generator=g).to(dev), 64, dev)
    fl = L.CONV_BIAS | L.CONV_RELU
    for n in (16, 5):
        x = torch.randn(n, 480, 854, 64, device=dev).to(torch.bfloat16)
        gf = 2 * 9 * 64 * 64 * n * 480 * 854 / 1e9
        t = timeit(lambda: ops.conv3x3_pool_only(x, wf, bp, 64, fl))
        print(f"{tag} conv1_2 fwd pool-only  batch {n}: {t:7.1f} us  {gf / t * 1e3:7.1f} TFLOP/s", flush=True)
        t = timeit(lambda: ops.conv3x3_pool(x, wf, bp, 64, fl))
        print(f"{tag} conv1_2 fwd y + pool   batch {n}: {t:7.1f} us  {gf / t * 1e3:7.1f} TFLOP/s", flush=True)
        del x
    n = 5
    dz = torch.randn(n, 480, 854, 64, device=dev).to(torch.bfloat16)
    m = torch.randn(n, 480, 854, 64, device=dev).clamp_min(0).to(torch.bfloat16)
    out = torch.empty_like(m)
    gf = 2 * 9 * 64 * 64 * n * 480 * 854 / 1e9
    t = timeit(lambda: ops.conv3x3(dz, wd, None, 64, L.CONV_MASK, mask=m, out=out))
    print(f"{tag} conv1_2 dgrad (mask)   batch {n}: {t:7.1f} us  {gf / t * 1e3:7.1f} TFLOP/s", flush=True)
    del dz, m, out
    dz = torch.randn(n, 240, 427, 128, device=dev).to(torch.bfloat16)
    m = torch.randn(n, 240, 427, 64, device=dev).clamp_min(0).to(torch.bfloat16)
    out = torch.empty_like(m)
    gf = 2 * 9 * 128 * 64 * n * 240 * 427 / 1e9
    t = timeit(lambda: ops.conv3x3(dz, w2d, None, 64, L.CONV_MASK, mask=m, out=out))
    print(f"{tag} conv2_1 dgrad (mask)   batch {n}: {t:7.1f} us  {gf / t * 1e3:7.1f} TFLOP/s", flush=True)


run("generic " if os.environ.get("FOSVOS_TC_NO_STACK") else "rowstack")
