"""Diagnostic sweep of the tcgen05 conv kernel: per layer shape, time the full kernel and variants with one
pipeline component dropped (A loads / B loads / epilogue stores / MMAs) and with forced BN.
   python tools/conv_probe.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fosvos_b200 import _lib as L
from fosvos_b200 import ops

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
LAYERS = [(480, 854, 8, 64), (480, 854, 64, 64), (240, 427, 64, 128), (240, 427, 128, 128), (240, 427, 128, 16),
          (120, 214, 128, 256), (120, 214, 256, 256), (120, 214, 256, 16), (60, 107, 256, 512), (60, 107, 512, 512),
          (60, 107, 512, 16), (30, 54, 512, 512), (30, 54, 512, 16),
          # data-gradient shapes not already covered (cin/cout swapped)
          (240, 427, 128, 64), (120, 214, 256, 128), (60, 107, 512, 256), (240, 427, 16, 128), (120, 214, 16, 256),
          (60, 107, 16, 512), (30, 54, 16, 512)]
DBG = {"full": 0}


def timeit(fn, reps=20):
    """GPU time per call: `reps` calls captured in one CUDA graph (no host launch gaps), replayed 3x."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            g.replay()
            e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


print(f"batch {batch}; times in us; TF = algorithmic TFLOP/s of the full kernel")
for (h, w, cin, cout) in LAYERS:
    x = torch.randn((batch, h, w, cin), device=dev).to(torch.bfloat16)
    kpad = (cin + 63) // 64 * 64
    wp = (torch.randn((cout * 9 * kpad,), device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.zeros(cout, device=dev)
    y = torch.empty((batch, h, w, cout), device=dev, dtype=torch.bfloat16)
    flops = 2.0 * batch * h * w * cin * cout * 9
    row = []
    for bn in ([None] if (cout <= 16 or os.environ.get("PROBE_BN", "1") != "1") else [None, 64, 128, 256]):
        if bn is not None and bn > max(cout, 64):
            continue
        if bn is None:
            os.environ.pop("FOSVOS_TC_BN", None)
        else:
            os.environ["FOSVOS_TC_BN"] = str(bn)
        for name, d in DBG.items():
            if bn is not None and name not in ("full", "onlyMMA"):
                continue
            t = timeit(lambda: ops.conv3x3(x, wp, bias, cout, L.CONV_BIAS | L.CONV_RELU | d, out=y))
            row.append(f"{'bn' + str(bn) + ':' if bn else ''}{name}={t:.1f}" + (f"({flops / t / 1e6:.0f}TF)" if name == "full" else ""))
    os.environ.pop("FOSVOS_TC_BN", None)
    if cout >= 64 and os.environ.get("PROBE_MASK", "1") == "1":
        # data-gradient epilogue: ReLU mask of the layer input, staged-slab path vs register path
        mk = (torch.rand((batch, h, w, cout), device=dev) > 0.5).to(torch.bfloat16)
        for name, env in (("mask_slab", None), ("mask_regs", "1")):
            if env is None:
                os.environ.pop("FOSVOS_TC_MASK_REGS", None)
            else:
                os.environ["FOSVOS_TC_MASK_REGS"] = env
            t = timeit(lambda: ops.conv3x3(x, wp, None, cout, L.CONV_MASK, mask=mk, out=y))
            row.append(f"{name}={t:.1f}")
        os.environ.pop("FOSVOS_TC_MASK_REGS", None)
    print(f"{h}x{w} {cin}->{cout}: " + " ".join(row), flush=True)
