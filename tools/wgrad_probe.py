"""Time the tcgen05 weight-gradient kernel per layer shape; check the Cin=3 (padded 8) case against the direct kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fosvos_b200 import ops

dev = torch.device("cuda:0")
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
LAYERS = [(480, 854, 3, 64), (480, 854, 64, 64), (240, 427, 64, 128), (240, 427, 128, 128), (240, 427, 128, 16),
          (120, 214, 128, 256), (120, 214, 256, 256), (120, 214, 256, 16), (60, 107, 256, 512), (60, 107, 512, 512),
          (60, 107, 512, 16), (30, 54, 512, 512), (30, 54, 512, 16)]


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


SPLITS = [None] + [int(v) for v in os.environ.get("PROBE_SPLITS", "1,2,3,4,6,8,12").split(",")]
for (h, w, cin, cout) in LAYERS:
    cinp = ops.pad8(cin)
    x = torch.randn((batch, h, w, cinp), device=dev).to(torch.bfloat16)
    if cinp != cin:
        x[..., cin:] = 0
    dz = (torch.randn((batch, h, w, cout), device=dev) * 0.1).to(torch.bfloat16)
    flops = 2.0 * batch * h * w * cin * cout * 9
    msg = f"{h}x{w} {cin}->{cout}:"
    ws = ops.wgrad_workspace(cinp, ops.pad8(cout), dev)
    db = torch.zeros(cout, device=dev)
    for sp in SPLITS:
        if sp is None:
            os.environ.pop("FOSVOS_WG_SPLITS", None)
        else:
            os.environ["FOSVOS_WG_SPLITS"] = str(sp)
        t = timeit(lambda: ops.conv3x3_wgrad_accumulate(x, dz, ws, db, cout))
        msg += f" {'auto' if sp is None else 's' + str(sp)}={t:.1f}" + (f"us({flops / t / 1e6:.0f}TF)" if sp is None else "")
    os.environ.pop("FOSVOS_WG_SPLITS", None)
    t2 = timeit(lambda: ops.conv3x3_wgrad_accumulate(x, dz, ws, None, cout))
    msg += f" | auto nobias={t2:.1f}"
    print(msg, flush=True)
