"""The side chain on the path the network takes (heads + separable upsample), bf16 activations: GPU time per call and
GB/s of the algorithmic bytes.  python tools/side_sep_probe.py [batch ...]   (A/B switches: FOSVOS_SIDE_HEADS=0|1,
FOSVOS_SIDE_SEP2_VARIANT=0..3, FOSVOS_SIDE_NO_SEP2=1)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fosvos_b200 as FB
from fosvos_b200 import ops


def timeit(fn, reps=20):
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


H, W = 480, 854
dev = torch.device("cuda:0")
net = FB.OSVOS_VGG(pretrained=0).to(dev)
params = net._side()
assert ops.side_separable(params)
out = []
for batch in [int(a) for a in sys.argv[1:]] or [1, 5, 16]:
    sps = []
    h, w = H, W
    for i in range(4):
        h, w = (h + 1) // 2, (w + 1) // 2
        sps.append(torch.randn((batch, h, w, 16), device=dev).to(torch.bfloat16))
    low = sum(t.shape[1] * t.shape[2] for t in sps)
    if os.environ.get("FOSVOS_SIDE_PROBE_HEADS", "done") == "done":
        # the path the network takes since round 2: the side_prep convolutions wrote the head maps (conv_side_tc.cu), the side
        # chain is the up-sampling kernel alone -- reads 8 B per low-res pixel, writes 5 maps + prob + mask
        hs, ws = [int(t_.shape[1]) for t_ in sps], [int(t_.shape[2]) for t_ in sps]
        zs = torch.randn((batch * low, 2), device=dev)
        t = timeit(lambda: ops.side_fwd_heads_done(zs, hs, ws, params, batch, H, W, general=2, want_prob=True, want_mask=True))
        b = (8 * low + 5 * H * W * 4 + H * W * 4 + H * W) * batch
    else:
        t = timeit(lambda: ops.side_fwd(sps, params, H, W, general=2, want_prob=True, want_mask=True))
        b = (16 * low * 2 + 5 * H * W * 4 + H * W * 4 + H * W) * batch
    out.append(f"b{batch}: {t:.1f} us {b / t / 1e3:.0f} GB/s")
print(" | ".join(out))
