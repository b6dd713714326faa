"""Side-chain and loss kernels alone: GPU time (CUDA graph of `reps` calls) and HBM GB/s.  python tools/side_probe.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fosvos_b200 import ops
from fosvos_b200.layers import interp_surgery  # noqa: F401

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W = 480, 854
dev = torch.device("cuda:0")


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


import fosvos_b200 as FB
net = FB.OSVOS_VGG(pretrained=0).to(dev)
params = net._side()
h, w = H, W
for dt in (torch.bfloat16, torch.float32):
    sps = []
    h, w = H, W
    for i in range(4):
        h, w = (h + 1) // 2, (w + 1) // 2
        sps.append(torch.randn((batch, h, w, 16), device=dev).to(dt))
    low = sum(t.shape[1] * t.shape[2] for t in sps)
    t = timeit(lambda: ops.side_fwd(sps, params, H, W, general=False, want_prob=True, want_mask=True))
    b = (16 * low * sps[0].element_size() + 5 * H * W * 4 + H * W * 4 + H * W) * batch
    print(f"side_fwd[{dt}] batch {batch}: {t:.1f} us  {b / t / 1e3:.0f} GB/s ({b / 1e6:.1f} MB)")
    if ops.side_separable(params):
        t = timeit(lambda: ops.side_fwd(sps, params, H, W, general=2, want_prob=True, want_mask=True))
        print(f"side_fwd[{dt}] separable path: {t:.1f} us  {b / t / 1e3:.0f} GB/s")
    t = timeit(lambda: ops.side_fwd(sps, params, H, W, general=False))
    b = (16 * low * sps[0].element_size() + 5 * H * W * 4) * batch
    print(f"side_fwd[{dt}] no prob/mask: {t:.1f} us  {b / t / 1e3:.0f} GB/s")
x = torch.randn((batch, 1, H, W), device=dev)
lab = (torch.rand((batch, 1, H, W), device=dev) > 0.8).float()
loss, stats = ops.bal_loss_fwd(x, lab, False)
t = timeit(lambda: ops.bal_loss_fwd(x, lab, False))
print(f"bal_loss_fwd batch {batch}: {t:.1f} us  {8 * x.numel() / t / 1e3:.0f} GB/s")
dx = torch.empty_like(x)
t = timeit(lambda: ops.bal_loss_bwd(x, lab, False, stats, None, 1.0, out=dx))
print(f"bal_loss_bwd batch {batch}: {t:.1f} us  {12 * x.numel() / t / 1e3:.0f} GB/s")
t = timeit(lambda: ops.bal_loss_fwd_bwd(x, lab, False, stats, None, 1.0, out=dx))
print(f"bal_loss_fwd_bwd (one pass) batch {batch}: {t:.1f} us  {12 * x.numel() / t / 1e3:.0f} GB/s")
