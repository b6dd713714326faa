"""maxpool backward timing per stage shape:  python tools/pool_probe.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fosvos_b200 import ops
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
tot = 0.0
for (h, w, c) in [(480, 854, 64), (240, 427, 128), (120, 214, 256), (60, 107, 512)]:
    x = torch.randn((batch, h, w, c), device=dev).clamp_min(0).to(torch.bfloat16)
    dy = torch.randn((batch, (h + 1) // 2, (w + 1) // 2, c), device=dev).to(torch.bfloat16)
    for _ in range(2):
        ops.maxpool2x2_bwd(x, dy)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.maxpool2x2_bwd(x, dy)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10 * 1e3
    b = (2 * x.numel() + dy.numel()) * 2
    tot += t
    print(f"{h}x{w}x{c} batch {batch}: {t:.1f} us  {b / t / 1e3:.0f} GB/s")
print(f"total {tot:.1f} us")
