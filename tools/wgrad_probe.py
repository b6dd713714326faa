"""Per-layer timing of the tcgen05 weight-gradient kernel at the fine-tune window's shapes (batch 5, 480x854 frame).
usage: python tools/wgrad_probe.py [reps]   (FOSVOS_WG_NO_CTA_PAIR=1 selects the single-CTA kernel for the wide layers)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fosvos_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
ONLY = os.environ.get("PROBE_LAYERS")
LAYERS = [("conv1_1", 8, 64, 480, 854), ("side2", 128, 16, 240, 427), ("side3", 256, 16, 120, 214), ("side4", 512, 16, 60, 107),
          ("side5", 512, 16, 30, 54), ("conv1_2", 64, 64, 480, 854), ("conv2_1", 64, 128, 240, 427), ("conv2_2", 128, 128, 240, 427),
          ("conv3_1", 128, 256, 120, 214), ("conv3_2", 256, 256, 120, 214), ("conv4_1", 256, 512, 60, 107),
          ("conv4_2", 512, 512, 60, 107), ("conv5_1", 512, 512, 30, 54)]
if ONLY:
    LAYERS = [l for l in LAYERS if l[0] in ONLY.split(",")]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tot = 0.0
for name, cin, cout, h, w in LAYERS:
    x = torch.randn(5, h, w, cin, device=dev).to(torch.bfloat16)
    dz = torch.randn(5, h, w, cout, device=dev).to(torch.bfloat16)
    ws = ops.wgrad_workspace(cin, cout, dev)
    db = None if os.environ.get("PROBE_NO_BIAS") else torch.zeros(cout, device=dev)
    for _ in range(3):
        ops.conv3x3_wgrad_accumulate(x, dz, ws, db, cout)
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.conv3x3_wgrad_accumulate(x, dz, ws, db, cout)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    med = ts[len(ts) // 2]
    gf = 2 * 9 * cin * cout * 5 * h * w / 1e9
    tot += med
    print(f"{name}: {cin}->{cout} {h}x{w}  median {med:7.1f} us  min {ts[0]:7.1f} us  {gf / med * 1e3:7.1f} TFLOP/s", flush=True)
print(f"sum of medians: {tot:.1f} us")
