"""One tcgen05 weight-gradient layer, a few launches (for ncu):  python tools/one_wgrad.py H W Cin Cout [batch] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fosvos_b200 import ops

h, w, cin, cout = (int(a) for a in sys.argv[1:5])
batch = int(sys.argv[5]) if len(sys.argv) > 5 else 1
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
dev = torch.device("cuda:0")
cinp = ops.pad8(cin)
x = torch.randn((batch, h, w, cinp), device=dev).to(torch.bfloat16)
dz = (torch.randn((batch, h, w, cout), device=dev) * 0.1).to(torch.bfloat16)
ws = ops.wgrad_workspace(cinp, ops.pad8(cout), dev)
db = torch.zeros(cout, device=dev)
for _ in range(reps):
    ops.conv3x3_wgrad_accumulate(x, dz, ws, db, cout)
torch.cuda.synchronize()
print("one_wgrad ok")
