"""One tcgen05 conv layer, a few launches (for ncu):  python tools/one_conv.py H W Cin Cout [batch] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fosvos_b200 import _lib as L
from fosvos_b200 import ops

h, w, cin, cout = (int(a) for a in sys.argv[1:5])
batch = int(sys.argv[5]) if len(sys.argv) > 5 else 1
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
dev = torch.device("cuda:0")
x = torch.randn((batch, h, w, cin), device=dev).to(torch.bfloat16)
kpad = (cin + 63) // 64 * 64
wp = (torch.randn((cout * 9 * kpad,), device=dev) * 0.05).to(torch.bfloat16)
bias = torch.zeros(cout, device=dev)
y = torch.empty((batch, h, w, cout), device=dev, dtype=torch.bfloat16)
for _ in range(reps):
    ops.conv3x3(x, wp, bias, cout, L.CONV_BIAS | L.CONV_RELU, out=y)
torch.cuda.synchronize()
print("one_conv ok")
