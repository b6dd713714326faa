"""Timing of the one-launch conv optimizer step (fold + SGD + repack, step.cu) over the 17 convs of OSVOS-VGG."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fosvos_b200 as FB
from fosvos_b200 import synth

dev = torch.device("cuda:0")
net = FB.OSVOS_VGG(pretrained=0)
net.load_state_dict(synth.make_state_dict(0, "parent"))
net = net.to(dev)
x, m = synth.make_frame(0, 0, 96, 128)
opt = FB.get_optimizer_online(net)
from fosvos_b200.online import OnlineTrainer
tr = OnlineTrainer(net, 96, 128, 5, opt, use_graph=False)
tr.set_frame(x.to(dev), m.to(dev))
for _ in range(2):
    tr.run(5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
from fosvos_b200 import ops
sh = tr._shared
ts = []
for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.conv_step_all(sh["conv_table"], opt.param_groups[0].get("momentum", 0.9))
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
print(f"conv_step_all: median {ts[len(ts) // 2]:.1f} us, min {ts[0]:.1f} us")
