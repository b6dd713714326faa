"""Small fixed workload for ncu: `reps` x (one 8-frame inference batch + one fine-tune accumulation window of
5 micro-iterations + one optimizer step) at 480x854, bf16.  Used for the launch list and the --set full captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fosvos_b200 as FB
from fosvos_b200 import synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mode = sys.argv[3] if len(sys.argv) > 3 else "both"      # both | ft | inf
dev = torch.device("cuda:0")
net = FB.OSVOS_VGG(pretrained=0)
net.load_state_dict(synth.make_state_dict(0, "structured"))
net = net.to(dev)
net.precision = os.environ.get("FOSVOS_PRECISION", "bf16")
x, m = synth.make_frame(0, 0, 480, 854)
xb = torch.cat([torch.roll(x, i, 3) for i in range(batch)]).to(dev)
x, m = x.to(dev), m.to(dev)
opt = FB.get_optimizer_online(net)
for _ in range(reps):
    if mode in ("both", "inf"):
        net.predict(xb)
    if mode in ("both", "ft"):
        FB.finetune(net, x, m, 5, 5, optimizer=opt)      # one accumulation window (5 micro-iterations, one batched pass) + step
torch.cuda.synchronize()
print("profile_step ok")
