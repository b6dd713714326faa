"""Data-parallel offline parent training (BASELINE configs[4]) on N GPUs of one box:
   torchrun --nproc-per-node N tools/offline_dp_check.py
Every rank fine-tunes on ITS frame for avg_grad_every_n / N micro-iterations with deep supervision, the flat fp32 gradient
is all-reduced over NCCL, every rank applies the same fused SGD step.  Rank 0 checks the result against the same global
batch processed by one trainer, and reports iterations/s and the exposed all-reduce time."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import fosvos_b200 as FB
from fosvos_b200 import sharding, synth
from fosvos_b200.online import OnlineTrainer

rank, world = sharding.init_distributed("nccl")
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
H, W = (int(os.environ.get("DP_H", "120")), int(os.environ.get("DP_W", "214")))
n = 2 * world
sd = synth.make_state_dict(0, "structured")
frames = [synth.make_frame(r, 0, H, W) for r in range(world)]


def make(dp):
    net = FB.OSVOS_VGG(pretrained=0)
    net.load_state_dict(sd)
    net = net.to(dev)
    net.precision = os.environ.get("FOSVOS_PRECISION", "bf16")
    opt = FB.get_optimizer_offline(net, learning_rate=1e-6)
    return net, OnlineTrainer(net, H, W, avg_grad_every_n=n, optimizer=opt, use_graph=True, deep_supervision=0.75,
                              data_parallel=dp, world_size=world)


net, tr = make(True)
x, m = frames[rank]
tr.set_frame(x.to(dev), m.to(dev))
tr.run(n // world)                       # one optimizer step
torch.cuda.synchronize()
ok = True
if rank == 0:
    net1, tr1 = make(False)
    for r in range(world):
        xr, mr = frames[r]
        tr1.set_frame(xr.to(dev), mr.to(dev))
        tr1.run(n // world)
    torch.cuda.synchronize()
    worst = 0.0
    for (k, a), (_, b) in zip(net.state_dict().items(), net1.state_dict().items()):
        d0 = (b.cpu() - sd[k]).abs().max().item()
        d = (a - b).abs().max().item()
        worst = max(worst, d / (d0 + 1e-12) if d0 > 0 else d)
    print(f"DP({world} ranks) vs single trainer: worst |dW_dp - dW_1| / |dW_1| = {worst:.3e}", flush=True)
    ok = worst < 2e-2
# timing: steps/s with and without the exchange
sharding.barrier()
for dp_name, reps in (("dp", 10),):
    torch.cuda.synchronize(); sharding.barrier()
    t0 = time.perf_counter()
    tr.run(reps * (n // world))
    torch.cuda.synchronize(); sharding.barrier()
    dt = time.perf_counter() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        sharding.allreduce_flat(tr.flat_grad)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{world} ranks, {H}x{W}: {reps * n / dt:.1f} frames/s ({reps / dt:.1f} optimizer steps/s); all-reduce of "
              f"{tr.flat_grad.numel() * 4 / 1e6:.1f} MB fp32 = {e0.elapsed_time(e1) / 20:.3f} ms "
              f"({tr.flat_grad.numel() * 4 / 1e9 / (e0.elapsed_time(e1) / 20 / 1e3):.0f} GB/s algorithmic)", flush=True)
if dist.is_initialized():
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
