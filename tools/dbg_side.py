import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import test_gpu_kernels as T
from fosvos_b200 import ops
from oracle import osvos_oracle as O
for rep in range(3):
  for HW in [(48, 72), (45, 70), (33, 17)]:
    for dt in (torch.float32, torch.bfloat16):
        H, W = HW
        sp, sd = T._side_inputs(2, H, W, 11, dt)
        params, dsd = T._side_params(sd)
        outs, prob, mask = ops.side_fwd([T._nhwc(t, dt) for t in sp], params, H, W, general=False, want_prob=True, want_mask=True)
        p = O.probabilities(outs[4].cpu())
        d = (prob.cpu() - p).abs()
        i = d.argmax()
        print(rep, HW, dt, "max|dp|", float(d.max()), "at logit", float(outs[4].cpu().flatten()[i]), "p", float(p.flatten()[i]), float(prob.cpu().flatten()[i]))
