"""Measured HBM streams on this GPU: pure write (fill), pure read (sum), copy -- context for the write-heavy kernels."""
import torch
dev = torch.device("cuda:0")
n = 1 << 28            # 1 GiB of fp32
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)


def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


ms = t(lambda: a.zero_())
print(f"fill  (write only): {4 * n / ms / 1e6:.0f} GB/s")
ms = t(lambda: a.sum())
print(f"sum   (read only) : {4 * n / ms / 1e6:.0f} GB/s")
ms = t(lambda: b.copy_(a))
print(f"copy  (read+write): {8 * n / ms / 1e6:.0f} GB/s")
small = torch.empty(82_000_000 // 4, device=dev)
ms = t(lambda: small.zero_(), reps=30)
print(f"fill 82 MB (fits L2): {82 / ms:.0f} GB/s ({ms * 1e3:.1f} us)")
