"""Device-memory bandwidth probes (write-only / read-only / copy) -- the denominators for the HBM-bound kernels."""
import torch

n = 1 << 30
x = torch.empty(n, dtype=torch.uint8, device="cuda")
y = torch.empty(n, dtype=torch.uint8, device="cuda")
xi = x.view(torch.int32)


def t(fn, reps=10):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


w = t(lambda: x.zero_())
c = t(lambda: y.copy_(x))
r = t(lambda: xi.sum())
print("write-only %.0f GB/s | copy (read+write) %.0f GB/s | read-only (int32 sum) %.0f GB/s" % (n / w / 1e6, 2 * n / c / 1e6, n / r / 1e6))
