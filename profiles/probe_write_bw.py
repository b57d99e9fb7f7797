"""Probe: write-only and copy bandwidth of this B200 with plain torch ops (context for the FK kernel,
whose traffic is 98 % writes).  Run under gpurun; prints GB/s."""
import torch
x = torch.empty(1 << 30, dtype=torch.float64, device="cuda")       # 8 GiB
y = torch.empty(1 << 29, dtype=torch.float64, device="cuda")
z = torch.empty(1 << 29, dtype=torch.float64, device="cuda")
def t(fn, bytes_, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return bytes_ / best / 1e6
print("fill  (write only) GB/s: %.0f" % t(lambda: x.fill_(1.0), x.numel() * 8))
print("copy  (read+write) GB/s: %.0f" % t(lambda: z.copy_(y), 2 * y.numel() * 8))
print("read  (sum)        GB/s: %.0f" % t(lambda: x.sum(), x.numel() * 8))
