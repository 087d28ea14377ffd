"""HBM probes that bound the fused epilogues: write-only fill, read+write copy, at the zT tensor's size."""
import torch
H, M = 3072, 50432
x = torch.empty((H, M), dtype=torch.bfloat16, device="cuda")
y = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(f, n=10):
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return 1e3 * tot / n
nb = x.numel() * 2
for name, f, bytes_ in [("fill (write only)", lambda: x.zero_(), nb), ("copy (read+write)", lambda: y.copy_(x), 2 * nb),
                        ("fill 2 tensors", lambda: (x.zero_(), y.zero_()), 2 * nb),
                        ("sum (read only)", lambda: x.view(torch.int16).sum(), nb)]:
    us = t(f)
    print(f"{name:22s} {us:8.1f} us  {bytes_ / us / 1e6:7.2f} TB/s")
