"""Reproduce the roofline denominators of MEASURED_PEAKS.json on the box at hand, with the method it records:
  HBM: b.copy_(a) over 1 Gi bf16 elements, read + write bytes, best of 10, CUDA events;
  bf16: torch.matmul 8192^3 (2 N^3 FLOP), best of 10 ("burst") and back to back for ~4 s ("sustained").
Also the one-directional figures the kernel notes in DESIGN.md lean on: a write-only fill and a read-only reduction of a
tensor the size of one z^T / h^T residual (310 MB at C2 vision), L2 flushed in between.
Usage (on a B200): python tools/bw_probe.py > gpurun_out/bw_probe.txt"""
import json
import os
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda")


def best_ms(fn, reps=10, flush=None):
    best = float("inf")
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {}
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
b = torch.empty_like(a)
for _ in range(3):
    b.copy_(a)
ms = best_ms(lambda: b.copy_(a))
out["hbm_copy_gbs"] = 2 * n * 2 / (ms * 1e-3) / 1e9
del a, b

# one-directional streams over a residual-sized tensor (H x M bf16 at C2 vision), L2 flushed before each repetition
H, M = 3072, 50432
t = torch.empty((H, M), dtype=torch.bfloat16, device=dev).normal_()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out["write_only_fill_gbs"] = t.numel() * 2 / (best_ms(lambda: t.zero_(), flush=flush) * 1e-3) / 1e9
t.normal_()
out["read_only_sum_gbs"] = t.numel() * 2 / (best_ms(lambda: t.view(torch.int16).sum(dtype=torch.int64), flush=flush) * 1e-3) / 1e9
del t, flush

N = 8192
x = torch.randn((N, N), device=dev).bfloat16()
y = torch.randn((N, N), device=dev).bfloat16()
for _ in range(3):
    torch.matmul(x, y)
out["bf16_tflops_burst"] = 2 * N ** 3 / (best_ms(lambda: torch.matmul(x, y)) * 1e-3) / 1e12
torch.cuda.synchronize()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 0
while time.perf_counter() - t0 < 4.0:
    for _ in range(50):
        torch.matmul(x, y)
    reps += 50
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
out["bf16_tflops_sustained"] = reps * 2 * N ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12

path = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(path):
    ref = json.load(open(path))
    out["measured_peaks_json"] = {k: ref.get(k) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained")}
print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in out.items()}, indent=1))
