"""Isolated StatefulLayer micro-benchmark at a BASELINE config's layer dims (default: C2 vision,
M = 256 x 197, d = 768, H = 3072): per-kernel CUDA-event times (L2 flushed between iterations) and the
fwd+bwd fraction of the measured burst bf16 peak.  Also the short command used for `ncu --set full`.
Usage: python tools/layer_bench.py [--M 50432 --d 768 --H 3072] [--iters 20] [--once]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

G.build()
import forde_b200  # noqa: E402,F401
from forde_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=50432)
ap.add_argument("--d", type=int, default=768)
ap.add_argument("--H", type=int, default=3072)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--once", action="store_true", help="one pass, no timing (for ncu)")
ap.add_argument("--save-z", type=int, default=0, help="1 = forward writes all of z^T (debug), 0 = only what the backward reads")
ap.add_argument("--sign-bits", type=int, default=1, help="1 = forward records (z > 0) bits, backward reads them instead of h^T")
ap.add_argument("--assign", default="mixed", choices=["mixed", "zeros", "blocks"],
                help="brain map: i.i.d. uniform{0,1,2} | all generalists (init) | contiguous thirds (smoothed map)")
a = ap.parse_args()
M, d, H = a.M, a.d, a.H
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((M, d), device="cuda", generator=g).bfloat16()
w_up = (torch.randn((d, H), device="cuda", generator=g) / d ** 0.5).bfloat16()
w_down = (torch.randn((H, d), device="cuda", generator=g) / H ** 0.5).bfloat16()
b_up = torch.randn(H, device="cuda", generator=g) * 0.1
b_down = torch.randn(d, device="cuda", generator=g) * 0.1
assign = torch.randint(0, 3, (H,), device="cuda", generator=g).to(torch.int8)
if a.assign == "zeros":
    assign.zero_()
elif a.assign == "blocks":
    assign = (torch.arange(H, device="cuda") * 3 // H).to(torch.int8)
dy = torch.randn((M, d), device="cuda", generator=g).bfloat16()
ld, P = ops.sample_pitch(M), ops.num_partials(M)
zT = torch.zeros((H, ld), device="cuda", dtype=torch.bfloat16); hT = torch.zeros_like(zT); dzT = torch.zeros_like(zT)
actp = torch.zeros((P, 3, H), device="cuda"); gradp = torch.zeros((P, 3, H), device="cuda")
y = torch.zeros((M, d), device="cuda", dtype=torch.bfloat16); dx = torch.zeros_like(y)
dw_up = torch.zeros((d, H), device="cuda"); dw_down = torch.zeros((H, d), device="cuda")
db_up = torch.zeros(H, device="cuda"); db_down = torch.zeros(d, device="cuda")
stats = torch.zeros((H, 5), device="cuda")
signs = torch.zeros(ops.sign_shape(M, H), device="cuda", dtype=torch.int32) if a.sign_bits else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

steps = [
    ("up_fwd", lambda: ops.sl_up_fwd(x, w_up, b_up, assign, zT, hT, actp, save_z=bool(a.save_z), sign_bits=signs)),
    ("down_fwd", lambda: ops.sl_down_fwd(hT, w_down, b_down, y, M)),
    ("dz_bwd", lambda: ops.sl_dz_bwd(dy, w_down, zT, hT, assign, dzT, gradp, sign_bits=signs)),
    ("dx_bwd", lambda: ops.sl_dx_bwd(dzT, w_up, dx, M)),
    ("dw_up", lambda: ops.sl_dw_up(x, dzT, dw_up)),
    ("dw_down", lambda: ops.sl_dw_down(hT, dy, dw_down)),
    ("db_down", lambda: ops.colsum_bf16(dy, db_down)),
    ("finalize", lambda: ops.stats_finalize(actp, gradp, P, H, M, step_stats=stats, db_up=db_up)),
]
if a.once:
    for _, f in steps:
        f()
    torch.cuda.synchronize()
    sys.exit(0)
for _ in range(3):
    for _, f in steps:
        f()
torch.cuda.synchronize()
res = {}
for name, f in steps:
    tot = 0.0
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    res[name] = 1e3 * tot / a.iters
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
fl = 2.0 * M * d * H
print(f"StatefulLayer M={M} d={d} H={H} assignments={a.assign}  (L2 flushed between launches; peak = {peaks['bf16_tflops']} TF/s burst)")
for name, us in res.items():
    tf = fl / (us * 1e-6) / 1e12 if name not in ("db_down", "finalize") else 0.0
    print(f"  {name:10s} {us:9.1f} us" + (f"  {tf:8.1f} TFLOP/s  {100 * tf / peaks['bf16_tflops']:5.1f}% of burst peak" if tf else ""))
total = sum(res.values())
print(f"  layer fwd+bwd+sensing {total:9.1f} us  {6 * fl / (total * 1e-6) / 1e12:8.1f} TFLOP/s  "
      f"{100 * 6 * fl / (total * 1e-6) / 1e12 / peaks['bf16_tflops']:5.1f}% of burst peak")
