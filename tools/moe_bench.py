"""MoE variant (BASELINE config 4: d=512, H=2048, E=8, top-2, M = 8 x 1024 tokens) stage timings on a B200."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

G.build()
import forde_b200  # noqa: E402,F401
from forde_b200 import ops  # noqa: E402

M, d, H, E, k = 8192, 512, 2048, 8, 2
if len(sys.argv) > 1:
    M = int(sys.argv[1])
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((M, d), device="cuda", generator=g).bfloat16()
w_r = (torch.randn((d, E), device="cuda", generator=g) * 0.02).bfloat16()
b_r = torch.zeros(E, device="cuda")
w_up = (torch.randn((E, d, H), device="cuda", generator=g) / d ** 0.5).bfloat16()
b_up = torch.zeros((E, H), device="cuda")
w_down = (torch.randn((E, H, d), device="cuda", generator=g) / H ** 0.5).bfloat16()
b_down = torch.zeros((E, d), device="cuda")
dout = torch.randn((M, d), device="cuda", generator=g).bfloat16()
st = {}


def run():
    st["r"] = ops.moe_route(x, w_r, b_r, k, 0.01)
    st["xg"] = ops.moe_gather(x, st["r"])
    st["zg"], st["hg"], st["yg"] = ops.moe_experts_fwd(st["xg"], w_up, b_up, w_down, b_down, st["r"])
    st["out"] = ops.moe_combine(st["yg"], st["r"])
    st["dyg"] = ops.moe_scatter(dout, st["r"])
    st["dw"] = ops.moe_gate_grad(dout, st["yg"], st["r"])
    st["bwd"] = ops.moe_experts_bwd(st["xg"], st["dyg"], w_up, w_down, st["zg"], st["hg"], st["r"])
    st["dx"] = ops.moe_combine(st["bwd"][0], st["r"], weighted=False)


stages = [
    ("route+plan", lambda: st.__setitem__("r", ops.moe_route(x, w_r, b_r, k, 0.01))),
    ("gather", lambda: ops.moe_gather(x, st["r"])),
    ("experts_fwd (2 grouped GEMMs)", lambda: ops.moe_experts_fwd(st["xg"], w_up, b_up, w_down, b_down, st["r"])),
    ("combine", lambda: ops.moe_combine(st["yg"], st["r"])),
    ("scatter", lambda: ops.moe_scatter(dout, st["r"])),
    ("gate_grad", lambda: ops.moe_gate_grad(dout, st["yg"], st["r"])),
    ("experts_bwd (4 grouped GEMMs + db)", lambda: ops.moe_experts_bwd(st["xg"], st["dyg"], w_up, w_down, st["zg"], st["hg"], st["r"])),
    ("combine dx", lambda: ops.moe_combine(st["bwd"][0], st["r"], weighted=False)),
]
for _ in range(3):
    run()
torch.cuda.synchronize()
tot = 0.0
print(f"MoE slot M={M} d={d} H={H} E={E} top-{k}: routed rows {M * k}, capacity {st['r']['R']}")
for name, f in stages:
    t = 0.0
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        t += a.elapsed_time(b)
    us = 1e3 * t / 20
    tot += us
    print(f"  {name:36s} {us:8.1f} us")
fl = 12.0 * M * k * d * H
print(f"  total fwd+bwd {tot:8.1f} us ; routed-token GEMM work 12*M*k*d*H = {fl / 1e9:.1f} GFLOP -> {fl / (tot * 1e-6) / 1e12:.1f} TFLOP/s")
print("  (stage times above are eager launches from Python with the GPU idle in between: they include the launch latency of")
print("   every kernel of the stage; the device-side cost is the CUDA-graph replay below)")
# the whole slot (fwd + bwd, 25 kernels) as ONE CUDA graph: nothing in it synchronises with the host (the routing plan
# stays on the device), so it replays back to back
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    run()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        run()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    a.record()
    for _ in range(n):
        graph.replay()
    b.record(); torch.cuda.synchronize()
    us = 1e3 * a.elapsed_time(b) / n
print(f"  whole slot fwd+bwd replayed as one CUDA graph: {us:8.1f} us -> {fl / (us * 1e-6) / 1e12:.1f} TFLOP/s on the routed work")

if "--profile" in sys.argv:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            run()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=20, max_name_column_width=70))
