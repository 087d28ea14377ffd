"""Time this repository's attention kernels (csrc/forde_attn.cu, forde_attn2.cu) against cuDNN flash SDPA at the towers'
shapes: isolated launches, CUDA events, L2 flushed between launches.  python tools/attn_bench.py [B]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G  # noqa: E402

G.build()
from forde_b200 import ops  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(512 * 2 ** 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for name, T, heads in (("vision ViT-B/16", 197, 12), ("text", 77, 8), ("T=256", 256, 12), ("T=129", 129, 12)):
    w = heads * 64
    g = torch.Generator(device="cuda").manual_seed(1)
    q, k, v, d_o = ((torch.randn((B * T, w), device=dev, generator=g)).bfloat16() for _ in range(4))
    o, lse = ops.attn_fwd(q, k, v, B, T, heads)
    t_f = timeit(lambda: ops.attn_fwd(q, k, v, B, T, heads))
    t_b = timeit(lambda: ops.attn_bwd(q, k, v, d_o, lse, B, T, heads, o=o))
    q4, k4, v4 = (t.view(B, T, heads, 64).transpose(1, 2).detach().requires_grad_(True) for t in (q, k, v))
    d4 = d_o.view(B, T, heads, 64).transpose(1, 2)
    with torch.nn.attention.sdpa_kernel(torch.nn.attention.SDPBackend.CUDNN_ATTENTION):
        t_cf = timeit(lambda: F.scaled_dot_product_attention(q4, k4, v4))
        out = F.scaled_dot_product_attention(q4, k4, v4)
        t_cb = timeit(lambda: torch.autograd.grad(out, (q4, k4, v4), d4, retain_graph=True))
    fl = 4.0 * B * heads * T * T * 64
    mb = B * T * w * 2 / 1e6
    print(f"{name:16s} B={B} T={T} heads={heads}: own fwd {t_f:7.1f} us ({fl / t_f / 1e6:6.1f} TF/s, {4 * mb / t_f * 1e-3 * 1e3:5.0f} GB/s)"
          f"  bwd {t_b:7.1f} us ({2.5 * fl / t_b / 1e6:6.1f} TF/s, {8 * mb / t_b:5.2f} TB/s) | cuDNN fwd {t_cf:7.1f} bwd {t_cb:7.1f} us")
