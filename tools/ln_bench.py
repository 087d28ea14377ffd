"""add+LayerNorm micro-benchmark at ViT-B/16 token dims (rows = 256 x 197, cols = 768)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G  # noqa: E402

G.build()
import forde_b200  # noqa: E402,F401
from forde_b200 import ops  # noqa: E402

_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
rows, cols = (50432, 768) if len(_pos) < 2 else (int(_pos[0]), int(_pos[1]))
x = torch.randn((rows, cols), device="cuda").bfloat16()
r = None if "--no-r" in sys.argv else torch.randn((rows, cols), device="cuda").bfloat16()   # --no-r: the input is already the sum
dy = torch.randn((rows, cols), device="cuda").bfloat16()
g = torch.ones(cols, device="cuda"); b = torch.zeros(cols, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out, mean, rstd = ops.add_layernorm_fwd(x, r, g, b)
if "--once" in sys.argv:
    ops.add_layernorm_bwd(dy, x, r, g, mean, rstd)
    torch.cuda.synchronize()
    sys.exit(0)
nt = 2 if r is None else 3
for name, f, nbytes in [("fwd", lambda: ops.add_layernorm_fwd(x, r, g, b), nt * rows * cols * 2),
                        ("bwd", lambda: ops.add_layernorm_bwd(dy, x, r, g, mean, rstd), (nt + 1) * rows * cols * 2)]:
    for _ in range(3):
        f()
    t = 0.0
    for _ in range(20):
        flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); c.record(); torch.cuda.synchronize()
        t += a.elapsed_time(c)
    us = 1e3 * t / 20
    print(f"add_layernorm {name} rows={rows} cols={cols}: {us:7.1f} us  {nbytes / us / 1e6:5.2f} TB/s of algorithmic bytes")
