"""clock64 timeline of CTA 0 of the blocked attention forward and backward (debug hook forde_attn2_set_trace)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G
G.build()
from forde_b200 import ops, _lib
lib = _lib.load()
lib.forde_attn2_set_trace.argtypes = [ctypes.c_void_p]
B, T, heads = 64, 197, 12
w = heads * 64
g = torch.Generator(device="cuda").manual_seed(1)
q, k, v, d_o = ((torch.randn((B * T, w), device="cuda", generator=g)).bfloat16() for _ in range(4))
o, lse = ops.attn_fwd(q, k, v, B, T, heads)
for _ in range(3):
    ops.attn_fwd(q, k, v, B, T, heads)
dbg = torch.zeros(16 * 16, dtype=torch.int64, device="cuda")
lib.forde_attn2_set_trace(dbg.data_ptr())
ops.attn_fwd(q, k, v, B, T, heads)
torch.cuda.synchronize()
lib.forde_attn2_set_trace(None)
d = dbg.cpu().view(16, 16)
t0 = int(d[0, 0])
fnames = ["start", "S ready", "max done", "P staged", "O ready", "O stored"]
print("forward (cycles from the first item's start; CTA 0, first thread of each row block's warp group)")
for n in range(6):
    for blk in range(2):
        print(f"item {n} block {blk}: " + "  ".join(f"{fnames[i]}={int(d[n, blk * 8 + i]) - t0:6d}" for i in range(6)))
print("backward")
for _ in range(3):
    ops.attn_bwd(q, k, v, d_o, lse, B, T, heads, o=o)
dbg = torch.zeros(16 * 16, dtype=torch.int64, device="cuda")
lib.forde_attn2_set_trace(dbg.data_ptr())
ops.attn_bwd(q, k, v, d_o, lse, B, T, heads, o=o)
torch.cuda.synchronize()
lib.forde_attn2_set_trace(None)
d = dbg.cpu().view(16, 16)
t0 = int(d[0, 0])
names = ["start", "D/lse ready", "sdp ready", "chunk 0 done", "bar_g ok", "stored+arrive"]
for gidx in range(12):
    print(f"tile {gidx:2d}: " + "  ".join(f"{names[i]}={int(d[gidx, i]) - t0:6d}" for i in range(6)))
