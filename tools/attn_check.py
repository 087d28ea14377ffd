"""Relative errors of the attention kernels against fp32 softmax attention (quick diagnostic, no asserts)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G
G.build()
from forde_b200 import ops

def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))

for (B, T, heads, blocked) in [(2, 197, 12, True), (3, 256, 2, True), (2, 129, 4, True), (4, 77, 8, True), (256, 197, 12, True)]:
    w = heads * 64
    g_ = torch.Generator(device="cuda").manual_seed(T * 131 + B)
    q, k, v = ((torch.randn((B, T, w), device="cuda", generator=g_) * s).bfloat16() for s in (1.5, 1.5, 1.0))
    d_o = torch.randn((B, T, w), device="cuda", generator=g_).bfloat16()
    q2, k2, v2, d2 = (t.view(B * T, w) for t in (q, k, v, d_o))
    o, lse = ops.attn_fwd(q2, k2, v2, B, T, heads, blocked=blocked)
    torch.cuda.synchronize()
    o_keep = o.clone()
    dq, dk, dv = ops.attn_bwd(q2, k2, v2, d2, lse, B, T, heads, o=o, blocked=blocked)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().view(B, T, heads, 64).transpose(1, 2).detach().requires_grad_(True) for t in (q, k, v))
    s = (qf @ kf.transpose(-1, -2)) * 0.125
    ref = torch.softmax(s, -1) @ vf
    ref.backward(d_o.float().view(B, T, heads, 64).transpose(1, 2))
    back = lambda t: t.transpose(1, 2).reshape(B * T, w)
    print((B, T, heads), "o", rel(o_keep, back(ref)), "o after bwd", rel(o, back(ref)), "dq", rel(dq, back(qf.grad)),
          "dk", rel(dk, back(kf.grad)), "dv", rel(dv, back(vf.grad)), "lse", rel(lse, torch.logsumexp(s, -1)))
