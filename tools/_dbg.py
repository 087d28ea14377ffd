import os, sys
import torch
sys.path.insert(0, "/root/repo")
import __graft_entry__ as G
G.build()
from forde_b200 import ops
B,T,heads=2,197,12
w=heads*64
g_ = torch.Generator(device="cuda").manual_seed(1)
q, k, v = ((torch.randn((B, T, w), device="cuda", generator=g_) * s).bfloat16() for s in (1.5, 1.5, 1.0))
o, lse = ops.attn_fwd(q.view(B*T,w), k.view(B*T,w), v.view(B*T,w), B, T, heads, blocked=True)
torch.cuda.synchronize()
qf, kf, vf = (t.float().view(B, T, heads, 64).transpose(1, 2) for t in (q, k, v))
s = (qf @ kf.transpose(-1, -2)) * 0.125
P = torch.softmax(s, -1)
ref = (P @ vf).transpose(1,2).reshape(B,T,heads,64)
og = o.float().view(B,T,heads,64)
for rows in [(0,32),(32,64),(64,128),(128,160),(160,197)]:
    for cols in [(0,32),(32,64)]:
        a=og[0,rows[0]:rows[1],0,cols[0]:cols[1]]; b=ref[0,rows[0]:rows[1],0,cols[0]:cols[1]]
        print(rows, cols, "rel", float((a-b).norm()/b.norm()), "median ratio", float((a/b).median()))
# compare to unnormalised
mx = s.max(-1, keepdim=True).values
un = (torch.exp(s-mx) @ vf).transpose(1,2).reshape(B,T,heads,64)
print("vs unnormalised:", float((og-un).norm()/un.norm()))
# vs keys-limited
for kk in (64,128,192,197):
    Pk = P[..., :kk] @ vf[..., :kk, :]
    print(kk, float((og - Pk.transpose(1,2).reshape(B,T,heads,64)).norm()/ref.norm()))
