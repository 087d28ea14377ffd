"""BASELINE config 4: FORDE LLM decoder (d=512, 12 layers, 8 experts top-2, H=2048, seq 1024, batch 8, vocab 50257;
src/forde/model.py:43-55, src/training/train.py:241) with the MoE slot on this repository's kernels: train-step time
and tokens/s on synthetic tokens (eager launches; a record for profiles/, not bench.py's metric)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G  # noqa: E402

G.build()
import forde_b200  # noqa: E402,F401
from forde_b200 import llm_harness as LM, ops  # noqa: E402

# everything runs on one non-default stream from the start: autograd's AccumulateGrad nodes remember the stream they
# were created on, and a CUDA-graph capture must not touch the default stream
_stream = torch.cuda.Stream()
torch.cuda.set_stream(_stream)
B, S = 8, 1024
cfg = LM.LLMConfig(vocab_size=50257, d_model=512, num_layers=12, num_heads=8, head_dim=64, max_seq_len=S,
                   num_experts=8, top_k_experts=2, expert_hidden_dim=2048)
model = LM.FORDEDecoderLM(cfg)
v = model.init(0, "cuda")
ids = torch.randint(0, cfg.vocab_size, (B, S), device="cuda", dtype=torch.int32)
opt = torch.optim.AdamW(list(LM.param_leaves(v["params"])), lr=1e-4, fused=True, capturable=True)


def step():
    (logits, aux), _ = model.apply(v, ids, mutable=["stats_buffer"])
    loss, _ = LM.lm_loss(logits, ids, aux)
    opt.zero_grad(set_to_none=False)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(list(LM.param_leaves(v["params"])), 1.0)
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
l0 = ops.launch_count()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
n = 10
for _ in range(n):
    loss = step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
print(f"FORDE LLM (config 4) B={B} S={S} d=512 L=12 E=8 top-2 H=2048 V=50257: {ms:.1f} ms/step, {B * S / ms * 1e3:.0f} tokens/s, "
      f"loss {float(loss.item()):.3f}, {(ops.launch_count() - l0) // n} kernels of this repository per step (eager launches)")
# the same step replayed as one CUDA graph (the routing plan never leaves the device, so nothing in the step syncs)
try:
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=_stream):
        gloss = step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        graph.replay()
    b.record(); torch.cuda.synchronize()
    msg = a.elapsed_time(b) / n
    print(f"  replayed as one CUDA graph: {msg:.1f} ms/step, {B * S / msg * 1e3:.0f} tokens/s, loss {float(gloss.item()):.3f}")
except Exception as e:  # noqa: BLE001
    print(f"  CUDA graph capture of the LLM step failed: {type(e).__name__}: {str(e)[:200]}")
