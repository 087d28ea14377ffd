"""Timeline of the captured train step (CUDA-graph replay) from the CUPTI kernel records of torch.profiler: per stream
busy time, time with NO kernel running on any stream, the distribution of the gaps between consecutive kernels and the
kernels that follow the longest gaps.  Usage (on a B200): python tools/gap_probe.py [--steps 3] > gpurun_out/gaps.txt"""
import argparse
import json
import os
import sys
import tempfile

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G  # noqa: E402

G.build()
import forde_b200  # noqa: E402,F401
from forde_b200 import dual_encoder, train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="vit_b16")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
cfg = dual_encoder.CONFIGS[a.config]()
st = train.create_train_state(cfg, 0, device="cuda")
b = {k: v.cuda() for k, v in train.synthetic_batch(cfg, a.batch).items()}
g = train.GraphedTrainStep(st, b)
for _ in range(3):
    g(b)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        g(b)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
print(f"{len(ev)} kernels over {a.steps} replays, {len(ev) / a.steps:.0f} per step; span {(t1 - t0) / a.steps / 1e3:.2f} ms per step")
streams = {}
for e in ev:
    streams.setdefault(e["args"].get("stream"), []).append(e)
for s, es in streams.items():
    print(f"stream {s}: {len(es) / a.steps:.0f} kernels/step, busy {sum(e['dur'] for e in es) / a.steps / 1e3:.2f} ms per step")
# union of busy intervals over all streams
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev)
busy, cur_s, cur_e, gaps = 0.0, iv[0][0], iv[0][1], []
for s, e in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, s))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
idle = (t1 - t0) - busy
print(f"any-stream busy {busy / a.steps / 1e3:.2f} ms per step, idle (no kernel on any stream) {idle / a.steps / 1e3:.2f} ms per step "
      f"in {len(gaps) / a.steps:.0f} gaps")
import collections
h = collections.Counter()
for d, _ in gaps:
    h[min(int(d), 20)] += 1
print("gap length (us, 20 = 20 or more): count per step")
for k in sorted(h):
    print(f"  {k:3d}: {h[k] / a.steps:7.1f}   ({k * h[k] / a.steps / 1e3:.3f} ms)")
# per-stream gaps on the main (busiest) stream: what follows the longest ones
main = max(streams.values(), key=len)
main.sort(key=lambda e: e["ts"])
pg = []
for p, n in zip(main, main[1:]):
    pg.append((n["ts"] - p["ts"] - p["dur"], p["name"][:60], n["name"][:60]))
pg.sort(reverse=True)
print("longest gaps on the main stream (us, previous kernel -> next kernel):")
for d, pn, nn in pg[:25]:
    print(f"  {d:8.1f}  {pn}  ->  {nn}")
tot = sum(d for d, _, _ in pg if d > 0)
print(f"sum of positive gaps on the main stream: {tot / a.steps / 1e3:.2f} ms per step; median gap {sorted(d for d, _, _ in pg)[len(pg) // 2]:.2f} us")
