"""Kernel-time breakdown of one dual-encoder train step (torch.profiler, CUDA activities) -> text table.
Usage (on a B200): python tools/profile_step.py [--config vit_b16] [--batch 256] > gpurun_out/step_profile.txt"""
import argparse
import os
import sys

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
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
cfg = dual_encoder.CONFIGS[a.config]()
st = train.create_train_state(cfg, 0, device="cuda")
b = {k: v.cuda() for k, v in train.synthetic_batch(cfg, a.batch).items()}
for _ in range(3):
    train.train_step(st, b)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        train.train_step(st, b)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
