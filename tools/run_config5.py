"""BASELINE config 5 in miniature: FORDE ViT-L/14 dual encoder, data parallel, in-kernel sensing every step, pooled
moments all-reduced across ranks and the slow loop (Sense -> Cluster -> Smooth -> Actuate -> Reset) every N steps.
Launch:  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/run_config5.py
         [--config vit_l14] [--batch 256] [--steps 45] [--interval 20]"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="vit_l14")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=45)
ap.add_argument("--interval", type=int, default=20)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import forde_b200  # noqa: E402,F401
from forde_b200 import dual_encoder, slow_loop, train  # noqa: E402

cfg = dual_encoder.CONFIGS[a.config]()
st = train.create_train_state(cfg, 0, device=torch.device("cuda", local))
b = {k: v.cuda() for k, v in train.synthetic_batch(cfg, a.batch, seed=42 + rank).items()}
g = train.GraphedTrainStep(st, b)
torch.cuda.synchronize()
t_fast, t_slow, censuses = 0.0, [], []
for step in range(a.steps):
    t0 = time.perf_counter()
    _, m = g(b)
    torch.cuda.synchronize()
    t_fast += time.perf_counter() - t0
    if step > 0 and step % a.interval == 0:
        t0 = time.perf_counter()
        mv = {"stats_buffer": st.variables["stats_buffer"], "neuron_assignments": st.variables["neuron_assignments"]}
        _, _, diag = slow_loop.slow_loop_step(st.params, mv, slow_loop.SlowLoopConfig(), key=step, step=step, group=st.group)
        torch.cuda.synchronize()
        t_slow.append(time.perf_counter() - t0)
        first = diag["layers"]["vision/layer_0/stateful_layer"]
        censuses.append((step, first["census"], first["changed"]))
if rank == 0:
    print(f"config {cfg.name}: {world} GPU(s) x {a.batch} pairs, {a.steps} fast steps {1e3 * t_fast / a.steps:.1f} ms each, "
          f"slow loop every {a.interval} steps: {[round(t, 2) for t in t_slow]} s each ({len(st.model.layers)} layers, GMM on host)")
    for s_, c, ch in censuses:
        print(f"  step {s_}: vision/layer_0 census {c} (relu, tanh, step), {ch} neurons re-assigned; loss {float(m['loss'].item()):.3f}")
if world > 1:
    torch.cuda.synchronize(); dist.barrier(); os._exit(0)
