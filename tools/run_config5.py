"""BASELINE config 5 in miniature: FORDE ViT-L/14 dual encoder, data parallel, in-kernel sensing every step, pooled
moments all-reduced across ranks and the slow loop (Sense -> Cluster -> Smooth -> Actuate -> Reset) every N steps.
Launch:  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/run_config5.py
         [--config vit_l14] [--batch 256] [--steps 45] [--interval 20] [--overlap K]
--overlap K > 0: Cluster + Smooth on a worker thread, maps actuated K steps after sensing (slow_loop.OverlappedSlowLoop);
the "slow loop" times printed are then only the Sense + Reset part that stays on the training thread."""
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
ap.add_argument("--overlap", type=int, default=0)
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
over = slow_loop.OverlappedSlowLoop(slow_loop.SlowLoopConfig(), a.overlap, st.group) if a.overlap > 0 else None
t_wall = time.perf_counter()


def note(diag):
    if diag is not None:
        first = diag["layers"]["vision/layer_0/stateful_layer"]
        censuses.append((diag["step"], first["census"], first["changed"]))


for step in range(a.steps):
    if over is not None:
        note(over.poll(step))
    t0 = time.perf_counter()
    _, m = g(b)
    torch.cuda.synchronize()
    t_fast += time.perf_counter() - t0
    if step > 0 and step % a.interval == 0:
        t0 = time.perf_counter()
        mv = {"stats_buffer": st.variables["stats_buffer"], "neuron_assignments": st.variables["neuron_assignments"]}
        if over is not None:
            note(over.start(mv, key=step, step=step))
        else:
            _, _, diag = slow_loop.slow_loop_step(st.params, mv, slow_loop.SlowLoopConfig(), key=step, step=step, group=st.group)
            note(diag)
        torch.cuda.synchronize()
        t_slow.append(time.perf_counter() - t0)
if over is not None:
    note(over.flush())
    over.close()
t_wall = time.perf_counter() - t_wall
if rank == 0:
    print(f"config {cfg.name}: {world} GPU(s) x {a.batch} pairs, {a.steps} fast steps {1e3 * t_fast / a.steps:.1f} ms each, "
          f"slow loop every {a.interval} steps: {[round(t, 2) for t in t_slow]} s each on the training thread "
          f"({len(st.model.layers)} layers, GMM on host, overlap {a.overlap}); whole loop {t_wall:.2f} s wall")
    for s_, c, ch in censuses:
        print(f"  step {s_}: vision/layer_0 census {c} (relu, tanh, step), {ch} neurons re-assigned; loss {float(m['loss'].item()):.3f}")
if world > 1:
    torch.cuda.synchronize(); dist.barrier(); os._exit(0)
