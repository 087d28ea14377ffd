"""Fast loop + slow loop together, through train.fit (the reference's outer loop, src/training/train.py:320-399):
throughput INCLUDING the slow-loop stalls, at the reference's cadence (--interval 100, train.py:219-224).
BASELINE config 5 = --config vit_l14 at 2/4/8 GPUs; config 3 = --config vit_b16 --batch 1024 at 8 GPUs.
Launch:  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/fit_bench.py
         [--config vit_l14] [--batch 256] [--steps 250] [--interval 100] [--overlap K] [--json out.json]
--overlap K > 0: Cluster + Smooth on a worker thread, maps actuated K steps after sensing (fit's
slow_loop_overlap_steps); 0 = the reference's synchronous update.  Timing: CUDA events around the whole fit() call
(graph capture excluded: the step is captured before), max over ranks; slow-loop phases from the diagnostics."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="vit_l14")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=250)
ap.add_argument("--interval", type=int, default=100)
ap.add_argument("--overlap", type=int, default=0)
ap.add_argument("--json", default=None)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import __graft_entry__ as G  # noqa: E402

if local == 0:
    G.build()
if world > 1:
    dist.barrier()
import forde_b200  # noqa: E402,F401
from forde_b200 import dual_encoder, slow_loop, train  # noqa: E402

cfg = dual_encoder.CONFIGS[a.config]()
dev = torch.device("cuda", local)
st = train.create_train_state(cfg, 0, device=dev)
batches = [{k: v.to(dev) for k, v in train.synthetic_batch(cfg, a.batch, seed=42 + rank * 100 + i).items()} for i in range(2)]
g = train.GraphedTrainStep(st, batches[0])
slow_loop.warm_up()
for i in range(3):                      # steady state before the clock starts
    g(batches[i % 2])
for tower in st.variables["stats_buffer"].values():        # the interval starts clean
    for l in tower.values():
        for t in l["stateful_layer"].values():
            t.zero_()
st.step = 0


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def stream():
    for i in range(a.steps):
        yield batches[i % 2]


# fast steps alone (no slow loop): the baseline the fit() figure is compared with
barrier()
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record()
for i in range(20):
    g(batches[i % 2])
f1.record()
barrier()
ms_fast = torch.tensor([f0.elapsed_time(f1) / 20], device=dev)
if world > 1:
    dist.all_reduce(ms_fast, op=dist.ReduceOp.MAX)
ms_fast = float(ms_fast.item())
for tower in st.variables["stats_buffer"].values():
    for l in tower.values():
        for t in l["stateful_layer"].values():
            t.zero_()
st.step = 0

barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
st, hist = train.fit(st, stream(), max_steps=a.steps, slow_loop_interval=a.interval, key=1,
                     slow_loop_overlap_steps=a.overlap, step_fn=g)
e1.record()
barrier()
wall = time.perf_counter() - t0
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms.item())
if rank == 0:
    loops = hist["slow_loop"]
    slow_s = sum(d.get("seconds", {}).get("total", 0.0) for d in loops)
    first = loops[-1]["layers"].get("vision/layer_0/stateful_layer", {}) if loops else {}
    rec = {
        "config": cfg.name, "n_gpus": world, "batch_per_gpu": a.batch, "global_batch": a.batch * world, "steps": a.steps,
        "slow_loop_interval": a.interval, "slow_loop_overlap_steps": a.overlap, "slow_loops_run": len(loops),
        "ms_total_device_max_over_ranks": ms, "wall_s_rank0": wall,
        "pairs_per_s_including_slow_loops": world * a.batch * a.steps / (ms * 1e-3),
        "ms_per_step_including_slow_loops": ms / a.steps,
        "ms_per_fast_step_alone": ms_fast, "pairs_per_s_fast_steps_alone": world * a.batch / (ms_fast * 1e-3),
        "slow_loop_overhead_fraction": (ms / a.steps - ms_fast) / ms_fast,
        "slow_loop_seconds_on_training_thread": [d.get("seconds") for d in loops],
        "slow_loop_share_of_wall": slow_s / (ms * 1e-3),
        "final_loss": hist["loss"][-1][1] if hist["loss"] else None,
        "census_vision_layer_0_after_last_update": first.get("census"),
        "layers": len(st.model.layers),
    }
    print(json.dumps(rec), flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rec, f, indent=1)
sys.stdout.flush()
if world > 1:
    torch.cuda.synchronize(); dist.barrier(); os._exit(0)
