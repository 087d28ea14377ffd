#!/usr/bin/env python
"""bench.py -- FORDE dual-encoder train-step throughput (image-text pairs/s) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config vit_b16] [--batch 256] [--impl reference]

One "step" = one fast-loop train step of the FORDE dual encoder (README.md:26-56) on one synthetic batch of
image/caption pairs: forward of both towers (every FFN slot is the sm_100a StatefulLayer with in-kernel
sensing), contrastive loss over the global batch, backward, gradient all-reduce (N > 1), global-norm clip,
AdamW, bf16 weight refresh.  Batch-sharded ("weak" scaling: --batch pairs per GPU).

JSON line keys follow the driver's contract: `value` = pairs/s with inputs resident in HBM, `e2e` = the same
through train_step() with HOST (pinned) batches, the H2D copy and the loss read-back inside the timed region;
`roofline` = the dominant StatefulLayer kernel (tensor bound) timed live with CUDA events on the launching
stream; `cpu_baseline` = the CPU port of the same step (oracle/dual_encoder_cpu.py) on the host cores.

--impl reference times that CPU port alone (the reference's JAX code cannot be installed here and contains no
dual-encoder code) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "image_text_pairs_per_sec"
UNIT = "pairs/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"burst": p["bf16_tflops"], "sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm": p["hbm_gbs"], "source": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


def workload_config(cfg, batch, world):
    """`config` of the JSON line: the workload only, identical for the B200 arm and the reference arm."""
    return {"workload": f"FORDE {cfg.name} dual-encoder train step: ViT {cfg.image_size}px/p{cfg.patch_size} "
                        f"d={cfg.vision.width} H={cfg.vision.hidden} L={cfg.vision.layers} + text S={cfg.text_len} "
                        f"d={cfg.text.width} H={cfg.text.hidden} L={cfg.text.layers}, embed {cfg.embed_dim}; forward, "
                        "contrastive loss, backward, in-kernel sensing, clip_by_global_norm(1.0) + AdamW",
            "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": f"dp{world}",
            "optimizer": "clip_by_global_norm(1.0)+AdamW fp32 master, bf16 compute",
            "l2": "per-step working set (activations + weights, several GB) exceeds the 126 MB L2; no flush"}


def lib_sha16():
    """First 16 hex digits of the SHA-256 of the shared library the run loads (evidence is tied to a build)."""
    import hashlib
    path = os.environ.get("FORDE_LIB", os.path.join(ROOT, "forde-vision-language_b200", "libforde_b200.so"))
    try:
        with open(path, "rb") as f:
            return hashlib.sha256(f.read()).hexdigest()[:16]
    except OSError:
        return None


def lib_src16():
    """First 16 hex digits of the hash of the SOURCES the loaded library was built from (__graft_entry__.build() records
    it next to the library); unlike the binary's hash it does not depend on the directory the build ran in."""
    path = os.environ.get("FORDE_LIB", os.path.join(ROOT, "forde-vision-language_b200", "libforde_b200.so")) + ".srchash"
    try:
        with open(path) as f:
            return f.read().strip()[:16] or None
    except OSError:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = time.mktime(time.strptime(f[0].split(".")[0], "%Y/%m/%d %H:%M:%S")) + float("0." + f[0].split(".")[1])
                if self.t0 is not None and not (self.t0 <= ts <= (self.t1 or ts)):
                    continue          # keep only samples taken DURING the timed region
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """CPU port of the same train step on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import forde_b200  # noqa: F401
    from forde_b200 import dual_encoder, train
    from oracle import dual_encoder_cpu as C

    cfg = dual_encoder.CONFIGS[args.config]()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = args.cpu_batch
    st = train.create_train_state(cfg, 0, device="cpu")
    tr = C.CpuTrainer(cfg, st.variables)
    b = train.synthetic_batch(cfg, sample, pinned=False)
    img = b["image"].float()
    for _ in range(args.warmup):
        tr.step(img, b["input_ids"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tr.step(img, b["input_ids"])
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args.batch, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step = {sample} synthetic pairs of the {cfg.name} workload (a bounded sample "
                                   f"of the {args.batch} pairs per GPU per step the B200 arm runs); "
                                   f"{1e3 * dt / args.steps / sample:.1f} ms per pair; torch CPU fp32 port of the "
                                   "restated reference on all host cores (JAX/XLA not installable here)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as G
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the FORDE hot path has no CPU fallback "
                         "(use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if local == 0:
        G.build()
    if world > 1:
        dist.barrier()
    import forde_b200  # noqa: F401
    from forde_b200 import dual_encoder, ops, train

    cfg = dual_encoder.CONFIGS[args.config]()
    dev = torch.device("cuda", local)
    state = train.create_train_state(cfg, 0, device=dev)
    host = [train.synthetic_batch(cfg, args.batch, seed=42 + rank * 100 + i, pinned=True) for i in range(2)]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batches, steps, read_loss, step_fn=None, prefetch=False):
        step_fn = step_fn or (lambda b: train.train_step(state, b))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for i in range(steps):
            if prefetch:   # start the NEXT batch's H2D copy on a side stream while this step computes
                _, m = step_fn(batches[i % len(batches)], prefetch=batches[(i + 1) % len(batches)])
            else:
                _, m = step_fn(batches[i % len(batches)])
            if read_loss:
                last = float(m["loss"].item())      # D2H of the step's result inside the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), last

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()       # started early (nvidia-smi takes a while to come up); samples are filtered by time
    for i in range(max(args.warmup, 3)):
        train.train_step(state, devb[i % 2])
    torch.cuda.synchronize()

    # ---- per-kernel pass (eager launches, CUDA events around every StatefulLayer kernel on the launching stream):
    #      the same K steps as the timed region below, which replays them as one CUDA graph and so cannot be
    #      instrumented per kernel.  Launch counts come from this pass too (a replay launches the same kernels).
    #      Kernels are timed one at a time: both towers on ONE stream here (the measured step overlaps them).
    concurrent = state.model.concurrent_towers
    state.model.concurrent_towers = False
    ops.enable_timing(True)
    l0 = ops.launch_count()
    ms_eager, _ = timed(devb, args.steps, read_loss=False)
    launches = ops.launch_count() - l0
    kt = ops.timing_summary()
    ops.enable_timing(False)
    state.model.concurrent_towers = concurrent

    step_fn, graphed = None, False
    if not args.no_graph:
        try:
            step_fn = train.GraphedTrainStep(state, devb[0])
            graphed = True
        except Exception as e:  # noqa: BLE001  (report and fall back to eager launches)
            print(f"[bench] CUDA graph capture failed, timing eager launches: {type(e).__name__}: {e}", file=sys.stderr)
            step_fn = None

    # ---- device-resident throughput (the metric), with clock sampling
    for i in range(2):
        (step_fn or (lambda b: train.train_step(state, b)))(devb[i % 2])
    sampler.mark_begin()
    ms, _ = timed(devb, args.steps, read_loss=False, step_fn=step_fn)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    value = world * args.batch * args.steps / (ms * 1e-3)

    # ---- end to end: host (pinned) batches -> public train-step API -> loss on the host, every step
    for i in range(2):
        (step_fn or (lambda b: train.train_step(state, b)))(host[i % 2])
    ms_e2e, loss = timed(host, args.steps, read_loss=True, step_fn=step_fn, prefetch=graphed)
    e2e = world * args.batch * args.steps / (ms_e2e * 1e-3)

    # ---- the same step on a post-slow-loop brain map: contiguous thirds of relu / tanh / step neurons in every layer
    #      (spatially smoothed maps are contiguous, smoothing.py:30-91).  tanh groups make the forward store z^T and
    #      the backward read it; the all-generalist map above is the cheapest case.  The graph sees the in-place rewrite.
    smoothed = None
    if not args.no_smoothed_map:
        for tower in state.variables["neuron_assignments"].values():
            for l in tower.values():
                a_ = l["stateful_layer"]["assignments"]
                a_.copy_((torch.arange(a_.numel(), device=dev) * 3 // a_.numel()).to(a_.dtype))
        for i in range(2):
            (step_fn or (lambda b: train.train_step(state, b)))(devb[i % 2])
        ms_s, _ = timed(devb, args.steps, read_loss=False, step_fn=step_fn)
        state.model.concurrent_towers = False
        ops.enable_timing(True)
        timed(devb, min(args.steps, 3), read_loss=False)
        kt_s = ops.timing_summary()
        ops.enable_timing(False)
        state.model.concurrent_towers = concurrent
        smoothed = {"map": "contiguous thirds of relu / tanh / binary_step neurons in every layer",
                    "value": world * args.batch * args.steps / (ms_s * 1e-3), "unit": UNIT,
                    "ms_per_step": ms_s / args.steps,
                    "kernels_us": {n: round(1e3 * r["ms_total"] / r["launches"], 2) for n, r in kt_s.items()
                                   if n.startswith(("sl_up_fwd", "sl_dz_bwd"))}}

    # ---- BASELINE config 3 (global batch 8192 = 1024 pairs per GPU on 8 GPUs): carried on the 8-GPU line
    c3 = None
    if world == 8 and args.batch != 1024 and not args.no_c3 and args.config == "vit_b16":
        try:
            del step_fn, state, devb, host
            torch.cuda.empty_cache()
            st3 = train.create_train_state(cfg, 0, device=dev)
            b3 = [{k: v.to(dev) for k, v in train.synthetic_batch(cfg, 1024, seed=7 + rank * 100 + i, pinned=False).items()}
                  for i in range(2)]
            g3 = train.GraphedTrainStep(st3, b3[0], warmup=2)
            for i in range(2):
                g3(b3[i % 2])
            n3 = 5
            ms3, _ = timed(b3, n3, read_loss=False, step_fn=g3)
            _, m3 = g3(b3[0])
            c3 = {"global_batch": 1024 * world, "batch_per_gpu": 1024, "steps": n3, "ms_per_step": ms3 / n3,
                  "value": world * 1024 * n3 / (ms3 * 1e-3), "unit": UNIT, "loss": float(m3["loss"].item()),
                  "local_logits": "1024 x 8192 fp32 per rank; text embeddings all-gathered (8.4 MB bf16), column "
                                  "log-sum-exp merged by one MAX + one SUM all-reduce, d(txt) by reduce-scatter",
                  "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1)}
        except Exception as e:  # noqa: BLE001  (the headline line must still be printed)
            c3 = {"error": f"{type(e).__name__}: {e}"[:300]}

    def finish():
        # NCCL communicators that were captured into a CUDA graph can hang in their destructor at interpreter
        # exit; everything is measured and printed by now, so synchronise, meet at a barrier and leave.
        sys.stdout.flush(); sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        finish()
        return
    peaks = measured_peaks()
    kernels = {}
    for name, r in kt.items():
        us = 1e3 * r["ms_total"] / r["launches"]
        kernels[name] = {"launches": r["launches"], "us_per_launch": round(us, 2),
                         "tflops": round(r["flops_total"] / r["launches"] / (us * 1e-6) / 1e12, 1)}
    roofline = None
    sl_ms = sum(r["ms_total"] for r in kt.values())
    sl_flops = sum(r["flops_total"] for r in kt.values())
    traffic_db, traffic_note = {}, "no ncu capture on record for this build"
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    sha, src = lib_sha16(), lib_src16()
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        # only quote DRAM bytes captured on THIS build of the library: the same binary, or one built from the same sources
        same_bin = tj.get("lib_sha256_16") is not None and tj.get("lib_sha256_16") == sha
        same_src = tj.get("src_sha256_16") is not None and tj.get("src_sha256_16") == src
        # ... or one whose GEMM engine (the only kernels in the table) is built from the same files and flags, when the
        # loaded library is the one build() made from this tree
        import __graft_entry__ as G
        same_gemm = (tj.get("gemm_src_sha256_16") is not None and src == G.source_hash()[:16]
                     and tj.get("gemm_src_sha256_16") == G.source_hash(G.GEMM_SOURCES)[:16])
        if same_bin or same_src or same_gemm:
            traffic_db = tj.get("kernels", {})
            how = "same binary" if same_bin else "same sources" if same_src else "same GEMM-engine sources"
            traffic_note = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch, one ncu --set full capture of "
                            f"library {tj.get('lib_sha256_16')} built from sources {tj.get('src_sha256_16')} "
                            f"({how}; {tj.get('source', 'profiles/')})")
        else:
            traffic_note = (f"profiles/ncu_traffic.json was captured on library {tj.get('lib_sha256_16')} (sources "
                            f"{tj.get('src_sha256_16')}), this run loads {sha} (sources {src}): not quoted")
    if kt:
        dom = max(kt, key=lambda n: kt[n]["ms_total"])
        r = kt[dom]
        ach = r["flops_total"] / (r["ms_total"] * 1e-3) / 1e12
        roofline = {"kernel": dom, "bound": "tensor", "achieved": round(ach, 1), "peak": peaks["sustained"],
                    "unit": "TFLOP/s", "frac": round(ach / peaks["sustained"], 4),
                    "peak_kind": f"{peaks['source']} sustained bf16 (kernel timed inside a long step)",
                    "flops_per_launch": r["flops_total"] / r["launches"], "launches": r["launches"],
                    "us_per_launch": round(1e3 * r["ms_total"] / r["launches"], 2),
                    "traffic": traffic_db.get(dom, {}).get("traffic_bytes"),
                    "algorithmic_bytes": traffic_db.get(dom, {}).get("algorithmic_bytes"),
                    "traffic_note": traffic_note}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": workload_config(cfg, args.batch, world),
        "run": {"launch": "whole step replayed as one CUDA graph" if graphed else "eager launches",
                "ms_per_step_eager_instrumented": ms_eager / args.steps, "lib_sha256_16": sha, "src_sha256_16": src,
                "towers": ("text tower on its own CUDA stream, forward and backward (fork/join inside the graph); the "
                           "per-kernel pass runs both towers on one stream") if concurrent else "one stream",
                "brain_map": "all-generalist (every layer's initial map, D7); see smoothed_map for a post-slow-loop map",
                "grad_exchange": (os.environ.get("FORDE_GRAD_COMM", "bf16") if world > 1 else "none (1 GPU)")},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "loss": loss},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "stateful_layer": {"flops_per_step": dual_encoder.stateful_layer_flops(cfg, args.batch),
                           "ms_per_step_in_kernels": sl_ms / args.steps,
                           "tflops": round(sl_flops / (sl_ms * 1e-3) / 1e12, 1) if sl_ms else None,
                           "frac_of_sustained_peak": round(sl_flops / (sl_ms * 1e-3) / 1e12 / peaks["sustained"], 4) if sl_ms else None,
                           "kernels": kernels},
    }
    if smoothed is not None:
        line["smoothed_map"] = smoothed
    if c3 is not None:
        line["c3_global_batch_8192"] = c3
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, args)
    print(json.dumps(line), flush=True)
    finish()


def cpu_baseline(cfg, args):
    """Bounded sample of the same workload on the host cores (about 10-30 s)."""
    import torch
    from forde_b200 import train
    from oracle import dual_encoder_cpu as C

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = args.cpu_batch
    st = train.create_train_state(cfg, 0, device="cpu")
    tr = C.CpuTrainer(cfg, st.variables)
    b = train.synthetic_batch(cfg, sample, pinned=False)
    img = b["image"].float()
    tr.step(img, b["input_ids"])
    t0 = time.perf_counter()
    n = 0
    while n < 2 or (time.perf_counter() - t0 < 12.0 and n < 50):
        tr.step(img, b["input_ids"])
        n += 1
    dt = time.perf_counter() - t0
    return {"value": sample * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} steps of {sample} synthetic pairs of the {cfg.name} workload (fwd+bwd+sensing pass+AdamW), "
                      f"{1e3 * dt / n / sample:.1f} ms per pair; torch CPU fp32 port of the restated reference on all "
                      "host cores (JAX not installable here)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="vit_b16", choices=["tiny", "vit_b16", "vit_l14"])
    ap.add_argument("--batch", type=int, default=256, help="pairs per GPU per step")
    ap.add_argument("--cpu-batch", type=int, default=32, help="pairs per step of the CPU port's bounded sample")
    ap.add_argument("--no-smoothed-map", action="store_true", help="skip the second measurement on a mixed brain map")
    ap.add_argument("--no-c3", action="store_true", help="on 8 GPUs: skip the extra global-batch-8192 measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of the CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
