"""Training-loop contract of the fast loop (src/training/train.py:48-171), B200-side.

``create_train_state`` / ``train_step`` keep the reference's names and roles: the state carries ``params``
plus the mutable collections (``stats_buffer``, and README's ``neuron_assignments``), ``train_step`` runs
forward + backward with the statistics threaded through ``mutable=[...]``, clips the global gradient norm to
1.0 and applies AdamW (optax chain at train.py:74-77), and the slow loop is triggered by the caller every
``slow_loop_interval`` steps (train.py:365-399).

B200 layout: every master parameter is a view into ONE flat fp32 buffer (gradients likewise, plus a flat bf16
compute mirror refreshed once per step), so the data-parallel gradient all-reduce, the global-norm clip, AdamW
and the bf16 refresh are each a single pass over contiguous memory, and the NCCL all-reduce is issued in a few
large chunks on a side stream while the remaining backward still runs.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import ops
from .dual_encoder import DualEncoderConfig, FORDEDualEncoder

_ALIGN = 64  # elements; keeps every bf16 mirror view 128-byte aligned (TMA needs 16 B)


def _leaves(tree, path=()):
    # insertion order (vision: embeds, layer_0..L-1, proj; text: ...; log_temp): contiguous layer ranges in the flat
    # buffers are what the bucketed gradient all-reduce slices
    for k in tree:
        v = tree[k]
        if isinstance(v, dict):
            yield from _leaves(v, path + (k,))
        else:
            yield path + (k,), v


def _set(tree, path, value):
    for k in path[:-1]:
        tree = tree[k]
    tree[path[-1]] = value


def _get(tree, path):
    for k in path:
        tree = tree[k]
    return tree


def _mirror_tree(tree):
    return {k: (_mirror_tree(v) if isinstance(v, dict) else None) for k, v in tree.items()}


@dataclass
class TrainState:
    step: int
    model: FORDEDualEncoder
    variables: Dict[str, dict]          # params (views of flat_p32), neuron_assignments, stats_buffer
    params_bf16: Dict[str, dict]        # same tree, views of flat_p16
    flat_p32: torch.Tensor
    flat_g32: torch.Tensor
    flat_p16: torch.Tensor
    opt_state: Dict[str, torch.Tensor]  # optax.adamw state over the flat buffer: mu, nu (fp32), count (i32 scalar)
    hyper: Dict[str, float]             # learning_rate, weight_decay, b1, b2, eps
    norm_partials: torch.Tensor         # scratch of the two-pass global-norm reduction
    grad_norm: torch.Tensor             # [1] fp32, the step's pre-clip global gradient norm (train.py:168)
    flat_g16: Optional[torch.Tensor] = None   # bf16 gradient exchange buffer (world > 1, grad_comm_dtype="bf16")
    group: Optional["dist.ProcessGroup"] = None
    max_grad_norm: float = 1.0
    buckets: Optional[list] = None      # gradient all-reduce buckets (see _GradBuckets)

    @property
    def params(self):
        return self.variables["params"]

    @property
    def stats_buffer(self):
        return self.variables["stats_buffer"]


def create_train_state(config: DualEncoderConfig, key: int, learning_rate: float = 1e-4, weight_decay: float = 0.01,
                       device="cuda", group=None, grad_comm_dtype: Optional[str] = None) -> TrainState:
    """``grad_comm_dtype``: "bf16" (default, or env FORDE_GRAD_COMM) exchanges gradients between ranks as bf16 (half
    the NVLink bytes and half the time NCCL's CTAs take SMs from the backward GEMMs; each rank still accumulates its
    local gradient in fp32 and AdamW's moments stay fp32), "fp32" exchanges the fp32 buffer itself."""
    model = FORDEDualEncoder(config)
    variables = model.init(key, device)
    leaves = list(_leaves(variables["params"]))
    offs, total = [], 0
    for _, t in leaves:
        offs.append(total)
        total += (t.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
    dev = torch.device(device)
    flat_p32 = torch.zeros(total, dtype=torch.float32, device=dev)
    flat_g32 = torch.zeros(total, dtype=torch.float32, device=dev)
    flat_p16 = torch.zeros(total, dtype=torch.bfloat16, device=dev)
    p16_tree = _mirror_tree(variables["params"])
    for (path, t), off in zip(leaves, offs):
        n = t.numel()
        flat_p32[off:off + n].copy_(t.reshape(-1))
        p = flat_p32[off:off + n].view(t.shape).requires_grad_(True)
        p.grad = flat_g32[off:off + n].view(t.shape)
        p._forde_grad = p.grad          # gradient sink for the functions that accumulate in place (grad_sink)
        _set(variables["params"], path, p)
        _set(p16_tree, path, flat_p16[off:off + n].view(t.shape))
    flat_p16.copy_(flat_p32)
    flat_p32.grad = flat_g32
    # optax.adamw(lr, wd) decays every leaf (no mask, train.py:74-77), so one flat parameter is equivalent; the
    # state is optax's ScaleByAdamState (count, mu, nu) over the flat buffer, stepped by the fused kernels of
    # csrc/forde_optim.cu (apply_gradients below)
    opt_state = {"mu": torch.zeros_like(flat_p32), "nu": torch.zeros_like(flat_p32),
                 "count": torch.zeros((), dtype=torch.int32, device=dev)}
    hyper = {"learning_rate": float(learning_rate), "weight_decay": float(weight_decay), "b1": 0.9, "b2": 0.999,
             "eps": 1e-8}
    n_part = ops.optim_num_partials(total) if dev.type == "cuda" else 1
    state = TrainState(0, model, variables, p16_tree, flat_p32, flat_g32, flat_p16, opt_state, hyper,
                       torch.zeros(n_part, dtype=torch.float32, device=dev),
                       torch.zeros(1, dtype=torch.float32, device=dev), None, group)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        comm = grad_comm_dtype or os.environ.get("FORDE_GRAD_COMM", "bf16")
        if comm not in ("bf16", "fp32"):
            raise ValueError(f"grad_comm_dtype must be 'bf16' or 'fp32', got {comm!r}")
        if comm == "bf16" and dev.type == "cuda":
            state.flat_g16 = torch.zeros(total, dtype=torch.bfloat16, device=dev)
        state.buckets = _GradBuckets(leaves, offs, total, flat_g32, group, flat_g16=state.flat_g16)
        state.buckets.attach([_get(variables["params"], path) for path, _ in leaves])
    return state


def apply_gradients(state: TrainState) -> torch.Tensor:
    """optax.chain(clip_by_global_norm(max_grad_norm), adamw) on the flat buffers + bf16 mirror refresh: two streaming
    kernels (state.apply_gradients of the reference, src/training/train.py:162).  Reads the exchanged bf16 gradients
    when the data-parallel exchange ran in bf16.  Returns the pre-clip global gradient norm ([1], device)."""
    g = state.flat_g16 if state.flat_g16 is not None else state.flat_g32
    h = state.hyper
    ops.grad_sqnorm(g, state.norm_partials, state.opt_state["count"])
    ops.adamw_clip_step(state.flat_p32.detach(), g, state.opt_state["mu"], state.opt_state["nu"], state.flat_p16,
                        state.norm_partials, state.opt_state["count"], state.grad_norm, state.max_grad_norm,
                        h["learning_rate"], h["b1"], h["b2"], h["eps"], h["weight_decay"])
    return state.grad_norm


class _GradBuckets:
    """Bucketed data-parallel gradient exchange overlapped with the backward pass.

    The flat gradient buffer is cut into a few contiguous ranges (groups of transformer layers per tower).  Every
    parameter has a post-accumulate hook; when the last gradient of a range has been accumulated its SUM all-reduce
    is issued asynchronously (NCCL's own stream, ordered after the kernels that produced the gradients), so it runs
    under the rest of the backward.  `finish()` issues whatever did not fire and makes the compute stream wait."""

    def __init__(self, leaves, offs, total, flat_g32, group, layers_per_bucket: int = 4, flat_g16=None):
        # (bucket shapes were swept on 2 GPUs in round 2 -- 1, 2, 4, 12 layers per bucket, tapered tails -- with no
        # measurable effect on the step, profiles/r2_n2_bucket_variants.md: four layers per bucket stays)
        self.flat, self.group = flat_g32, group
        self.flat16 = flat_g16      # when set, a bucket is cast to bf16 on the side stream and exchanged as bf16
        ranges = []  # (key, start, end, count)
        for (path, t), off in zip(leaves, offs):
            tower = path[0]
            layer = next((k for k in path if k.startswith("layer_")), None)
            if layer is not None:
                idx = int(layer.split("_")[1]) // layers_per_bucket
            elif len(path) > 1 and ("embed" in path[1] or path[1] == "cls"):
                idx = -1            # front end of the tower: its gradients arrive last
            else:
                idx = 10 ** 6       # projection head / log-temperature: arrive first
            key = (tower, idx)
            end = off + (t.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            if ranges and ranges[-1][0] == key:
                ranges[-1][2], ranges[-1][3] = end, ranges[-1][3] + 1
            else:
                ranges.append([key, off, end, 1])
        self.ranges = ranges
        self.pending = [r[3] for r in ranges]
        self.works = []
        self.fired = [False] * len(ranges)
        self.enabled = False
        self.comm_stream = None
        self.epoch = 0
        self.mode = os.environ.get("FORDE_BUCKET_MODE", "stream")
        # FORDE_GRAD_OVERLAP=0: exchange every bucket after the backward instead of under it (A/B of the overlap)
        self.overlap = os.environ.get("FORDE_GRAD_OVERLAP", "1") != "0"

    def attach(self, params_in_order):
        """params_in_order: the leaf tensors in the same order as `leaves`.  The hook fires from autograd's
        AccumulateGrad, or is called directly (``p._forde_ready``) by the functions that write into the gradient views
        themselves (stateful_layer.grad_sink)."""
        i = 0
        for bi, r in enumerate(self.ranges):
            for _ in range(r[3]):
                hook = self._make_hook(bi)
                params_in_order[i].register_post_accumulate_grad_hook(hook)
                params_in_order[i]._forde_ready = hook
                i += 1

    def _make_hook(self, bi):
        def hook(param):
            if not self.enabled:
                return
            # idempotent per step: a parameter whose gradient went through a sink is announced by the function that
            # wrote it AND (in current PyTorch) by AccumulateGrad's hook, which also runs for a None gradient
            if getattr(param, "_forde_mark", -1) == self.epoch:
                return
            param._forde_mark = self.epoch
            self.pending[bi] -= 1
            if self.pending[bi] == 0 and self.overlap:
                self._fire(bi)
        return hook

    def _fire(self, bi):
        _, a, b, _ = self.ranges[bi]
        if self.flat.is_cuda and self.mode == "stream":
            # explicit side stream: it waits for everything enqueued so far on the compute stream (the kernels that
            # produced this bucket's gradients), the collective is ordered on it, finish() joins it back
            cur = torch.cuda.current_stream(self.flat.device)
            if self.comm_stream is None:
                self.comm_stream = torch.cuda.Stream(device=self.flat.device)
            self.comm_stream.wait_stream(cur)
            with torch.cuda.stream(self.comm_stream):
                if self.flat16 is not None:
                    ops.cast_f32_bf16(self.flat[a:b], self.flat16[a:b])
                    dist.all_reduce(self.flat16[a:b], op=dist.ReduceOp.SUM, group=self.group)
                else:
                    dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group)
        else:
            buf = self.flat
            if self.flat16 is not None:
                ops.cast_f32_bf16(self.flat[a:b], self.flat16[a:b])
                buf = self.flat16
            self.works.append(dist.all_reduce(buf[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.fired[bi] = True

    def begin(self):
        self.enabled = True
        self.epoch += 1
        self.pending = [r[3] for r in self.ranges]
        self.fired = [False] * len(self.ranges)
        self.works = []

    def finish(self):
        self.enabled = False
        for bi in range(len(self.ranges)):
            if not self.fired[bi]:      # a range whose parameters did not all receive gradients this step
                self._fire(bi)
        for w in self.works:
            w.wait()
        self.works = []
        if self.comm_stream is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)


def sync_gradients(flat_g32: torch.Tensor, group=None) -> None:
    """Data-parallel gradient exchange.  The loss every rank differentiates is the GLOBAL-batch loss and each
    rank holds the partial derivative through its own samples, so the full gradient is the SUM over ranks
    (not the mean).  One all-reduce over the flat buffer; no-op on a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_g32, op=dist.ReduceOp.SUM, group=group)


def _to_device(t: torch.Tensor, dev) -> torch.Tensor:
    return t if t.device == dev else t.to(dev, non_blocking=True)


def train_step(state: TrainState, batch: Dict[str, torch.Tensor], sense: bool = True):
    """One fast-loop step.  batch = {"image": (B,S,S,3), "input_ids": (B,T)}; host (pinned) or device tensors.
    Returns (state, metrics) with metrics["loss"], metrics["grad_norm"] as device scalars (train.py:167-171)."""
    dev = state.flat_p32.device
    images = _to_device(batch["image"], dev)
    ids = _to_device(batch["input_ids"], dev)
    state.flat_g32.zero_()
    loss, _ = state.model.apply(state.variables, images, ids, mutable=["stats_buffer"] if sense else [],
                                params_bf16=state.params_bf16, group=state.group)
    if state.buckets is not None:
        state.buckets.begin()
        loss.backward()
        if dev.type == "cuda":
            state.model.join_streams(dev)   # the text tower's backward ran on its own stream
        state.buckets.finish()          # bucketed SUM all-reduce, overlapped with the backward
    else:
        loss.backward()
        if dev.type == "cuda":
            state.model.join_streams(dev)
        sync_gradients(state.flat_g32, state.group)
    gnorm = apply_gradients(state)      # clip_by_global_norm(1.0) -> AdamW -> bf16 mirror, two fused passes
    state.step += 1
    return state, {"loss": loss.detach(), "grad_norm": gnorm.reshape(())}


class GraphedTrainStep:
    """The whole fast-loop step (forward, loss, backward, gradient all-reduce, clip, AdamW, bf16 refresh, in-kernel
    sensing and stats accumulation) captured ONCE into a CUDA graph and replayed per step: the step is ~1600
    kernel launches, and launching them from Python costs more than running them on a B200.

    Usage: ``g = GraphedTrainStep(state, example_batch); state, metrics = g(batch)``.  ``batch`` may live on the
    host (pinned) or the device; it is copied into the graph's static input buffers on the current stream.
    ``metrics`` are the graph's static output scalars (valid until the next call).  The slow loop may rewrite
    ``neuron_assignments`` in place and zero ``stats_buffer`` between replays; both are read/updated by kernels
    inside the graph."""

    def __init__(self, state: TrainState, example_batch: Dict[str, torch.Tensor], sense: bool = True, warmup: int = 3):
        dev = state.flat_p32.device
        self.state = state
        self.static = {k: torch.empty_like(v, device=dev) for k, v in example_batch.items()}
        for k, v in example_batch.items():
            self.static[k].copy_(v, non_blocking=True)
        # The warm-up runs are real train steps (they size workspaces, prime cuBLAS/cuDNN/NCCL and the TMA descriptor
        # cache).  Everything they mutate is snapshotted here and restored IN PLACE after the capture (addresses stay
        # valid for the graph), so building the graph trains on nothing: one optimizer update per replayed batch, as
        # in the reference's loop (src/training/train.py:320-343).
        snap = _snapshot_mutable(state)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                train_step(state, self.static, sense)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._staging, self._copy_stream, self._staged_for, self._d2d_done = None, None, None, None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            _, self.metrics = train_step(state, self.static, sense)
        torch.cuda.synchronize(dev)
        _restore_mutable(state, snap)
        torch.cuda.synchronize(dev)

    def __call__(self, batch: Dict[str, torch.Tensor], prefetch: Optional[Dict[str, torch.Tensor]] = None):
        """Run one step on ``batch``.  ``prefetch`` (optional) is the NEXT step's host batch: its host-to-device copy
        is started on a side stream into a staging buffer so it overlaps this step's compute; pass that same batch
        object as ``batch`` next time and only a device-to-device move remains on the critical path."""
        dev = self.state.flat_p32.device
        cur = torch.cuda.current_stream(dev)
        for k, v in batch.items():
            if tuple(v.shape) != tuple(self.static[k].shape) or v.dtype != self.static[k].dtype:
                raise ValueError(f"GraphedTrainStep was captured for {k} of shape {tuple(self.static[k].shape)} "
                                 f"{self.static[k].dtype}; got {tuple(v.shape)} {v.dtype} (capture a new one)")
        if self._staged_for is not None and self._staged_for is batch:
            cur.wait_stream(self._copy_stream)
            for k in self.static:
                self.static[k].copy_(self._staging[k], non_blocking=True)
            self._d2d_done = torch.cuda.Event()
            self._d2d_done.record(cur)          # the staging buffers may be refilled once this has passed
        else:
            for k, v in batch.items():
                if v.data_ptr() != self.static[k].data_ptr():
                    self.static[k].copy_(v, non_blocking=True)
        self._staged_for = None
        self.graph.replay()
        if prefetch is not None:
            if self._staging is None:
                self._staging = {k: torch.empty_like(v) for k, v in self.static.items()}
                self._copy_stream = torch.cuda.Stream(device=dev)
            if self._d2d_done is not None:
                self._copy_stream.wait_event(self._d2d_done)   # NOT wait_stream(cur): that would wait for this step
            with torch.cuda.stream(self._copy_stream):
                for k, v in prefetch.items():
                    self._staging[k].copy_(v, non_blocking=True)
            self._staged_for = prefetch
        self.state.step += 1
        return self.state, self.metrics


def _stat_leaves(tree):
    for v in tree.values():
        if isinstance(v, dict):
            yield from _stat_leaves(v)
        else:
            yield v


def _snapshot_mutable(state: TrainState):
    """Clones of everything a train step mutates: parameters, optimizer state, stats_buffer leaves, step counter."""
    return {"p32": state.flat_p32.detach().clone(), "opt": {k: v.clone() for k, v in state.opt_state.items()},
            "stats": [t.clone() for t in _stat_leaves(state.variables["stats_buffer"])], "step": state.step}


def _restore_mutable(state: TrainState, snap) -> None:
    state.flat_p32.detach().copy_(snap["p32"])
    state.flat_p16.copy_(state.flat_p32)
    for k, v in state.opt_state.items():
        v.copy_(snap["opt"][k])
    for t, s in zip(_stat_leaves(state.variables["stats_buffer"]), snap["stats"]):
        t.copy_(s)
    state.step = snap["step"]


def save_checkpoint(state: TrainState, path: str) -> None:
    """Save params, neuron_assignments, the stat accumulators, the optimizer moments and the step counter (the
    reference saves the whole TrainState once at the end, src/training/train.py:412-422, and has no restore).
    Only tensors, numbers and strings are written, so the file loads with ``weights_only=True``."""
    import dataclasses
    cfg = state.model.cfg
    cfg_dict = {k: (dataclasses.asdict(v) if dataclasses.is_dataclass(v) else v)
                for k, v in (dataclasses.asdict(cfg) if dataclasses.is_dataclass(cfg) else dict(vars(cfg))).items()}
    torch.save({"step": state.step, "config": cfg_dict, "flat_p32": state.flat_p32.detach().cpu(),
                "opt_state": {k: v.detach().cpu() for k, v in state.opt_state.items()}, "hyper": dict(state.hyper),
                "neuron_assignments": _tree_to(state.variables["neuron_assignments"], "cpu"),
                "stats_buffer": _tree_to(state.variables["stats_buffer"], "cpu")}, path)


def restore_checkpoint(state: TrainState, path: str) -> TrainState:
    """Load a checkpoint written by save_checkpoint into an already created TrainState of the same config, in place:
    parameter views, bf16 mirror, brain maps, accumulators AND the optimizer moments keep their addresses, so captured
    CUDA graphs and cached TMA descriptors stay valid."""
    ck = torch.load(path, map_location="cpu", weights_only=True)
    if ck["flat_p32"].numel() != state.flat_p32.numel():
        raise ValueError("checkpoint does not match this model configuration")
    state.flat_p32.detach().copy_(ck["flat_p32"])
    state.flat_p16.copy_(state.flat_p32)
    for k, v in state.opt_state.items():          # optimizer moments and count, in place
        v.copy_(ck["opt_state"][k])
    state.hyper.update(ck.get("hyper", {}))
    _tree_copy_(state.variables["neuron_assignments"], ck["neuron_assignments"])
    _tree_copy_(state.variables["stats_buffer"], ck["stats_buffer"])
    state.step = int(ck["step"])
    return state


def _tree_to(tree, device):
    return {k: (_tree_to(v, device) if isinstance(v, dict) else v.detach().to(device)) for k, v in tree.items()}


def _tree_copy_(dst, src):
    for k, v in dst.items():
        if isinstance(v, dict):
            _tree_copy_(v, src[k])
        else:
            v.copy_(src[k])


def fit(state: TrainState, batches, max_steps: int, slow_loop_interval: int = 100, slow_loop_config=None, key: int = 0,
        use_cuda_graph: bool = True, log=None, plot_dir: Optional[str] = None, slow_loop_overlap_steps: int = 0,
        step_fn=None):
    """The reference's outer loop (src/training/train.py:320-399): fast steps, and every ``slow_loop_interval`` steps
    (train.py:219-224 default 100; ``step > 0 and step % interval == 0``, train.py:365-369) the slow loop -- Sense
    (all-reduced pooled moments) -> Cluster -> Smooth -> Actuate -> Reset -- which rewrites ``neuron_assignments`` in
    place and zeroes ``stats_buffer`` (train.py:387-399).  ``batches`` is an iterable of batch dicts.
    With ``plot_dir`` rank 0 also writes each layer's brain scan and feature-space scatter after every slow loop
    (the reference's plot_brain_scan / plot_feature_space, src/utils/logging.py:7-52), into one sub-directory per layer.
    ``slow_loop_overlap_steps`` = k > 0 takes the slow loop's host work (GMM + smoothing, ~0.2 s per update at
    ViT-B/16 sizes = 5 fast steps) off the training thread: statistics are sensed and reset at the slow-loop step,
    clustered on a worker thread while fast steps continue, and the new maps take effect exactly k steps later on
    every rank (slow_loop.OverlappedSlowLoop).  0 = the reference's synchronous update.
    ``step_fn``: an already captured ``GraphedTrainStep`` of this state to reuse (otherwise one is captured from the
    first batch).  Every slow-loop diagnostics dict carries ``seconds``: host time the training thread spent in it.
    Returns (state, history) with history = list of (step, loss) and the slow-loop diagnostics."""
    from .brain_scan import plot_brain_scan, plot_feature_space
    from .slow_loop import OverlappedSlowLoop, SlowLoopConfig, slow_loop_step

    cfg = slow_loop_config or SlowLoopConfig()
    warm = None
    if max_steps > slow_loop_interval:   # a slow loop will run: pay its imports now, under the first fast steps
        import threading
        from .slow_loop import warm_up
        warm = threading.Thread(target=warm_up, daemon=True)
        warm.start()
    history, diagnostics = [], []
    overlapped = OverlappedSlowLoop(cfg, slow_loop_overlap_steps, state.group) if slow_loop_overlap_steps > 0 else None

    def record(diag):
        if diag is None:
            return
        diagnostics.append(diag)
        if plot_dir is not None and (not dist.is_initialized() or dist.get_rank(state.group) == 0):
            for name, d in diag["layers"].items():
                if "brain_scan" not in d:
                    continue
                sub = os.path.join(plot_dir, name.replace("/", "_"))
                os.makedirs(sub, exist_ok=True)
                plot_brain_scan(d["brain_scan"], diag["step"], 0, out_dir=sub)
                plot_feature_space(d["features"], d["brain_scan"].reshape(-1), diag["step"], 0, out_dir=sub)

    it = iter(batches)
    for step in range(max_steps):
        try:
            batch = next(it)
        except StopIteration:
            break
        if overlapped is not None:
            record(overlapped.poll(step))      # a background update that is due takes effect before this step
        if use_cuda_graph and state.flat_p32.is_cuda:
            if step_fn is None:
                step_fn = GraphedTrainStep(state, batch)
            _, m = step_fn(batch)
        else:
            _, m = train_step(state, batch)
        if log is not None or step == max_steps - 1 or (step + 1) % slow_loop_interval == 0:
            history.append((step, float(m["loss"].item())))
            if log is not None:
                log(step, history[-1][1])
        if step > 0 and step % slow_loop_interval == 0:
            if warm is not None:
                warm.join()
                warm = None
            mv = {"stats_buffer": state.variables["stats_buffer"],
                  "neuron_assignments": state.variables["neuron_assignments"]}
            if overlapped is not None:
                record(overlapped.start(mv, key=key + step, epoch=0, step=step))
            else:
                _, _, diag = slow_loop_step(state.params, mv, cfg, key=key + step, epoch=0, step=step, group=state.group)
                record(diag)
    if overlapped is not None:
        record(overlapped.flush())
        overlapped.close()
    return state, {"loss": history, "slow_loop": diagnostics}


def synthetic_batch(config: DualEncoderConfig, batch: int, seed: int = 42, pinned: bool = True):
    """Synthetic image/caption pairs of the reference's tensor layout (src/data/download_to_drive.py:106-108,
    171-178): image (B,S,S,3) NHWC, input_ids (B,T) int32; fixed-seed ints as src/data/dataset.py:27,39."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randn((batch, config.image_size, config.image_size, 3), generator=g).to(torch.bfloat16)
    ids = torch.randint(0, config.vocab_size, (batch, config.text_len), generator=g, dtype=torch.int32)
    if pinned and torch.cuda.is_available():
        img, ids = img.pin_memory(), ids.pin_memory()
    return {"image": img, "input_ids": ids}
