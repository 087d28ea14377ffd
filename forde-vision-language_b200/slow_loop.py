"""Slow loop: Sense -> Cluster -> Smooth -> Actuate -> Reset, every N fast steps (README.md:17-20).

Host-side, exactly like the reference (src/forde/moe_slow_loop.py:253-464 is "non-jit-able" Python around
scikit-learn and a convolution); it touches (H,5) floats per layer once every ~100 steps, so it is not
accelerated (out of scope).  What is kept is the INTERFACE:

  * input / output are plain nested dicts of array leaves; layers are found under ``layer_{i}`` keys and the
    statistics leaves by name (moe_slow_loop.py:145-170); the call returns
    ``(params, mutable_variables, diagnostics)`` (moe_slow_loop.py:455-464);
  * Sense consumes the accumulators the fast loop filled (``stats_buffer``), Cluster is
    ``GaussianMixture(n_components, random_state=seed)`` (clustering.py:27-30), Smooth is the per-cluster box
    filter + argmax on a 2-D neuron grid (smoothing.py:30-91), Actuate rewrites ``neuron_assignments``
    (README.md:20), and every leaf of ``stats_buffer`` is zeroed afterwards (moe_slow_loop.py:441-449,
    tests/verify_reset.py:61-67).

Multi-GPU: the raw-moment accumulators are additive, so ranks all-reduce(SUM) ``moments`` and ``sample_count``
(never the finished statistics, Hoyer and variance are non-linear) -- ONE fp64 all-reduce for all layers -- and every
rank then runs the same seeded slow loop.  Replicas are kept identical by construction AND by a broadcast of rank 0's
new maps before Actuate (``SlowLoopConfig.broadcast``): bit-identical scikit-learn fits on every rank are likely but
not something to stake the replicas' consistency on.

Deviations from the reference's slow loop, all switchable: ``canonicalize`` (D10) relabels
the GMM's arbitrary cluster ids by ascending mean grad_gini so that id 0/1/2 keep meaning relu/tanh/step from one
update to the next (``cluster_neurons_gmm`` returns raw ids); ``pooled`` senses from raw moments pooled over the
interval and the ranks instead of ``neuron_stats / step_count`` (moe_slow_loop.py:298).  ``canonicalize=False,
pooled=False`` is the reference's behaviour.
"""
from __future__ import annotations

from dataclasses import dataclass
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import ops


@dataclass
class SlowLoopConfig:
    num_clusters: int = 3            # README.md:52-54: three processing paths
    kernel_size: int = 3             # smoothing.py:31
    pooled: bool = True              # Sense from pooled raw moments (exact across ranks); False = mean of
                                     # per-step statistics (neuron_stats / step_count, moe_slow_loop.py:298)
    canonicalize: bool = True        # D10: relabel clusters by ascending mean grad_gini -> 0 relu, 1 tanh, 2 step
    forde_lite: bool = False         # README.md:58-59 ablation: is_spec = grad_gini > threshold, no GMM
    forde_lite_threshold: float = 0.8
    broadcast: bool = True           # multi-rank: rank 0's new maps are broadcast before Actuate (one int32 tensor)


def default_grid(H: int) -> Tuple[int, int]:
    """Most square factorisation of H (3072 -> 48x64, 2048 -> 32x64, 4096 -> 64x64, 512 -> 16x32)."""
    best = (1, H)
    for a in range(1, int(H ** 0.5) + 1):
        if H % a == 0:
            best = (a, H // a)
    return best


def box_smooth(grid: np.ndarray, kernel_size: int, num_clusters: int) -> np.ndarray:
    """Majority-style smoothing of an integer label grid (smoothing.py:30-91): per-label density by a
    zero-padded k x k box filter, then the label of highest density; ties resolve to the smallest label."""
    from scipy.signal import convolve2d

    h, w = grid.shape
    box = np.full((kernel_size, kernel_size), 1.0 / kernel_size ** 2, dtype=np.float32)
    # the reference pads tiny grids so each side exceeds the kernel (smoothing.py:59-68) and crops back
    ph, pw = max(0, kernel_size + 1 - h), max(0, kernel_size + 1 - w)
    top, left = ph // 2, pw // 2
    dens = np.empty((h, w, num_clusters), dtype=np.float32)
    for c in range(num_clusters):
        plane = np.pad((grid == c).astype(np.float32), ((top, ph - top), (left, pw - left)))
        dens[:, :, c] = convolve2d(plane, box, mode="same")[top:top + h, left:left + w]
    return dens.argmax(axis=-1).astype(np.int32)


def warm_up() -> None:
    """Import scikit-learn / SciPy and run one tiny fit so the first slow loop does not pay ~4 s of imports.
    ``train.fit`` calls this on a background thread while the first fast steps run."""
    rng = np.random.RandomState(0)
    # a realistically sized job through the same code path as an update (thread pool, threadpoolctl, sklearn's
    # lazily imported linear algebra): the first real update otherwise pays seconds of one-time set-up
    jobs = [(f"warm_{i}", rng.rand(1024, 5).astype(np.float32), None) for i in range(2)]
    slow_loop_compute(jobs, SlowLoopConfig(), 0)


def gmm_cluster(stats: np.ndarray, num_clusters: int, seed: int):
    """Cluster stage (clustering.py:6-39)."""
    from sklearn.mixture import GaussianMixture

    gmm = GaussianMixture(n_components=num_clusters, random_state=int(seed))
    gmm.fit(stats)
    labels = gmm.predict(stats).astype(np.int32)
    return labels, {"weights": gmm.weights_, "means": gmm.means_, "covariances": gmm.covariances_}


def _relabel_by_grad_gini(labels: np.ndarray, stats: np.ndarray, k: int) -> np.ndarray:
    means = np.array([stats[labels == c, 0].mean() if np.any(labels == c) else np.inf for c in range(k)])
    order = np.argsort(means, kind="stable")
    lut = np.empty(k, dtype=np.int32)
    lut[order] = np.arange(k, dtype=np.int32)
    return lut[labels]


def _seed_from_key(key) -> int:
    """The reference derives the sklearn seed from a JAX PRNGKey (clustering.py:25); here `key` is an int or a
    torch.Generator."""
    if isinstance(key, torch.Generator):
        return int(torch.randint(0, 2 ** 31 - 1, (), generator=key).item())
    return int(key) % (2 ** 31 - 1)


def _world(group) -> int:
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def sense_layer(stats_leaf: Dict[str, torch.Tensor], pooled: bool, group=None) -> Optional[np.ndarray]:
    """(H,5) [grad_gini, grad_gdp, act_gini, act_gdp, act_variance] for one layer, or None if nothing was
    accumulated.  With `pooled`, the raw moments are all-reduced across ranks first."""
    return _sense_layers([stats_leaf], pooled, group)[0]


def _sense_layers_device(leaves, pooled: bool, group=None):
    """Device half of Sense for a list of stats_buffer leaf dicts: ONE cross-rank all-reduce, the finalisation kernels
    and ONE device-to-host copy into pinned memory, all enqueued on the current stream without waiting for them.
    Returns a handle for ``_sense_layers_host``."""
    world = _world(group)
    dev = leaves[0]["moments"].device
    sizes = [int(l["moments"].shape[0]) for l in leaves]
    n = len(leaves)
    if pooled:
        # fp64 = high + low fp32 parts of the double-single accumulators; counts ride along (exact in fp64)
        flat = torch.cat([ops.pooled_moments(l).reshape(-1) for l in leaves] +
                         [torch.stack([l["sample_count"].double() for l in leaves])])
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        counts = flat[-n:].to(torch.int32)
        out = torch.zeros((sum(sizes), 5), dtype=torch.float32, device=dev)
        off = 0
        for i, H in enumerate(sizes):
            # finalisation runs in the CUDA kernel (ops rejects host tensors: no CPU fallback); zero counts give zeros
            ops.stats_from_moments(flat[off * 5:(off + H) * 5].view(H, 5), counts[i:i + 1], out=out[off:off + H])
            off += H
        counts = counts.float()
    else:
        flat = torch.cat([l["neuron_stats"].reshape(-1) for l in leaves] +
                         [torch.stack([l["step_count"].float() for l in leaves])])
        counts = flat[-n:].clone()
        if world > 1:       # mean over ranks of each rank's mean over steps (moe_slow_loop.py:298 per rank)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        out = flat[:-n].view(-1, 5)
    pin = dev.type == "cuda"
    host = torch.empty(out.shape, dtype=torch.float32, pin_memory=pin)
    counts_h = torch.empty(n, dtype=torch.float32, pin_memory=pin)
    host.copy_(out, non_blocking=True)
    counts_h.copy_(counts, non_blocking=True)
    event = None
    if dev.type == "cuda":
        event = torch.cuda.Event()
        event.record()
    return {"host": host, "counts": counts_h, "event": event, "sizes": sizes, "pooled": pooled, "world": world}


def _sense_layers_host(handle):
    """Host half: wait for the copy (only this thread blocks) and cut the flat array into per-layer (H,5) arrays."""
    if handle["event"] is not None:
        handle["event"].synchronize()
    host, counts_h = handle["host"].numpy(), handle["counts"].numpy()
    res, off = [], 0
    for i, H in enumerate(handle["sizes"]):
        if counts_h[i] == 0:
            res.append(None)
        elif handle["pooled"]:
            res.append(host[off:off + H].copy())
        else:
            res.append(host[off:off + H] / (float(counts_h[i]) * handle["world"]))
        off += H
    return res


def _sense_layers(leaves, pooled: bool, group=None):
    """Sense a list of stats_buffer leaf dicts with ONE cross-rank all-reduce and ONE device-to-host copy (the slow
    loop stalls the training thread, so no per-layer synchronisation).  Returns a list of (H,5) arrays / None."""
    if not leaves:
        return []
    return _sense_layers_host(_sense_layers_device(leaves, pooled, group))


def _find_layers(mutable_variables):
    found = []

    def walk(st, asg, path):
        if isinstance(st, dict) and "neuron_stats" in st:
            found.append(("/".join(path), st, asg))
            return
        if isinstance(st, dict):
            for k in sorted(st):
                walk(st[k], asg[k], path + (k,))

    walk(mutable_variables["stats_buffer"], mutable_variables["neuron_assignments"], ())
    return found


def slow_loop_sense(mutable_variables, config: SlowLoopConfig, group=None):
    """SENSE for every layer (device work + one cross-rank all-reduce, layers in a fixed order on every rank).
    Returns the job list [(layer path, (H,5) stats on the host or None, that layer's assignment leaf dict)]."""
    found = _find_layers(mutable_variables)
    stats = _sense_layers([st for _, st, _ in found], config.pooled, group)
    return [(name, s, asg) for (name, _, asg), s in zip(found, stats)]


def slow_loop_compute(jobs, config: SlowLoopConfig, seed: int):
    """CLUSTER + SMOOTH per layer on host threads (pure host work on the sensed statistics: each layer is independent
    and seeded, so the result does not depend on the scheduling; safe to run while the device executes fast steps).
    Returns {layer path: (raw labels (H,), smoothed grid, gmm parameters)}."""
    def cluster_and_smooth(stats):
        H = stats.shape[0]
        k = config.num_clusters
        if config.forde_lite:
            labels = np.where(stats[:, 0] > config.forde_lite_threshold, 2, 0).astype(np.int32)
            gmm_params = None
        else:
            labels, gmm_params = gmm_cluster(np.nan_to_num(stats), k, seed)             # CLUSTER
            if config.canonicalize:
                labels = _relabel_by_grad_gini(labels, stats, k)
        grid = labels.reshape(default_grid(H))                                          # SMOOTH
        smoothed = box_smooth(grid, config.kernel_size, max(k, 3) if config.forde_lite else k)
        return labels, smoothed, gmm_params

    from concurrent.futures import ThreadPoolExecutor
    from contextlib import nullcontext
    try:    # (H, 5) problems are far too small for multi-threaded BLAS / OpenMP: with the library pools left on,
        # one GaussianMixture.fit takes 23 ms instead of 8 ms and 8 of them at once fight over the cores
        from threadpoolctl import threadpool_limits
        one_thread_kernels = threadpool_limits(limits=1)
    except ImportError:
        one_thread_kernels = nullcontext()
    todo = [(name, stats) for name, stats, _ in jobs if stats is not None]
    workers = max(1, min(os.cpu_count() or 1, 16, len(todo)))
    with one_thread_kernels, ThreadPoolExecutor(max_workers=workers) as pool:
        return dict(zip([n for n, _ in todo], pool.map(lambda t: cluster_and_smooth(t[1]), todo)))


def computes_here(config: SlowLoopConfig, group=None) -> bool:
    """With ``broadcast`` only rank 0 clusters and smooths (its maps are the ones every rank writes anyway); eight ranks
    each running the same scikit-learn fits on 16 host threads oversubscribe the box's cores and stretch an update
    from 0.7 s to 3 s (measured at 4 and 8 ranks, profiles/r2_c5_fit_*)."""
    if _world(group) > 1 and config.broadcast:
        return dist.get_rank(group) == 0
    return True


def slow_loop_actuate(jobs, results, config: SlowLoopConfig, diagnostics, group=None) -> None:
    """ACTUATE in layer order: overwrite each layer's ``assignments`` in place and fill ``diagnostics["layers"]``.
    Multi-rank with ``config.broadcast``: the maps every rank writes are rank 0's (one broadcast of all layers);
    the other ranks pass empty ``results``."""
    k = config.num_clusters
    live = [(name, stats, asg) for name, stats, asg in jobs if stats is not None]
    for name, stats, _ in jobs:
        if stats is None:
            diagnostics["layers"][name] = {"skipped": "no statistics accumulated"}
    if not live:
        return
    dev = live[0][2]["assignments"].device
    new_maps = [results[name][1].reshape(-1).astype(np.int32) if name in results
                else np.zeros(stats.shape[0], np.int32) for name, stats, _ in live]
    if _world(group) > 1 and config.broadcast:
        flat = torch.from_numpy(np.concatenate(new_maps)).to(dev)
        dist.broadcast(flat, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        flat_h, off = flat.cpu().numpy(), 0
        for i, m in enumerate(new_maps):
            new_maps[i] = flat_h[off:off + m.size]
            off += m.size
    old = torch.cat([asg["assignments"].reshape(-1) for _, _, asg in live]).cpu().numpy()    # one device-to-host copy
    off = 0
    for (name, stats, asg), new in zip(live, new_maps):
        labels, _, gmm_params = results.get(name, (None, None, None))
        smoothed = new.reshape(default_grid(new.size))
        changed = int((old[off:off + new.size] != new).sum())
        off += new.size
        asg["assignments"].copy_(torch.from_numpy(np.ascontiguousarray(new)))           # in place: bumps _version
        diagnostics["layers"][name] = {
            "census": np.bincount(new, minlength=k).tolist(),                           # README.md:66 "The Census"
            "raw_census": np.bincount(labels, minlength=k).tolist() if labels is not None else None,
            "changed": changed,
            "brain_scan": smoothed,                                                     # README.md:64
            "features": stats,                                                          # README.md:65
            "gmm": gmm_params,
        }


def _zero_tree(t) -> None:
    """RESET: zero every leaf of stats_buffer (moe_slow_loop.py:441-449)."""
    if isinstance(t, dict):
        for v in t.values():
            _zero_tree(v)
    else:
        t.zero_()


def slow_loop_step(params, mutable_variables, config: SlowLoopConfig, key, epoch: int = 0, step: int = 0,
                   group=None):
    """One brain update: Sense -> Cluster -> Smooth -> Actuate -> Reset.  `mutable_variables` holds ``stats_buffer``
    and ``neuron_assignments`` trees keyed by ``layer_{i}`` (any nesting above that is walked).
    Returns (params, mutable_variables, diagnostics)."""
    import time
    t0 = time.perf_counter()
    diagnostics = {"epoch": epoch, "step": step, "layers": {}}
    jobs = slow_loop_sense(mutable_variables, config, group)
    t1 = time.perf_counter()
    results = slow_loop_compute(jobs, config, _seed_from_key(key)) if computes_here(config, group) else {}
    t2 = time.perf_counter()
    slow_loop_actuate(jobs, results, config, diagnostics, group)
    _zero_tree(mutable_variables["stats_buffer"])
    t3 = time.perf_counter()
    # host seconds the training thread spent in this update (the device idles meanwhile)
    diagnostics["seconds"] = {"sense": t1 - t0, "cluster_smooth": t2 - t1, "actuate_reset": t3 - t2, "total": t3 - t0}
    return params, mutable_variables, diagnostics


class OverlappedSlowLoop:
    """The same brain update with Cluster + Smooth taken off the training thread.

    The host work of one update (scikit-learn GMM + smoothing per layer, ~0.2 s for the 24 layers of the ViT-B/16
    dual encoder) would otherwise stall the device for several fast steps.  ``start`` senses (device -> host) and
    resets the accumulators at step s, hands Cluster + Smooth to a worker thread, and ``poll`` actuates at step
    s + delay -- a FIXED delay, so every data-parallel rank switches brain maps at the same step and the replicas
    stay identical.  The device keeps running fast steps (CUDA-graph replays) under the old map meanwhile.
    Deviation from the reference's synchronous loop (moe_slow_loop.py:253-464), hence opt-in: the new map takes
    effect ``delay`` steps late and the statistics of those steps count towards the next update."""

    def __init__(self, config: SlowLoopConfig, delay: int, group=None):
        if delay < 1:
            raise ValueError("delay must be at least one step (use slow_loop_step for the synchronous update)")
        from concurrent.futures import ThreadPoolExecutor
        self.config, self.delay, self.group = config, int(delay), group
        self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="forde-slow-loop")
        self._pending = None    # (future, due step, jobs, diagnostics)

    @property
    def busy(self) -> bool:
        return self._pending is not None

    def start(self, mutable_variables, key, epoch: int = 0, step: int = 0):
        """Sense + Reset now, Cluster + Smooth in the background.  Returns the diagnostics of a still-pending
        earlier update that had to be actuated first (None normally)."""
        import time
        flushed = self.flush()
        t0 = time.perf_counter()
        diagnostics = {"epoch": epoch, "step": step, "layers": {}}
        # Sense is only ENQUEUED here (all-reduce, finalisation kernels, copy into pinned memory, then Reset, all
        # stream-ordered behind the fast step in flight): the training thread does not wait for the device, the
        # worker thread does before it clusters.
        found = _find_layers(mutable_variables)
        handle = _sense_layers_device([st for _, st, _ in found], self.config.pooled, self.group) if found else None
        _zero_tree(mutable_variables["stats_buffer"])
        diagnostics["seconds"] = {"sense": time.perf_counter() - t0}
        config, seed = self.config, _seed_from_key(key)
        here = computes_here(config, self.group)

        def work():
            stats = _sense_layers_host(handle) if handle is not None else []
            jobs = [(name, st, asg) for (name, _, asg), st in zip(found, stats)]
            return jobs, (slow_loop_compute(jobs, config, seed) if here else {})

        fut = self._pool.submit(work)
        self._pending = (fut, step + self.delay, None, diagnostics)
        return flushed

    def poll(self, step: int):
        """Call once per fast step (before it).  Actuates and returns the diagnostics when ``step`` has reached the
        update's due step (waiting for the worker if it is not done yet), else returns None."""
        if self._pending is None or step < self._pending[1]:
            return None
        return self.flush()

    def flush(self):
        """Actuate a pending update now, whatever the step (end of training, or a new update is due)."""
        if self._pending is None:
            return None
        import time
        fut, due, _, diagnostics = self._pending
        self._pending = None
        t0 = time.perf_counter()
        jobs, results = fut.result()           # normally done long ago; the wait (if any) is training-thread time
        slow_loop_actuate(jobs, results, self.config, diagnostics, self.group)
        sec = diagnostics.setdefault("seconds", {})
        sec["actuate_reset"] = time.perf_counter() - t0
        sec["total"] = sec.get("sense", 0.0) + sec["actuate_reset"]     # on the training thread; Cluster+Smooth overlapped
        diagnostics["actuated_at_step"] = due
        return diagnostics

    def close(self):
        self._pool.shutdown(wait=True)


# ------------------------------------------------------------------------------------------ MoE variant (config 4)
def box_smooth_3d(grid: np.ndarray, kernel_size: int = 3, num_clusters: int = 3) -> np.ndarray:
    """3-D majority smoothing of expert labels (smoothing.py:94-171): per-label k^3 box filter with 'same'
    convolution, edge padding for grids thinner than the kernel, label of highest density (lowest id on ties)."""
    from scipy.signal import convolve

    box = np.full((kernel_size,) * 3, 1.0 / kernel_size ** 3, dtype=np.float32)
    d, h, w = grid.shape
    pads = [max(0, kernel_size - n) for n in grid.shape]
    padding = tuple((p // 2, p - p // 2) for p in pads)
    dens = []
    for c in range(num_clusters):
        plane = (grid == c).astype(np.float32)
        if any(pads):
            plane = np.pad(plane, padding, "edge")
        sm = convolve(plane, box, mode="same").astype(np.float32)
        dens.append(sm[padding[0][0]:padding[0][0] + d, padding[1][0]:padding[1][0] + h, padding[2][0]:padding[2][0] + w])
    return np.stack(dens, -1).argmax(-1).astype(np.int32)


def router_adjustments(assignments: np.ndarray, usage: np.ndarray, target_balance: float = 0.1) -> np.ndarray:
    """Router-bias nudges (moe_slow_loop.py:211-250): push towards uniform usage, capped at +-0.1, halved for the
    'specialist' cluster (id 1)."""
    usage = np.asarray(usage, np.float32)
    adj = np.clip(-(usage - np.float32(1.0 / usage.shape[0])) * np.float32(target_balance), -0.1, 0.1).astype(np.float32)
    return np.where(np.asarray(assignments) == 1, adj * np.float32(0.5), adj).astype(np.float32)


def _walk_leaves(tree, path=()):
    if isinstance(tree, dict):
        for k, v in tree.items():
            yield from _walk_leaves(v, path + (k,))
    else:
        yield path, tree


def moe_slow_loop_step(model_params, mutable_variables, config, key, epoch: int = 0, step: int = 0, group=None):
    """Slow loop of the MoE variant, same contract as src/forde/moe_slow_loop.py:253-464:
    ``(params, mutable_variables, diagnostics)`` in and out are plain nested dicts; ``config`` needs ``num_layers`` and
    ``num_experts``.  SENSE: every ``expert_usage`` leaf of ``stats_buffer`` (one per layer, found by name like
    moe_slow_loop.py:145-170) divided by the largest ``step_count`` (:298); ranks average their usage first.
    CLUSTER: GMM(3) on (mean, variance over layers) per expert (:314-323).  SMOOTH: labels on a (1, gh, gw) grid,
    3-D box filter (:337-363).  ANALYSE: load imbalance, relative routing entropy (:372-381).  ACTUATE: add the bias
    nudges to every ``router_linear/bias`` of matching shape, in place (:418-439).  RESET: zero every leaf of
    ``stats_buffer`` (:441-449)."""
    E, L = int(config.num_experts), int(config.num_layers)
    sb = mutable_variables.get("stats_buffer", {})
    usages, step_count = [], 0
    for path, leaf in _walk_leaves(sb):
        if path[-1] == "expert_usage" and leaf.shape[0] == E:
            usages.append(leaf)
        elif path[-1] == "step_count":
            step_count = max(step_count, int(leaf.item()))
    if step_count == 0 or not usages:
        return model_params, mutable_variables, {"skipped": True}
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    usage = torch.stack([u.detach().float() for u in usages[:L]] + [torch.zeros_like(usages[0], dtype=torch.float32)] * max(0, L - len(usages)))
    if world > 1:
        dist.all_reduce(usage, op=dist.ReduceOp.SUM, group=group)
        usage = usage / world
    usage = (usage / step_count).cpu().numpy().astype(np.float32)                     # (L, E)
    usage_mean, usage_var = usage.mean(0), usage.var(0)
    labels, gmm_params = gmm_cluster(np.stack([usage_mean, usage_var], -1), 3, _seed_from_key(key))
    gh = int(np.sqrt(E))
    while E % gh != 0:
        gh -= 1
    smoothed = box_smooth_3d(labels.reshape(1, gh, E // gh), 3, 3).reshape(-1)
    imbalance = float(np.abs(usage_mean - 1.0 / E).mean())
    entropy = float(-(usage_mean * np.log(usage_mean + 1e-8)).sum() / np.log(E))
    adj = router_adjustments(smoothed, usage_mean)
    updated = 0
    for path, leaf in _walk_leaves(model_params):
        if "router_linear" in path and path[-1] == "bias" and tuple(leaf.shape) == adj.shape:
            with torch.no_grad():
                leaf.add_(torch.from_numpy(adj).to(leaf.device, leaf.dtype))               # in place: views stay valid
            updated += 1
    for _, leaf in _walk_leaves(sb):
        leaf.zero_()
    return model_params, mutable_variables, {
        "expert_usage": usage_mean, "assignments": smoothed, "raw_assignments": labels, "load_imbalance": imbalance,
        "routing_entropy": entropy, "adjustments": adj, "step_count": step_count, "router_biases_updated": updated,
        "gmm": gmm_params, "epoch": epoch, "step": step,
    }
