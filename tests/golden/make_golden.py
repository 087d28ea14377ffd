"""Generate the golden fixtures in this directory FROM THE REFERENCE'S OWN SOURCE FILES.

JAX / Flax are not installable in the build container (no network, not in the wheelhouse), so the
reference modules cannot be imported as they are.  Their pure functions only use a handful of ``jax.numpy``
calls, so this script installs a minimal NumPy-backed stand-in for the ``jax`` package, imports
``/root/reference/src/forde/sensing.py`` and ``smoothing.py`` *unchanged* from where they lie, and records
their outputs on seeded inputs.  What is pinned is therefore the reference's algorithm as written (guards,
reduction axes, stat order, padding rule, argmax tie rule) evaluated with NumPy fp32 arithmetic instead of
XLA's; summation order may differ from XLA in the last ulp.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The .npz files it writes are committed; nothing at test time reads /root/reference.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def install_jax_stub():
    import scipy.signal

    class _At:
        def __init__(self, arr):
            self.arr = arr

        def __getitem__(self, idx):
            arr = self.arr

            class _Setter:
                @staticmethod
                def set(v):
                    out = np.array(arr, copy=True).view(_JArr)
                    out[idx] = v
                    return out
            return _Setter

    class _JArr(np.ndarray):
        """ndarray with jax's functional ``x.at[i].set(v)``."""

        @property
        def at(self):
            return _At(self)

    class _F32NumPy(types.ModuleType):
        """numpy with float32 defaults for the constructors the reference uses."""

        def __getattr__(self, name):
            return getattr(np, name)

        @staticmethod
        def ones(shape, dtype=np.float32):
            return np.ones(shape, dtype=dtype)

        @staticmethod
        def zeros(shape, dtype=np.float32):
            return np.zeros(shape, dtype=dtype).view(_JArr)

        @staticmethod
        def sqrt(x):
            # jnp.sqrt of a Python int yields float32
            if isinstance(x, (int, float)):
                return np.sqrt(np.float32(x))
            return np.sqrt(x)

        @staticmethod
        def where(c, a, b):
            r = np.where(c, a, b)
            return r.astype(np.float32) if r.dtype == np.float64 else r

        @staticmethod
        def asarray(x, dtype=None):
            return np.asarray(x, dtype=dtype)

    jnp = _F32NumPy("jax.numpy")
    jnp.ndarray = np.ndarray
    jnp.float32 = np.float32
    jnp.int32 = np.int32

    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.vmap = lambda f: (lambda xs: np.stack([f(x) for x in xs]))
    nn = types.ModuleType("jax.nn")
    nn.one_hot = lambda x, n: (np.asarray(x)[..., None] == np.arange(n)).astype(np.float32)

    def softmax(x, axis=-1):
        x = np.asarray(x, np.float32)
        e = np.exp(x - x.max(axis=axis, keepdims=True))
        return (e / e.sum(axis=axis, keepdims=True)).astype(np.float32)
    nn.softmax = softmax
    jax.nn = nn
    lax = types.ModuleType("jax.lax")

    def lax_top_k(x, k):
        # jax.lax.top_k: values descending, equal values keep the lower index first (stable)
        idx = np.argsort(-np.asarray(x), axis=-1, kind="stable")[..., :k]
        return np.take_along_axis(np.asarray(x), idx, axis=-1), idx.astype(np.int32)
    lax.top_k = lax_top_k
    jax.lax = lax
    jnp.bincount = lambda x, length=None, minlength=0: np.bincount(
        np.asarray(x).reshape(-1), minlength=length or minlength)[: length or None]
    jsp = types.ModuleType("jax.scipy")
    jsig = types.ModuleType("jax.scipy.signal")
    jsig.convolve2d = lambda a, k, mode="full": scipy.signal.convolve2d(a, k, mode=mode).astype(np.float32)
    jsig.convolve = lambda a, k, mode="full": scipy.signal.convolve(a, k, mode=mode).astype(np.float32)
    jsp.signal = jsig
    jax.scipy = jsp
    rnd = types.ModuleType("jax.random")
    rnd.PRNGKey = lambda seed: np.array([0, int(seed)], np.uint32)
    rnd.split = lambda key, num=2: [np.array([int(key[0]) + i + 1, int(key[1])], np.uint32) for i in range(num)]
    # clustering.py:25 draws the sklearn seed with jax.random.randint; the stub derives a deterministic int instead
    # (threefry is not reproduced), so GMM-dependent outputs are pinned only where they do not depend on the seed
    rnd.randint = lambda key, shape, lo, hi: np.int64((int(key[0]) * 1000003 + int(key[1])) % (hi - lo) + lo)
    jax.random = rnd
    tree = types.ModuleType("jax.tree")

    def tree_map(fn, t):
        return {k: tree_map(fn, v) for k, v in t.items()} if isinstance(t, dict) else fn(t)
    tree.map = tree_map
    jax.tree = tree
    flax = types.ModuleType("flax")
    fcore = types.ModuleType("flax.core")
    fcore.unfreeze = lambda t: ({k: fcore.unfreeze(v) for k, v in t.items()} if isinstance(t, dict) else t)
    ftu = types.ModuleType("flax.traverse_util")

    def flatten_dict(t, prefix=()):
        out = {}
        for k, v in t.items():
            if isinstance(v, dict):
                out.update(flatten_dict(v, prefix + (k,)))
            else:
                out[prefix + (k,)] = v
        return out

    def unflatten_dict(flat):
        out = {}
        for path, v in flat.items():
            d = out
            for k in path[:-1]:
                d = d.setdefault(k, {})
            d[path[-1]] = v
        return out
    ftu.flatten_dict, ftu.unflatten_dict = flatten_dict, unflatten_dict
    flax.core, flax.traverse_util = fcore, ftu
    # flax.linen: just enough for `src/forde/moe.py` to IMPORT (class bodies, decorators).  Its array-only helper
    # methods (_top_k_gating, _compute_expert_outputs, _load_balancing_loss) are then called unbound; nothing here
    # emulates flax's module system.
    linen = types.ModuleType("flax.linen")

    class _Module:
        pass
    linen.Module = _Module
    linen.compact = lambda f: f
    flax.linen = linen
    sys.modules["flax.linen"] = linen
    sys.modules["jax.lax"] = lax
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "jax.nn": nn, "jax.scipy": jsp, "jax.scipy.signal": jsig,
                        "jax.random": rnd, "jax.tree": tree, "flax": flax, "flax.core": fcore,
                        "flax.traverse_util": ftu})


def load_ref(name):
    path = os.path.join(REF, "src", "forde", name + ".py")
    spec = importlib.util.spec_from_file_location("ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    install_jax_stub()
    sensing = load_ref("sensing")
    smoothing = load_ref("smoothing")
    rng = np.random.RandomState(1234)

    # ---- sensing: hoyer known answers (tests/verify_hoyer.py:10-22) + random vectors
    hoy_in = [np.zeros(10, np.float32), np.ones(10, np.float32), np.eye(10, dtype=np.float32)[0],
              rng.randn(64).astype(np.float32), rng.randn(7, 33).astype(np.float32), np.array([3.0], np.float32)]
    hoy_out = [np.asarray(sensing.hoyer_sparsity(v), np.float32) for v in hoy_in]
    np.savez(os.path.join(OUT, "sensing_hoyer.npz"), **{f"in{i}": v for i, v in enumerate(hoy_in)},
             **{f"out{i}": v for i, v in enumerate(hoy_out)})

    # ---- sensing: calculate_neuron_stats, incl. the sensing.py:86-97 spike / all-ones cases
    cases = {}
    a = rng.randn(8, 128).astype(np.float32); g = rng.randn(8, 128).astype(np.float32)
    cases["rand2d"] = (a, g)
    a3 = (rng.randn(4, 9, 48) * 2 + 0.5).astype(np.float32); g3 = (rng.randn(4, 9, 48) * 1e-3).astype(np.float32)
    cases["rand3d"] = (a3, g3)
    spike = np.zeros((8, 128), np.float32); spike[0, 0] = 10.0
    cases["spike"] = (spike, g)
    cases["ones"] = (np.ones((8, 128), np.float32), g)
    sparse_g = np.where(rng.rand(32, 16) < 0.1, rng.randn(32, 16), 0).astype(np.float32)
    cases["sparse_grad"] = (rng.randn(32, 16).astype(np.float32), sparse_g)
    cases["single_sample"] = (rng.randn(1, 16).astype(np.float32), rng.randn(1, 16).astype(np.float32))
    blob = {}
    for k, (a_, g_) in cases.items():
        blob[k + "_act"] = a_
        blob[k + "_grad"] = g_
        blob[k + "_stats"] = np.asarray(sensing.calculate_neuron_stats(a_, g_), np.float32)
    np.savez(os.path.join(OUT, "sensing_stats.npz"), **blob)

    # ---- smoothing: 2-D grids incl. tiny grids that exercise the padding rule (smoothing.py:59-68)
    blob = {}
    for i, (shape, k, nc) in enumerate([((16, 32), 3, 3), ((48, 64), 3, 3), ((2, 3), 3, 3), ((1, 8), 3, 3),
                                        ((12, 12), 5, 4), ((32, 64), 3, 3)]):
        grid = rng.randint(0, nc, size=shape).astype(np.int32)
        out = np.asarray(smoothing.smooth_assignments(smoothing.assignments_to_grid(grid.reshape(-1), shape), k, nc))
        blob[f"grid{i}"] = grid
        blob[f"params{i}"] = np.array([k, nc], np.int32)
        blob[f"out{i}"] = out.astype(np.int32)
    np.savez(os.path.join(OUT, "smoothing.npz"), **blob)

    # ---- MoE slow loop (moe_slow_loop.py): 3-D smoothing, router adjustments, and one whole step on the pytrees of
    #      tests/verify_reset.py:17-35 (uniform usage: every expert lands in one GMM cluster whatever the seed)
    sys.path.insert(0, os.path.join(REF, "src", "forde"))
    msl = load_ref("moe_slow_loop")
    blob = {}
    for i, shape in enumerate([(1, 2, 4), (1, 1, 4), (3, 4, 4), (2, 3, 5)]):
        grid = rng.randint(0, 3, size=shape).astype(np.int32)
        blob[f"grid3d_{i}"] = grid
        blob[f"out3d_{i}"] = np.asarray(smoothing.smooth_assignments_3d(grid, 3, 3)).astype(np.int32)
    for i in range(3):
        assign = rng.randint(0, 3, size=8).astype(np.int32)
        usage = rng.dirichlet(np.ones(8)).astype(np.float32)
        blob[f"adj_assign_{i}"], blob[f"adj_usage_{i}"] = assign, usage
        blob[f"adj_out_{i}"] = np.asarray(msl.compute_router_adjustments(assign, usage), np.float32)

    class Cfg:
        num_layers, num_experts = 1, 4
    params = {"params": {"layer_0": {"moe": {"router_linear": {"bias": np.zeros(4, np.float32)}}}}}
    mv = {"stats_buffer": {"layer_0": {"moe": {"expert_usage": np.array([4.0, 3.0, 2.0, 1.0], np.float32),
                                               "step_count": np.array(10, np.int32)}}}}
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        new_p, new_mv, diag = msl.moe_slow_loop_step(params, mv, Cfg, sys.modules["jax.random"].PRNGKey(42), 0, 100)
    blob["step_usage_in"] = np.array([4.0, 3.0, 2.0, 1.0], np.float32)
    blob["step_usage_mean"] = np.asarray(diag["expert_usage"], np.float32)
    blob["step_bias_out"] = np.asarray(new_p["params"]["layer_0"]["moe"]["router_linear"]["bias"], np.float32)
    blob["step_adjustments"] = np.asarray(diag["adjustments"], np.float32)
    blob["step_assignments"] = np.asarray(diag["assignments"], np.int32)
    blob["step_imbalance"] = np.float32(diag["load_imbalance"])
    blob["step_entropy"] = np.float32(diag["routing_entropy"])
    blob["step_usage_after"] = np.asarray(new_mv["stats_buffer"]["layer_0"]["moe"]["expert_usage"], np.float32)
    blob["step_count_after"] = np.asarray(new_mv["stats_buffer"]["layer_0"]["moe"]["step_count"], np.int32)
    np.savez(os.path.join(OUT, "moe_slow_loop.npz"), **blob)
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


def moe_layer_golden():
    """MoE layer routing math from the reference's own `src/forde/moe.py` (its array-only helper methods called
    unbound with a plain namespace as `self`): top-k gating + renormalisation (moe.py:116-135), the gather /
    weighted combine (moe.py:137-187) on given per-expert outputs, and the load-balancing loss (moe.py:189-222).
    `jax.lax.top_k`, `jax.nn.softmax` and `jnp.bincount` come from the stub above (jax's documented semantics);
    what is pinned is how the reference composes them."""
    install_jax_stub()
    moe = load_ref("moe")
    rng = np.random.RandomState(4321)
    blob = {}
    for i, (B, S, d, E, k) in enumerate([(2, 5, 8, 4, 2), (1, 7, 16, 8, 2), (3, 4, 8, 8, 1), (2, 3, 4, 4, 4)]):
        ns = types.SimpleNamespace(num_experts=E, top_k=k, aux_loss_weight=0.01)
        logits = rng.randn(B, S, E).astype(np.float32)
        if i == 1:
            logits[0, 0, :] = np.array([1, 3, 3, 2, 3, 0, -1, 3], np.float32)      # ties
        x = rng.randn(B, S, d).astype(np.float32)
        expert_out = rng.randn(E, B, S, d).astype(np.float32)                      # what expert e returns for x
        experts = [(lambda xx, e=e: expert_out[e]) for e in range(E)]
        idx, w = moe.MoELayer._top_k_gating(ns, logits)
        out = moe.MoELayer._compute_expert_outputs(ns, x, experts, idx, w)
        probs = sys.modules["jax.nn"].softmax(logits, axis=-1)
        aux = moe.MoELayer._load_balancing_loss(ns, probs, idx)
        blob.update({f"logits{i}": logits, f"x{i}": x, f"expert_out{i}": expert_out,
                     f"cfg{i}": np.array([E, k], np.int32), f"idx{i}": np.asarray(idx, np.int32),
                     f"w{i}": np.asarray(w, np.float32), f"out{i}": np.asarray(out, np.float32),
                     f"probs{i}": np.asarray(probs, np.float32), f"aux{i}": np.float32(aux)})
    np.savez(os.path.join(OUT, "moe_layer.npz"), **blob)
    print("wrote moe_layer.npz")


def clustering_golden():
    """`src/forde/clustering.py:6-39` run as it is (scikit-learn is installed).  The sklearn seed it derives with
    jax.random.randint comes from the stub's deterministic stand-in and is recorded, so what is pinned is the call
    sequence (GaussianMixture(n_components, random_state=seed).fit(stats).predict(stats), label dtype, the returned
    parameter dict) for a GIVEN seed -- not jax's PRNG."""
    install_jax_stub()
    clustering = load_ref("clustering")
    rnd = sys.modules["jax.random"]
    rng = np.random.RandomState(77)
    blob = {}
    for i, (H, k, key_seed) in enumerate([(192, 3, 0), (512, 3, 42), (96, 4, 7)]):
        centres = rng.randn(k, 5) * 3
        stats = (centres[rng.randint(0, k, H)] + 0.3 * rng.randn(H, 5)).astype(np.float32)
        key = rnd.PRNGKey(key_seed)
        labels, params = clustering.cluster_neurons_gmm(stats, k, key)
        blob[f"stats{i}"] = stats
        blob[f"k{i}"] = np.int32(k)
        blob[f"seed{i}"] = np.int64(rnd.randint(key, (), 0, 2 ** 31 - 1))
        blob[f"labels{i}"] = np.asarray(labels, np.int32)
        blob[f"means{i}"] = np.asarray(params["means"], np.float64)
        blob[f"weights{i}"] = np.asarray(params["weights"], np.float64)
    import sklearn
    blob["sklearn_version"] = np.array(sklearn.__version__)
    np.savez(os.path.join(OUT, "clustering.npz"), **blob)
    print("wrote clustering.npz")


if __name__ == "__main__":
    if "--moe-layer-only" in sys.argv:      # adds moe_layer.npz without rewriting the other fixtures
        moe_layer_golden()
    elif "--clustering-only" in sys.argv:
        clustering_golden()
    else:
        main()
        moe_layer_golden()
        clustering_golden()
