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

    class _F32NumPy(types.ModuleType):
        """numpy with float32 defaults for the constructors the reference uses."""

        def __getattr__(self, name):
            return getattr(np, name)

        @staticmethod
        def ones(shape, dtype=np.float32):
            return np.ones(shape, dtype=dtype)

        @staticmethod
        def zeros(shape, dtype=np.float32):
            return np.zeros(shape, dtype=dtype)

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
    jax.nn = nn
    jsp = types.ModuleType("jax.scipy")
    jsig = types.ModuleType("jax.scipy.signal")
    jsig.convolve2d = lambda a, k, mode="full": scipy.signal.convolve2d(a, k, mode=mode).astype(np.float32)
    jsig.convolve = lambda a, k, mode="full": scipy.signal.convolve(a, k, mode=mode).astype(np.float32)
    jsp.signal = jsig
    jax.scipy = jsp
    rnd = types.ModuleType("jax.random")
    jax.random = rnd
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "jax.nn": nn, "jax.scipy": jsp, "jax.scipy.signal": jsig,
                        "jax.random": rnd})


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
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
