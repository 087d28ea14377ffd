"""The oracle against every pin the reference offers, plus self-consistency checks of
the parts that are parity-unpinned (StatefulLayer, CLIP loss): finite differences and the explicit STE rule."""
import os

import numpy as np
import pytest

from oracle import forde_oracle as O


# ------------------------------------------------------------------ sensing: golden vectors from sensing.py
def test_hoyer_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "sensing_hoyer.npz"))
    for i in range(6):
        got = O.hoyer_sparsity(z[f"in{i}"])
        np.testing.assert_allclose(got, z[f"out{i}"], rtol=1e-6, atol=1e-7)


def test_hoyer_known_answers():
    # tests/verify_hoyer.py:10-22,34-38 (printed there, asserted here)
    assert float(O.hoyer_sparsity(np.zeros(10, np.float32))) == 0.0
    assert abs(float(O.hoyer_sparsity(np.ones(10, np.float32)))) < 1e-6
    assert abs(float(O.hoyer_sparsity(np.eye(10, dtype=np.float32)[0])) - 1.0) < 1e-6
    assert float(O.hoyer_sparsity(np.array([3.0], np.float32))) == 0.0  # n == 1 guard, sensing.py:25,32


@pytest.mark.parametrize("case", ["rand2d", "rand3d", "spike", "ones", "sparse_grad", "single_sample"])
def test_neuron_stats_golden(golden_dir, case):
    z = np.load(os.path.join(golden_dir, "sensing_stats.npz"))
    got = O.calculate_neuron_stats(z[case + "_act"], z[case + "_grad"])
    np.testing.assert_allclose(got, z[case + "_stats"], rtol=1e-5, atol=1e-6)


def test_neuron_stats_spike_known_answer():
    # sensing.py:86-97: one spike of 10 in an (8,128) zero array -> neuron 0 maximally sparse
    a = np.zeros((8, 128), np.float32); a[0, 0] = 10.0
    s = O.calculate_neuron_stats(a, np.ones_like(a))
    assert abs(s[0, 2] - 1.0) < 1e-6 and abs(s[0, 3] - 1.25) < 1e-6 and abs(s[0, 4] - 10.9375) < 1e-5
    assert s[1, 2] == 0.0  # an all-zero column
    s1 = O.calculate_neuron_stats(np.ones((8, 128), np.float32), np.ones((8, 128), np.float32))
    assert np.all(np.abs(s1[:, 2]) < 1e-6)


def test_moments_equal_direct_stats():
    rng = np.random.RandomState(0)
    a = (rng.randn(300, 40) * 1.5 + 0.3).astype(np.float32); g = (rng.randn(300, 40) * 1e-3).astype(np.float32)
    m, n = O.raw_moments(a, g)
    np.testing.assert_allclose(O.stats_from_moments(m, n), O.calculate_neuron_stats(a, g), rtol=2e-5, atol=1e-6)
    # additivity over sample blocks (what lets CTAs, steps and ranks be summed)
    m1, n1 = O.raw_moments(a[:100], g[:100]); m2, n2 = O.raw_moments(a[100:], g[100:])
    np.testing.assert_allclose(m1 + m2, m, rtol=1e-12)
    assert n1 + n2 == n


# ------------------------------------------------------------------ smoothing: golden vectors from smoothing.py
@pytest.mark.parametrize("i", range(6))
def test_smoothing_golden(golden_dir, i):
    z = np.load(os.path.join(golden_dir, "smoothing.npz"))
    k, nc = (int(v) for v in z[f"params{i}"])
    got = O.smooth_assignments(O.assignments_to_grid(z[f"grid{i}"].reshape(-1), z[f"grid{i}"].shape), k, nc)
    np.testing.assert_array_equal(got, z[f"out{i}"])


# ------------------------------------------------------------------ top-k / MoE pins
def test_topk_index_sets():
    # tests/test_moe.py:46-70
    x = np.array([[10.0, 2.0, 5.0, 8.0], [1.0, 9.0, 3.0, 4.0]], np.float32)
    _, idx = O.top_k(x, 2)
    assert set(idx[0].tolist()) == {0, 3} and set(idx[1].tolist()) == {1, 3}
    # tests/test_topk_correctness.py:7-23: index set equals argsort's last k
    rng = np.random.RandomState(42)
    y = rng.randn(32, 100).astype(np.float32)
    _, idx = O.top_k(y, 10)
    ref = np.argsort(y, axis=-1)[..., -10:]
    for r in range(32):
        assert set(idx[r].tolist()) == set(ref[r].tolist())


def test_topk_tie_rule():
    # jax.lax.top_k: descending, equal values -> lower index first
    vals, idx = O.top_k(np.array([1.0, 3.0, 3.0, 2.0, 3.0], np.float32), 2)
    assert idx.tolist() == [1, 2] and vals.tolist() == [3.0, 3.0]


def test_moe_router_probs_and_aux():
    rng = np.random.RandomState(1)
    x = rng.randn(64, 32).astype(np.float32)
    w = (rng.randn(32, 8) * 0.02).astype(np.float32)
    probs, idx, wts = O.moe_router(x, w, np.zeros(8, np.float32), 2)
    np.testing.assert_allclose(probs.sum(-1), 1.0, atol=1e-5)  # tests/test_moe.py:38
    np.testing.assert_allclose(wts.sum(-1), 1.0, atol=1e-6)
    aux, counts = O.moe_aux_loss(probs, idx, 8, 2, 0.01)
    assert counts.sum() == 64 * 2 and counts.dtype == np.int32
    # uniform routing gives aux = weight (Switch loss lower bound E * sum (1/E)(1/E) = 1)
    pu = np.full((64, 8), 1 / 8, np.float32); iu = np.stack([np.arange(64) % 8, (np.arange(64) + 1) % 8], -1)
    au, _ = O.moe_aux_loss(pu, iu, 8, 2, 0.01)
    assert abs(au - 0.01) < 1e-7


def test_moe_layer_matches_routed_only_computation():
    rng = np.random.RandomState(2)
    M, d, H, E, k = 24, 16, 32, 4, 2
    x = rng.randn(M, d).astype(np.float32)
    experts = [((rng.randn(d, H) / 4).astype(np.float32), np.zeros(H, np.float32),
                (rng.randn(H, d) / 6).astype(np.float32), np.zeros(d, np.float32)) for _ in range(E)]
    wr = (rng.randn(d, E) * 0.5).astype(np.float32)
    out, aux, probs, idx, counts = O.moe_layer_fwd(x, wr, np.zeros(E, np.float32), experts, k)
    # result-equivalent sparse evaluation (only the routed experts), the form the CUDA path uses
    _, idx2, w2 = O.moe_router(x, wr, np.zeros(E, np.float32), k)
    ref = np.zeros_like(x)
    for m in range(M):
        for j in range(k):
            wu, bu, wd, bd = experts[idx2[m, j]]
            ref[m] += w2[m, j] * (O.gelu_tanh(x[m] @ wu + bu) @ wd + bd)
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------ StatefulLayer (parity unpinned: self-consistency)
def _layer(seed=0, M=20, d=12, H=18):
    rng = np.random.RandomState(seed)
    x = rng.randn(M, d).astype(np.float32)
    wu = (rng.randn(d, H) / np.sqrt(d)).astype(np.float32); bu = (rng.randn(H) * 0.1).astype(np.float32)
    wd = (rng.randn(H, d) / np.sqrt(H)).astype(np.float32); bd = (rng.randn(d) * 0.1).astype(np.float32)
    return x, wu, bu, wd, bd


def test_multiplex_and_gate_rules():
    z = np.array([[-1.0, -1.0, -1.0, -1.0], [0.0, 0.0, 0.0, 0.0], [2.0, 2.0, 2.0, 2.0]], np.float32)
    a = np.array([0, 1, 2, 7])  # 7: outside {0,1,2} behaves as a specialist (D5)
    f = O.multiplex(z, a)
    np.testing.assert_allclose(f[0], [0.0, np.tanh(-1.0), 0.0, 0.0], rtol=1e-6)
    np.testing.assert_allclose(f[1], [0.0, 0.0, 0.0, 0.0])                      # step(0) = 0 (D3)
    np.testing.assert_allclose(f[2], [2.0, np.tanh(2.0), 1.0, 1.0], rtol=1e-6)
    np.testing.assert_allclose(O.gate_of(a), [1.0, 1.0, 0.1, 0.1])              # D4
    h = O.gated_residual(z, a)
    np.testing.assert_allclose(h[2], [4.0, np.tanh(2.0) + 2.0, 1.2, 1.2], rtol=1e-6)


def test_backward_finite_difference_on_smooth_units_and_ste_on_step_units():
    x, wu, bu, wd, bd = _layer()
    a = np.arange(18) % 3
    fw = O.stateful_layer_fwd(x, wu, bu, wd, bd, a, emulate_bf16=False)
    rng = np.random.RandomState(5)
    dy = rng.randn(*fw["y"].shape).astype(np.float32)
    bw = O.stateful_layer_bwd(x, wu, wd, a, fw["z"], fw["h"], dy, emulate_bf16=False)
    # STE: specialist units pass dh through with slope 1 (+ gate 0.1) regardless of z
    dh = dy @ wd.T
    np.testing.assert_allclose(bw["dz"][:, a == 2], 1.1 * dh[:, a == 2], rtol=1e-6)
    # finite differences of L = sum(y * dy) through the smooth (relu / tanh) units only
    keep = a != 2
    wu_s, bu_s, wd_s = wu[:, keep], bu[keep], wd[keep]
    a_s = a[keep]

    def loss(xx):
        f = O.stateful_layer_fwd(xx, wu_s, bu_s, wd_s, bd, a_s, emulate_bf16=False)
        return float((f["y"].astype(np.float64) * dy).sum())

    f_s = O.stateful_layer_fwd(x, wu_s, bu_s, wd_s, bd, a_s, emulate_bf16=False)
    b_s = O.stateful_layer_bwd(x, wu_s, wd_s, a_s, f_s["z"], f_s["h"], dy, emulate_bf16=False)
    eps = 1e-2
    for (i, j) in [(0, 0), (3, 5), (19, 11)]:
        xp = x.copy(); xp[i, j] += eps
        xm = x.copy(); xm[i, j] -= eps
        fd = (loss(xp) - loss(xm)) / (2 * eps)
        assert abs(fd - b_s["dx"][i, j]) < 5e-2 * max(1.0, abs(fd))


def test_grad_capture_point_is_dense_output():
    # tests/reproduce_grad_sink.py:59-67: the sink is added to the Dense output, so its gradient is dL/dz.
    x, wu, bu, wd, bd = _layer(1)
    a = np.zeros(18, np.int64)
    fw = O.stateful_layer_fwd(x, wu, bu, wd, bd, a, emulate_bf16=False)
    dy = np.ones_like(fw["y"])
    bw = O.stateful_layer_bwd(x, wu, wd, a, fw["z"], fw["h"], dy, emulate_bf16=False)
    eps = 1e-3
    i, j = 2, 3

    def loss(sink):
        z = x @ wu + bu + sink
        h = O.gated_residual(z.astype(np.float32), a)
        return float((h @ wd + bd).sum())

    s = np.zeros_like(fw["z"]); s[i, j] = eps
    fd = (loss(s) - loss(-s)) / (2 * eps)
    assert abs(fd - bw["dz"][i, j]) < 1e-2 * max(1.0, abs(fd))


def test_bf16_rounding_helpers():
    import torch
    v = np.random.RandomState(0).randn(1000).astype(np.float32) * 100
    ref = torch.from_numpy(v).bfloat16().float().numpy()
    np.testing.assert_array_equal(O.bf16_round(v), ref)
    np.testing.assert_array_equal(O.bf16_from_bits(O.bf16_bits(v)), ref)


# ------------------------------------------------------------------ CLIP loss (parity unpinned: self-consistency)
def test_clip_loss_gradients_by_finite_difference():
    rng = np.random.RandomState(3)
    img = rng.randn(6, 10).astype(np.float32); txt = rng.randn(6, 10).astype(np.float32)
    lt = 1.3
    dimg, dtxt, dlt, _ = O.clip_loss_bwd(img, txt, lt)
    f = lambda i, t, l: float(O.clip_loss_fwd(i.astype(np.float64), t.astype(np.float64), l)[0])
    eps = 1e-3
    ip = img.copy(); ip[2, 4] += eps; im = img.copy(); im[2, 4] -= eps
    assert abs((f(ip, txt, lt) - f(im, txt, lt)) / (2 * eps) - dimg[2, 4]) < 2e-3
    tp = txt.copy(); tp[5, 1] += eps; tm = txt.copy(); tm[5, 1] -= eps
    assert abs((f(img, tp, lt) - f(img, tm, lt)) / (2 * eps) - dtxt[5, 1]) < 2e-3
    assert abs((f(img, txt, lt + eps) - f(img, txt, lt - eps)) / (2 * eps) - dlt) < 2e-3


def test_clip_loss_perfect_alignment_is_small():
    e = np.eye(8, 16, dtype=np.float32)
    loss, s, _, _ = O.clip_loss_fwd(e, e, np.log(100.0))
    assert loss < 1e-3 and abs(s[0, 0] - 100.0) < 1e-3


@pytest.mark.parametrize("i", range(4))
def test_moe_routing_math_matches_the_reference_golden(golden_dir, i):
    """tests/golden/moe_layer.npz: outputs of the reference's own MoELayer._top_k_gating / _compute_expert_outputs /
    _load_balancing_loss (src/forde/moe.py:116-222), executed by tests/golden/make_golden.py.  Indices bit-exact."""
    z = np.load(os.path.join(golden_dir, "moe_layer.npz"))
    E, k = (int(v) for v in z[f"cfg{i}"])
    logits = z[f"logits{i}"]
    B, S, _ = logits.shape
    vals, idx = O.top_k(logits.reshape(B * S, E), k)
    np.testing.assert_array_equal(idx, z[f"idx{i}"].reshape(B * S, k))
    w = O.softmax(vals)
    np.testing.assert_allclose(w, z[f"w{i}"].reshape(B * S, k), rtol=1e-6, atol=1e-7)
    probs = O.softmax(logits.reshape(B * S, E))
    np.testing.assert_allclose(probs, z[f"probs{i}"].reshape(B * S, E), rtol=1e-6, atol=1e-7)
    eo = z[f"expert_out{i}"]
    out = O.moe_combine(eo.reshape(E, B * S, -1), idx, w)
    np.testing.assert_allclose(out, z[f"out{i}"].reshape(B * S, -1), rtol=1e-5, atol=1e-6)
    aux, counts = O.moe_aux_loss(probs, idx, E, k, 0.01)
    np.testing.assert_allclose(aux, z[f"aux{i}"], rtol=1e-5)
    assert int(counts.sum()) == B * S * k
    if i == 1:   # the tie row [1,3,3,2,3,0,-1,3]: lower index first
        np.testing.assert_array_equal(idx[0], [1, 2])


@pytest.mark.parametrize("i", range(3))
def test_gmm_clustering_matches_the_reference_call_for_a_given_seed(golden_dir, i):
    """tests/golden/clustering.npz: labels and parameters returned by the reference's own cluster_neurons_gmm
    (src/forde/clustering.py:6-39) for a recorded sklearn seed.  Same scikit-learn, same call -> same labels; if the
    installed scikit-learn differs from the one that wrote the fixture only the partition is compared."""
    import sklearn
    z = np.load(os.path.join(golden_dir, "clustering.npz"))
    stats, k, seed = z[f"stats{i}"], int(z[f"k{i}"]), int(z[f"seed{i}"])
    labels, params = O.cluster_neurons_gmm(stats, k, seed)
    assert labels.dtype == np.int32 and set(params) == {"weights", "means", "covariances"}
    ref = z[f"labels{i}"]
    # the partition must be the reference's (well-separated blobs); label ids and parameters are compared after
    # ordering the components by their means, so a k-means++ draw that differs in the last ulp on another CPU
    # (a relabelling, nothing else) cannot fail the test
    pairs = set(zip(labels.tolist(), ref.tolist()))
    assert len(pairs) == k, "different partition of the neurons"
    if str(z["sklearn_version"]) == sklearn.__version__:
        order = np.lexsort(params["means"].T[::-1])
        ref_order = np.lexsort(z[f"means{i}"].T[::-1])
        np.testing.assert_allclose(params["means"][order], z[f"means{i}"][ref_order], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(params["weights"][order], z[f"weights{i}"][ref_order], rtol=1e-5, atol=1e-7)


def test_optax_clip_adamw_restatement_known_answers():
    """Hand-computed first AdamW step (bias correction makes |update| = lr for every nonzero gradient) and the
    clip_by_global_norm trigger (strictly-below-threshold gradients pass unchanged, no epsilon)."""
    p = np.array([1.0, -2.0, 0.5, 0.0], np.float32)
    g = np.array([3.0, -4.0, 0.0, 0.0], np.float32)               # norm 5 -> clipped to norm 1: (0.6, -0.8, 0, 0)
    z = np.zeros(4, np.float32)
    p1, mu, nu, count, norm = O.optax_clip_adamw_step(p, g, z, z, 0, lr=0.1, wd=0.0)
    assert count == 1 and abs(float(norm) - 5.0) < 1e-6
    np.testing.assert_allclose(mu, 0.1 * np.array([0.6, -0.8, 0, 0]), rtol=1e-6)
    np.testing.assert_allclose(nu, 0.001 * np.array([0.36, 0.64, 0, 0]), rtol=1e-5)
    np.testing.assert_allclose(p1, p - 0.1 * np.array([1, -1, 0, 0]), rtol=1e-5, atol=1e-6)   # mu_hat/sqrt(nu_hat) = sign(g)
    # weight decay is decoupled: with zero gradients the step is p * (1 - lr * wd)
    p2, *_ = O.optax_clip_adamw_step(p, z, z, z, 0, lr=0.1, wd=0.5)
    np.testing.assert_allclose(p2, p * 0.95, rtol=1e-6)
    # below the threshold nothing is scaled
    _, mu3, _, _, n3 = O.optax_clip_adamw_step(p, 0.1 * g, z, z, 0, lr=0.1, wd=0.0)
    np.testing.assert_allclose(mu3, 0.1 * 0.1 * g, rtol=1e-6)
    assert abs(float(n3) - 0.5) < 1e-6
