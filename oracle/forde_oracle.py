"""CPU oracle for the FORDE fast-loop hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import this module, and only as the checker / the timed CPU baseline.  The product path
(``forde-vision-language_b200/``) never imports it and has no CPU fallback.

What is restated, and how firmly it is pinned:

* ``hoyer_sparsity`` / ``calculate_neuron_stats`` follow /root/reference/src/forde/sensing.py:5-69 line by
  line.  PINNED: tests/golden/sensing_*.npz were produced by executing the reference's own sensing.py under
  a NumPy-backed ``jax`` stub (tests/golden/make_golden.py; JAX itself is not installable here), and the
  known answers of tests/verify_hoyer.py:10-22 and sensing.py:86-97 are asserted.
* ``smooth_assignments`` / ``assignments_to_grid`` follow src/forde/smoothing.py:14-91.  PINNED the same way
  (tests/golden/smoothing_*.npz).
* ``cluster_neurons_gmm`` follows src/forde/clustering.py:6-39 (sklearn GaussianMixture; the seed is passed
  as an int because jax.random is unavailable).  PINNED for a given seed: tests/golden/clustering.npz holds the
  partition and GMM parameters the reference's own function returns (scikit-learn is installed, the function runs
  unchanged under the stub); the jax PRNG step that turns a key into that seed is not reproduced.
* ``smooth_assignments_3d`` / ``compute_router_adjustments`` / ``moe_usage_diagnostics`` follow smoothing.py:94-171 and
  src/forde/moe_slow_loop.py:211-250,372-381.  PINNED: tests/golden/moe_slow_loop.npz holds outputs of the reference's
  own functions and of one whole ``moe_slow_loop_step`` on the pytrees of tests/verify_reset.py, executed under the
  same stub (GMM labels excepted: their seed comes from jax.random, which the stub cannot reproduce).
* ``top_k`` / ``moe_*`` follow src/forde/moe.py:88-93,116-135,163-187,189-222.  Index SETS pinned by
  tests/test_moe.py:46-70 and tests/test_topk_correctness.py:7-23; order/ties follow jax.lax.top_k's
  documented rule (descending, lower index first).  PINNED: tests/golden/moe_layer.npz holds outputs of the
  reference's own MoELayer._top_k_gating / _compute_expert_outputs / _load_balancing_loss (renormalised gates, the
  gather + weighted combine, the load-balancing loss) executed under the same stub.
* ``stateful_layer_*`` and ``clip_loss_*`` restate /root/reference/README.md:47-56 and :35 under the
  decisions D1-D9 of SURVEY.md section 8a.  **PARITY UNPINNED**: the reference snapshot contains no
  StatefulLayer / contrastive-loss code and no test vectors for them; the gradient-capture point is pinned
  by tests/reproduce_grad_sink.py:59-67 (dL/dz at the Dense output).

All arithmetic is NumPy fp32 (the reference is fp32) unless a function says otherwise.  ``bf16``
helpers emulate the storage rounding of the CUDA path so discrete decisions (relu mask, binary step, top-k)
can be compared bit-exactly on the same stored values.
"""
from __future__ import annotations

import math

import numpy as np

GATE_GENERALIST = 1.0
GATE_SPECIALIST = 0.1


# ----------------------------------------------------------------------------------------------- bf16
def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 to the nearest bf16 (ties to even) and return it as fp32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def bf16_bits(x: np.ndarray) -> np.ndarray:
    """uint16 bit pattern of bf16_round(x) (for bit-exact comparisons)."""
    return (bf16_round(x).view(np.uint32) >> np.uint32(16)).astype(np.uint16)


def bf16_from_bits(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


# ----------------------------------------------------------------------------------------------- sensing
def hoyer_sparsity(x: np.ndarray) -> np.ndarray:
    """sensing.py:5-32.  Hoyer sparsity along the last axis, with the zero-vector and n==1 guards."""
    x = np.asarray(x, dtype=np.float32)
    n = x.shape[-1]
    l1_norm = np.sum(np.abs(x), axis=-1, dtype=np.float32)
    l2_norm = np.sqrt(np.sum(np.square(x), axis=-1, dtype=np.float32))
    safe_l2 = np.where(l2_norm == 0, np.float32(1.0), l2_norm)                 # :22
    denom = np.float32(1.0) if n == 1 else np.float32(np.sqrt(np.float32(n)) - np.float32(1))  # :25
    sparsity = (np.sqrt(np.float32(n)) - (l1_norm / safe_l2)) / denom          # :27
    sparsity = np.where(l2_norm == 0, np.float32(0.0), sparsity)               # :30
    if n == 1:                                                                 # :32
        return np.zeros_like(sparsity, dtype=np.float32)
    return sparsity.astype(np.float32)


def calculate_neuron_stats(activations: np.ndarray, gradients: np.ndarray) -> np.ndarray:
    """sensing.py:34-69.  (..., F) activations (pre-nonlinearity z) and dL/dz -> (F, 5)
    [grad_gini, grad_gdp, act_gini, act_gdp, act_variance]."""
    f = activations.shape[-1]
    a = np.asarray(activations).reshape(-1, f).astype(np.float32)              # :49-53
    g = np.asarray(gradients).reshape(-1, f).astype(np.float32)
    act_gini = hoyer_sparsity(a.T)                                             # :57 (vmap over neurons)
    act_gdp = np.mean(np.abs(a), axis=0, dtype=np.float32)                     # :58
    act_var = np.var(a, axis=0, dtype=np.float32)                              # :59 (ddof=0, two-pass)
    grad_gini = hoyer_sparsity(g.T)                                            # :62
    grad_gdp = np.mean(np.abs(g), axis=0, dtype=np.float32)                    # :63
    return np.stack([grad_gini, grad_gdp, act_gini, act_gdp, act_var], axis=-1).astype(np.float32)  # :67


def raw_moments(activations: np.ndarray, gradients: np.ndarray):
    """The additive sufficient statistics the CUDA path accumulates (float64 here):
    (sum|z|, sum z, sum z^2, sum|g|, sum g^2) per neuron, plus the sample count."""
    f = activations.shape[-1]
    a = np.asarray(activations, dtype=np.float64).reshape(-1, f)
    g = np.asarray(gradients, dtype=np.float64).reshape(-1, f)
    m = np.stack([np.abs(a).sum(0), a.sum(0), (a * a).sum(0), np.abs(g).sum(0), (g * g).sum(0)], axis=-1)
    return m, a.shape[0]


def stats_from_moments(moments: np.ndarray, n: int) -> np.ndarray:
    """(H,5) stats from pooled raw moments -- algebraically identical to calculate_neuron_stats."""
    m = np.asarray(moments, dtype=np.float64)
    a_abs, a_sum, a_sq, g_abs, g_sq = (m[:, i] for i in range(5))

    def hoyer(l1, l2sq):
        if n == 1:
            return np.zeros_like(l1, dtype=np.float32)
        l2 = np.sqrt(l2sq.astype(np.float32))
        safe = np.where(l2 == 0, np.float32(1), l2)
        sn = np.sqrt(np.float32(n))
        s = (sn - l1.astype(np.float32) / safe) / (sn - np.float32(1))
        return np.where(l2 == 0, np.float32(0), s).astype(np.float32)

    mean = a_sum / n
    var = np.maximum(a_sq / n - mean * mean, 0.0)
    return np.stack([hoyer(g_abs, g_sq), (g_abs / n).astype(np.float32), hoyer(a_abs, a_sq),
                     (a_abs / n).astype(np.float32), var.astype(np.float32)], axis=-1).astype(np.float32)


# ----------------------------------------------------------------------------------------------- StatefulLayer
def binary_step(z: np.ndarray) -> np.ndarray:
    """README.md:54; D3: strict, step(0) = 0."""
    return (z > 0).astype(np.float32)


def multiplex(z: np.ndarray, assign: np.ndarray) -> np.ndarray:
    """README.md:50-54: F = where(a==0, relu(z), where(a==1, tanh(z), binary_step(z))); a broadcasts over the
    last axis.  D5: anything outside {0,1} takes the last branch."""
    a = np.asarray(assign)
    return np.where(a == 0, np.maximum(z, 0), np.where(a == 1, np.tanh(z), binary_step(z))).astype(np.float32)


def gate_of(assign: np.ndarray) -> np.ndarray:
    """README.md:56; D4: specialist (2, and anything outside {0,1}) -> 0.1, generalist/pooling -> 1.0."""
    a = np.asarray(assign)
    return np.where((a == 0) | (a == 1), np.float32(GATE_GENERALIST), np.float32(GATE_SPECIALIST)).astype(np.float32)


def multiplex_grad(z: np.ndarray, assign: np.ndarray) -> np.ndarray:
    """dF/dz with the identity straight-through estimator for binary_step (README.md:55; D3)."""
    a = np.asarray(assign)
    return np.where(a == 0, (z > 0).astype(np.float32),
                    np.where(a == 1, 1 - np.tanh(z) ** 2, np.float32(1.0))).astype(np.float32)


def gated_residual(z: np.ndarray, assign: np.ndarray) -> np.ndarray:
    """h = F(z) + gate * z in the same arithmetic as the CUDA epilogue: one fused multiply-add in fp32
    (emulated exactly through float64, whose product of two fp32 values is exact)."""
    f = multiplex(z, assign).astype(np.float64)
    g = gate_of(assign).astype(np.float64)
    return (g * z.astype(np.float64) + f).astype(np.float32)


def stateful_layer_fwd(x, w_up, b_up, w_down, b_down, assign, emulate_bf16=True):
    """README.md:47-56 under D1-D5.  x (M,d), w_up (d,H) [Flax kernel (in,out)], w_down (H,d).
    With emulate_bf16 the stored tensors z, h, y are rounded to bf16 exactly where the CUDA path rounds them
    (GEMM accumulation stays fp32).  Returns dict(z, h, y) as fp32 arrays."""
    x = np.asarray(x, np.float32); w_up = np.asarray(w_up, np.float32); w_down = np.asarray(w_down, np.float32)
    z = x @ w_up + np.asarray(b_up, np.float32)
    if emulate_bf16:
        z = bf16_round(z)
    h = gated_residual(z, assign)
    if emulate_bf16:
        h = bf16_round(h)
    y = h @ w_down + np.asarray(b_down, np.float32)
    if emulate_bf16:
        y = bf16_round(y)
    return {"z": z, "h": h, "y": y}


def stateful_layer_bwd(x, w_up, w_down, assign, z, h, dy, emulate_bf16=True):
    """Autodiff of stateful_layer_fwd with the STE (SURVEY a-3).  `z`, `h` are the saved forward tensors.
    Returns dict(dz, dx, dw_up, db_up, dw_down, db_down); dz is dL/dz, the tensor the gradient statistics
    are taken of (tests/reproduce_grad_sink.py:16-17)."""
    x = np.asarray(x, np.float32); w_up = np.asarray(w_up, np.float32); w_down = np.asarray(w_down, np.float32)
    dy = np.asarray(dy, np.float32)
    dh = dy @ w_down.T
    dz = dh * (multiplex_grad(z, assign) + gate_of(assign))
    if emulate_bf16:
        dz = bf16_round(dz)
    dx = dz @ w_up.T
    if emulate_bf16:
        dx = bf16_round(dx)
    return {
        "dz": dz, "dx": dx,
        "dw_up": x.T @ dz, "db_up": dz.sum(0),
        "dw_down": h.T @ dy, "db_down": dy.sum(0),
    }


def stateful_layer_fwd_from_z(z_stored: np.ndarray, assign: np.ndarray) -> np.ndarray:
    """h (bf16-rounded) recomputed from a stored bf16 z: the bit-exact check of the multiplex + gate."""
    return bf16_round(gated_residual(np.asarray(z_stored, np.float32), assign))


# ----------------------------------------------------------------------------------------------- top-k
def top_k(x: np.ndarray, k: int):
    """jax.lax.top_k semantics (moe.py:130): values descending, equal values -> lower index first."""
    x = np.asarray(x)
    idx = np.argsort(-x, axis=-1, kind="stable")[..., :k]
    return np.take_along_axis(x, idx, axis=-1), idx.astype(np.int32)


def topk_sparsify(h: np.ndarray, k: int):
    """D8 (SURVEY a-10): per row keep the k largest |h| over the last axis, zero the rest.
    Returns (sparse_h, indices) with indices in lax.top_k order."""
    _, idx = top_k(np.abs(h), k)
    out = np.zeros_like(h)
    np.put_along_axis(out, idx, np.take_along_axis(h, idx, axis=-1), axis=-1)
    return out, idx


# ----------------------------------------------------------------------------------------------- MoE (config 4)
def softmax(x, axis=-1):
    x = np.asarray(x, np.float32)
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return (e / e.sum(axis=axis, keepdims=True)).astype(np.float32)


def gelu_tanh(x):
    """flax.linen.gelu default (approximate=True), moe.py:27."""
    x = np.asarray(x, np.float32)
    return (0.5 * x * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * x ** 3)))).astype(np.float32)


def moe_router(x, w_r, b_r, k):
    """moe.py:88-93,116-135.  x (M,d). Returns probs (M,E), top-k idx (M,k) int32, top-k weights (M,k)."""
    logits = np.asarray(x, np.float32) @ np.asarray(w_r, np.float32) + np.asarray(b_r, np.float32)
    probs = softmax(logits)
    vals, idx = top_k(logits, k)
    return probs, idx, softmax(vals)


def moe_aux_loss(probs, idx, num_experts, k, weight):
    """moe.py:189-222: E * sum_e (bincount_e / (M k)) * mean_m probs_e * weight."""
    counts = np.bincount(idx.reshape(-1), minlength=num_experts)[:num_experts]
    frac = counts.astype(np.float32) / np.float32(probs.shape[0] * k)
    return np.float32(num_experts) * np.sum(frac * probs.mean(0)) * np.float32(weight), counts.astype(np.int32)


def moe_layer_fwd(x, w_r, b_r, experts, k, aux_weight=0.01):
    """moe.py:72-114,137-187: every expert on every token, gather the top-k, weighted sum.
    experts: list of (w_up, b_up, w_down, b_down)."""
    x = np.asarray(x, np.float32)
    probs, idx, w = moe_router(x, w_r, b_r, k)
    outs = np.stack([gelu_tanh(x @ wu + bu) @ wd + bd for (wu, bu, wd, bd) in experts], axis=0)  # (E,M,d)
    out = moe_combine(outs, idx, w)
    aux, counts = moe_aux_loss(probs, idx, len(experts), k, aux_weight)
    return out, aux, probs, idx, counts


def moe_combine(expert_outs, idx, w):
    """moe.py:163-187: out[m] = sum_j w[m, j] * expert_outs[idx[m, j], m].  expert_outs (E, M, d); idx, w (M, k).
    Pinned by tests/golden/moe_layer.npz (the reference's own _compute_expert_outputs)."""
    expert_outs = np.asarray(expert_outs, np.float32)
    out = np.zeros(expert_outs.shape[1:], np.float32)
    rows = np.arange(expert_outs.shape[1])
    for j in range(idx.shape[1]):
        out = out + w[:, j:j + 1] * expert_outs[idx[:, j], rows, :]
    return out


# ----------------------------------------------------------------------------------------------- CLIP loss
def clip_loss_fwd(img, txt, log_temp):
    """README.md:35 under D9: cosine similarity of every (image, text) pair scaled by exp(tau); symmetric
    cross entropy against the diagonal.  Returns (loss, S, u, v)."""
    img = np.asarray(img, np.float32); txt = np.asarray(txt, np.float32)
    u = img / np.linalg.norm(img, axis=-1, keepdims=True)
    v = txt / np.linalg.norm(txt, axis=-1, keepdims=True)
    s = np.float32(np.exp(log_temp)) * (u @ v.T)
    b = s.shape[0]

    def lse(a, axis):
        m = a.max(axis=axis, keepdims=True)
        return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)

    d = np.diag(s)
    loss = 0.5 * ((lse(s, 1) - d).mean() + (lse(s, 0) - d).mean())
    return np.float32(loss), s, u, v


def clip_loss_bwd(img, txt, log_temp):
    """Analytic gradient of clip_loss_fwd w.r.t. img, txt and log_temp."""
    img = np.asarray(img, np.float64); txt = np.asarray(txt, np.float64)
    ni = np.linalg.norm(img, axis=-1, keepdims=True); nt = np.linalg.norm(txt, axis=-1, keepdims=True)
    u = img / ni; v = txt / nt
    scale = np.exp(np.float64(log_temp))
    s = scale * (u @ v.T)
    b = s.shape[0]
    pr = np.exp(s - s.max(1, keepdims=True)); pr /= pr.sum(1, keepdims=True)
    pc = np.exp(s - s.max(0, keepdims=True)); pc /= pc.sum(0, keepdims=True)
    ds = (pr + pc) / (2 * b) - np.eye(b) / b
    du = scale * ds @ v
    dv = scale * ds.T @ u
    dimg = (du - u * (u * du).sum(-1, keepdims=True)) / ni
    dtxt = (dv - v * (v * dv).sum(-1, keepdims=True)) / nt
    dlog_temp = (ds * s).sum()
    return dimg.astype(np.float32), dtxt.astype(np.float32), np.float32(dlog_temp), ds.astype(np.float32)


# ----------------------------------------------------------------------------------------------- optimizer
def optax_clip_adamw_step(params, grads, mu, nu, count, lr, wd, max_norm=1.0, b1=0.9, b2=0.999, eps=1e-8):
    """One update of optax.chain(optax.clip_by_global_norm(max_norm), optax.adamw(lr, weight_decay=wd)), the
    optimizer the reference builds at src/training/train.py:74-77 (optax 0.2.6 is not installable here; this restates
    its published update rules in fp32): clip_by_global_norm scales by max_norm / norm only when norm >= max_norm (no
    epsilon); scale_by_adam keeps mu, nu, count with bias correction 1 - b^count and eps OUTSIDE the square root
    (eps_root = 0); add_decayed_weights(wd) adds wd * p; the update is -lr times that.
    Returns (params, mu, nu, count, global_norm) -- global_norm is the pre-clip norm train.py:168 logs."""
    p = np.asarray(params, np.float32); g = np.asarray(grads, np.float32)
    norm = np.float32(np.sqrt(np.sum(g.astype(np.float64) ** 2)))
    if norm >= np.float32(max_norm):
        g = (g / norm) * np.float32(max_norm)
    count = int(count) + 1
    mu = (np.float32(b1) * mu + np.float32(1 - b1) * g).astype(np.float32)
    nu = (np.float32(b2) * nu + np.float32(1 - b2) * g * g).astype(np.float32)
    mu_hat = mu / np.float32(1 - b1 ** count)
    nu_hat = nu / np.float32(1 - b2 ** count)
    upd = mu_hat / (np.sqrt(nu_hat) + np.float32(eps)) + np.float32(wd) * p
    return (p - np.float32(lr) * upd).astype(np.float32), mu, nu, count, norm


# ----------------------------------------------------------------------------------------------- slow loop
def assignments_to_grid(assignments, grid_size):
    """smoothing.py:14-27."""
    return np.asarray(assignments).reshape(grid_size)


def smooth_assignments(assignment_grid, kernel_size=3, num_clusters=3):
    """smoothing.py:30-91: one-hot -> per-cluster box filter (zero-filled 'same' convolution) -> argmax
    (first maximum wins, i.e. ties go to the lowest cluster id)."""
    from scipy.signal import convolve2d

    grid = np.asarray(assignment_grid)
    kernel = np.ones((kernel_size, kernel_size), np.float32) / np.float32(kernel_size ** 2)     # :49
    one_hot = (grid[..., None] == np.arange(num_clusters)).astype(np.float32)                    # :52
    oh, ow, _ = one_hot.shape
    pad_h = max(0, kernel_size + 1 - oh)                                                         # :59-60
    pad_w = max(0, kernel_size + 1 - ow)
    cfg = ((pad_h // 2, pad_h - pad_h // 2), (pad_w // 2, pad_w - pad_w // 2), (0, 0))           # :64-68
    padded = np.pad(one_hot, cfg, "constant")
    sm = np.stack([convolve2d(padded[:, :, i], kernel, mode="same") for i in range(num_clusters)], axis=-1)  # :72-78
    sm = sm[cfg[0][0]:cfg[0][0] + oh, cfg[1][0]:cfg[1][0] + ow, :]                               # :81-85
    return np.argmax(sm, axis=-1)                                                                # :89


def cluster_neurons_gmm(aggregated_stats, num_clusters, seed: int):
    """clustering.py:6-39 with the jax PRNG-derived seed passed in as an int."""
    from sklearn.mixture import GaussianMixture

    stats = np.asarray(aggregated_stats)
    gmm = GaussianMixture(n_components=num_clusters, random_state=int(seed))                    # :27
    gmm.fit(stats)                                                                               # :28
    assignments = gmm.predict(stats)                                                             # :30
    params = {"weights": gmm.weights_, "means": gmm.means_, "covariances": gmm.covariances_}    # :33-37
    return assignments.astype(np.int32), params


def forde_lite_assignments(stats, threshold=0.8):
    """README.md:58-59 ablation rule: is_spec = grad_gini > 0.8 (specialist = 2, else generalist = 0)."""
    return np.where(np.asarray(stats)[:, 0] > threshold, 2, 0).astype(np.int32)


# ----------------------------------------------------------------------------------------------- MoE fwd+bwd (torch)
def moe_layer_torch(x, w_r, b_r, w_up, b_up, w_down, b_down, k, aux_weight, dout=None, d_aux=1.0):
    """The reference formulation (moe.py:72-114,137-222) in float64 torch with autograd: EVERY expert on every token,
    gather the top-k, weighted sum, Switch-style aux loss.  w_up (E,d,H), w_down (E,H,d) stacked.
    Returns dict(out, aux, probs, idx, counts[, dx, dw_r, db_r, dw_up, db_up, dw_down, db_down])."""
    import torch

    t = lambda a: torch.tensor(np.asarray(a, np.float64), dtype=torch.float64, requires_grad=True)
    x_, wr_, br_, wu_, bu_, wd_, bd_ = (t(a) for a in (x, w_r, b_r, w_up, b_up, w_down, b_down))
    M, E = x_.shape[0], wr_.shape[1]
    logits = x_ @ wr_ + br_                                                    # moe.py:88
    probs = torch.softmax(logits, -1)                                          # moe.py:90
    order = torch.sort(-logits.detach(), dim=-1, stable=True).indices[:, :k]   # lax.top_k order (moe.py:130)
    top_logits = torch.gather(logits, 1, order)
    w = torch.softmax(top_logits, -1)                                          # moe.py:133

    def gelu(v):
        return 0.5 * v * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (v + 0.044715 * v ** 3)))

    all_out = torch.stack([gelu(x_ @ wu_[e] + bu_[e]) @ wd_[e] + bd_[e] for e in range(E)], 0)   # moe.py:163
    rows = torch.arange(M)
    out = sum(w[:, j:j + 1] * all_out[order[:, j], rows, :] for j in range(k))                  # moe.py:169-185
    counts = torch.bincount(order.reshape(-1), minlength=E)                                       # moe.py:216
    aux = E * torch.sum(counts.double() / (M * k) * probs.mean(0)) * aux_weight                   # moe.py:217-222
    res = {"out": out.detach().numpy(), "aux": float(aux.detach()), "probs": probs.detach().numpy(),
           "idx": order.numpy().astype(np.int32), "counts": counts.numpy().astype(np.int32),
           "logits": logits.detach().numpy()}
    if dout is not None:
        loss = (out * torch.tensor(np.asarray(dout, np.float64))).sum() + d_aux * aux
        loss.backward()
        for name, v in (("dx", x_), ("dw_r", wr_), ("db_r", br_), ("dw_up", wu_), ("db_up", bu_), ("dw_down", wd_),
                        ("db_down", bd_)):
            res[name] = v.grad.numpy()
    return res


# ----------------------------------------------------------------------------------------------- MoE slow loop
def smooth_assignments_3d(assignment_grid, kernel_size=3, num_clusters=3):
    """smoothing.py:94-171: one-hot -> per-cluster k^3 box filter ('same' convolution; grids smaller than the kernel
    are edge-padded first and cropped back) -> argmax (first maximum wins)."""
    from scipy.signal import convolve

    grid = np.asarray(assignment_grid)
    kernel = np.ones((kernel_size,) * 3, np.float32) / np.float32(kernel_size ** 3)            # :114
    one_hot = (grid[..., None] == np.arange(num_clusters)).astype(np.float32)                  # :117
    d, h, w, _ = one_hot.shape
    chans = []
    for c in range(num_clusters):
        ch = one_hot[..., c]
        pads = [max(0, kernel_size - n) for n in ch.shape]                                     # :133-135
        padding = tuple((p // 2, p - p // 2) for p in pads)                                    # :139-143
        padded = np.pad(ch, padding, "edge") if any(pads) else ch                              # :144-146
        sm = convolve(padded, kernel, mode="same").astype(np.float32)                          # :152
        if any(pads):
            sm = sm[padding[0][0]:padding[0][0] + d, padding[1][0]:padding[1][0] + h, padding[2][0]:padding[2][0] + w]
        chans.append(sm)
    return np.argmax(np.stack(chans, -1), axis=-1)                                             # :166-169


def compute_router_adjustments(expert_assignments, expert_usage, target_balance=0.1):
    """moe_slow_loop.py:211-250: bias nudge = clip(-(usage - 1/E) * target_balance, +-0.1), halved for cluster 1."""
    usage = np.asarray(expert_usage, np.float32)
    uniform = np.float32(1.0 / usage.shape[0])
    adj = np.clip(-(usage - uniform) * np.float32(target_balance), -0.1, 0.1).astype(np.float32)
    return np.where(np.asarray(expert_assignments) == 1, adj * np.float32(0.5), adj).astype(np.float32)


def moe_usage_diagnostics(usage_mean):
    """moe_slow_loop.py:372-381: mean |usage - 1/E| and routing entropy relative to log E."""
    u = np.asarray(usage_mean, np.float32)
    e = u.shape[0]
    imbalance = np.abs(u - np.float32(1.0 / e)).mean()
    entropy = -(u * np.log(u + np.float32(1e-8))).sum() / np.log(np.float32(e))
    return np.float32(imbalance), np.float32(entropy)
