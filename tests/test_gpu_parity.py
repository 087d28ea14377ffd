"""Parity tests proper: the sm_100a kernels, called through the C ABI (ctypes), against the CPU oracle on the
same seeded inputs.  Tolerances (none exist upstream):
  z, h, y, dz, dx vs the fp32-accumulating oracle with bf16 storage emulated: relative Frobenius <= 1e-2 and
      EVERY element within 2^-7 |ref| (one bf16 ulp at most) + 1e-3 * max(1, rms(ref)) of the oracle
      (assert_close_bf16; no element is forgiven);
  dW, db (fp32 outputs): relative Frobenius <= 1e-3;
  statistics (H,5): rtol 1e-3, atol 1e-5, computed by the oracle from the kernel's own stored z / dz;
  discrete outputs (relu mask, binary step, gate selection): BIT-exact given the same stored bf16 z.
"""
import numpy as np
import pytest
import torch

from oracle import forde_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import __graft_entry__ as g
    g.build()
    from forde_b200 import ops as _ops
    return _ops


def f32(t):
    return t.detach().float().cpu().numpy()


def relerr(got, ref):
    return float(np.linalg.norm(got.astype(np.float64) - ref) / (np.linalg.norm(ref) + 1e-30))


def assert_close_bf16(got, ref, name, fro=1e-2):
    assert relerr(got, ref) <= fro, f"{name}: rel Frobenius {relerr(got, ref):.3e}"
    scale = max(1.0, float(np.sqrt(np.mean(np.square(ref, dtype=np.float64)))))
    bound = 2.0 ** -7 * np.abs(ref) + 1e-3 * scale
    err = np.abs(got - ref)
    bad = err > bound
    assert not bad.any(), (f"{name}: {bad.sum()} of {bad.size} elements outside 2^-7 |ref| + 1e-3 * {scale:.2f}; "
                           f"worst excess {float((err - bound).max()):.3e}")


# ------------------------------------------------------------------------------------------ generic GEMM
@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("shape", [(256, 512, 256), (200, 136, 328), (128, 8, 64), (1000, 264, 72)])
def test_gemm_bf16_all_operand_layouts(ops, a_mn, b_mn, shape):
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda", generator=g).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda", generator=g).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = ops.gemm_bf16(A, B, a_mn, b_mn, bias_col=bias)
    Af = f32(A.t() if a_mn else A); Bf = f32(B.t() if b_mn else B)
    ref = O.bf16_round(Af @ Bf.T + f32(bias))
    assert_close_bf16(f32(out), ref, "gemm_bf16")


@pytest.mark.parametrize("accumulate", [False, True])
def test_gemm_f32_split_k_accumulate(ops, accumulate):
    M, N, K = 384, 520, 4096
    g = torch.Generator(device="cuda").manual_seed(2)
    A = torch.randn((K, M), device="cuda", generator=g).bfloat16()
    B = torch.randn((N, K), device="cuda", generator=g).bfloat16()
    out = torch.full((M, N), 1.0 if accumulate else 7.0, device="cuda")
    ops.gemm_f32(A, B, True, False, out=out, accumulate=accumulate)
    ref = f32(A).T.astype(np.float64) @ f32(B).T.astype(np.float64) + (1.0 if accumulate else 0.0)
    assert relerr(f32(out), ref) < 1e-5


# ------------------------------------------------------------------------------------------ StatefulLayer
def run_layer(ops, M, d, H, assign, seed=0, bias_scale=0.1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = {}
    t["x"] = torch.randn((M, d), device="cuda", generator=g).bfloat16()
    t["w_up"] = (torch.randn((d, H), device="cuda", generator=g) / d ** 0.5).bfloat16()
    t["w_down"] = (torch.randn((H, d), device="cuda", generator=g) / H ** 0.5).bfloat16()
    t["b_up"] = torch.randn(H, device="cuda", generator=g) * bias_scale
    t["b_down"] = torch.randn(d, device="cuda", generator=g) * bias_scale
    t["dy"] = torch.randn((M, d), device="cuda", generator=g).bfloat16()
    t["assign"] = torch.as_tensor(assign, device="cuda").to(torch.int8)
    ld, P = ops.sample_pitch(M), ops.num_partials(M)
    t["zT"] = torch.zeros((H, ld), device="cuda", dtype=torch.bfloat16)
    t["hT"] = torch.zeros_like(t["zT"]); t["dzT"] = torch.zeros_like(t["zT"])
    t["actp"] = torch.zeros((P, 3, H), device="cuda"); t["gradp"] = torch.zeros((P, 3, H), device="cuda")
    t["y"] = torch.zeros((M, d), device="cuda", dtype=torch.bfloat16); t["dx"] = torch.zeros_like(t["y"])
    t["dw_up"] = torch.zeros((d, H), device="cuda"); t["dw_down"] = torch.zeros((H, d), device="cuda")
    t["db_up"] = torch.zeros(H, device="cuda"); t["db_down"] = torch.zeros(d, device="cuda")
    t["stats"] = torch.zeros((H, 5), device="cuda")
    ops.sl_up_fwd(t["x"], t["w_up"], t["b_up"], t["assign"], t["zT"], t["hT"], t["actp"])
    ops.sl_down_fwd(t["hT"], t["w_down"], t["b_down"], t["y"], M)
    ops.sl_dz_bwd(t["dy"], t["w_down"], t["zT"], t["hT"], t["assign"], t["dzT"], t["gradp"])
    ops.sl_dx_bwd(t["dzT"], t["w_up"], t["dx"], M)
    ops.sl_dw_up(t["x"], t["dzT"], t["dw_up"])
    ops.sl_dw_down(t["hT"], t["dy"], t["dw_down"])
    ops.colsum_bf16(t["dy"], t["db_down"])
    ops.stats_finalize(t["actp"], t["gradp"], P, H, M, step_stats=t["stats"], db_up=t["db_up"])
    torch.cuda.synchronize()
    t["z"] = t["zT"][:, :M].t(); t["h"] = t["hT"][:, :M].t(); t["dz"] = t["dzT"][:, :M].t()
    return t


LAYER_SHAPES = [
    (160, 128, 512),     # C1 vision: 32 images x 5 tokens
    (512, 128, 512),     # C1 text:   32 captions x 16 tokens
    (1000, 256, 640),    # ragged: M, H not multiples of the tile
    (7, 64, 136),        # fewer samples than one tile row, H = 128 + 8
    (1, 128, 128),       # a single sample (Hoyer's n == 1 guard)
    (2048, 768, 3072),   # C2 vision dims at a reduced batch
]


@pytest.mark.parametrize("shape", LAYER_SHAPES)
@pytest.mark.parametrize("pattern", ["mixed", "all0", "all1", "all2"])
def test_stateful_layer_forward_backward_parity(ops, shape, pattern):
    M, d, H = shape
    if pattern != "mixed" and shape not in (LAYER_SHAPES[0], LAYER_SHAPES[2]):
        pytest.skip("uniform-assignment extremes are covered on two shapes")
    rng = np.random.RandomState(3)
    assign = {"mixed": rng.randint(0, 3, H), "all0": np.zeros(H, int), "all1": np.ones(H, int),
              "all2": np.full(H, 2)}[pattern]
    t = run_layer(ops, M, d, H, assign)
    x, wu, wd = f32(t["x"]), f32(t["w_up"]), f32(t["w_down"])
    bu, bd, dy = f32(t["b_up"]), f32(t["b_down"]), f32(t["dy"])
    fw = O.stateful_layer_fwd(x, wu, bu, wd, bd, assign)
    assert_close_bf16(f32(t["z"]), fw["z"], "z")
    # everything downstream of z is checked from the kernel's OWN stored z (threshold consistency)
    z_k = f32(t["z"])
    h_ref = O.stateful_layer_fwd_from_z(z_k, assign)
    h_k = f32(t["h"])
    # bit-exact for relu and binary-step neurons (mask, step and gate selection are discrete decisions)
    for a_val in (0, 2):
        m = assign == a_val
        if m.any():
            np.testing.assert_array_equal(O.bf16_bits(h_k[:, m]), O.bf16_bits(h_ref[:, m]))
    m1 = assign == 1
    if m1.any():  # tanh.approx.f32: ~2^-11 relative error before the bf16 rounding -> at most one bf16 ulp
        assert np.abs(h_k[:, m1] - h_ref[:, m1]).max() <= 2.0 ** -7 * np.abs(h_ref[:, m1]).max() + 1e-6
    y_ref = O.bf16_round(h_k @ wd + bd)
    assert_close_bf16(f32(t["y"]), y_ref, "y")
    bw = O.stateful_layer_bwd(x, wu, wd, assign, z_k, h_k, dy)
    assert_close_bf16(f32(t["dz"]), bw["dz"], "dz")
    # STE: specialist neurons see dz = 1.1 * dh regardless of z (binary_step's straight-through estimator)
    dz_k = f32(t["dz"])
    m2 = assign == 2
    if m2.any():
        dh = dy @ wd.T
        assert_close_bf16(dz_k[:, m2], O.bf16_round(np.float32(1.1) * dh[:, m2]), "dz(STE)")
    assert_close_bf16(f32(t["dx"]), O.bf16_round(dz_k @ wu.T), "dx")
    assert relerr(f32(t["dw_up"]), x.T.astype(np.float64) @ dz_k) < 1e-3
    assert relerr(f32(t["dw_down"]), h_k.T.astype(np.float64) @ dy) < 1e-3
    np.testing.assert_allclose(f32(t["db_down"]), dy.sum(0), rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(f32(t["db_up"]), dz_k.astype(np.float64).sum(0), rtol=1e-3, atol=1e-3)
    # sensing: the in-kernel statistics equal the reference's separate pass over (z, dL/dz)
    stats_ref = O.calculate_neuron_stats(z_k, dz_k)
    np.testing.assert_allclose(f32(t["stats"]), stats_ref, rtol=1e-3, atol=1e-5)


def test_padding_columns_do_not_leak_into_statistics(ops):
    # M = 1000 leaves 24 out-of-range sample columns in the last 256-wide tile: they must not be counted.
    M, d, H = 1000, 128, 256
    t = run_layer(ops, M, d, H, np.zeros(H, int), bias_scale=5.0)  # a large bias makes a leak obvious
    z_k = f32(t["z"])
    np.testing.assert_allclose(f32(t["actp"]).sum(0)[1], z_k.astype(np.float64).sum(0), rtol=1e-4, atol=1e-2)


def test_statistics_are_deterministic(ops):
    a = np.random.RandomState(0).randint(0, 3, 512)
    s1 = f32(run_layer(ops, 3000, 128, 512, a)["stats"])
    s2 = f32(run_layer(ops, 3000, 128, 512, a)["stats"])
    np.testing.assert_array_equal(s1, s2)  # one writer per partial slot + fixed-order finalisation


def test_moments_are_additive_over_sample_blocks(ops):
    # the property that lets steps and ranks be pooled: moments(M) == moments(first half) + moments(second half)
    M, d, H = 1024, 128, 256
    a = np.random.RandomState(1).randint(0, 3, H)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((M, d), device="cuda", generator=g).bfloat16()
    w = (torch.randn((d, H), device="cuda", generator=g) / d ** 0.5).bfloat16()
    b = torch.zeros(H, device="cuda")
    ai8 = torch.as_tensor(a, device="cuda").to(torch.int8)

    def moments(xs):
        m = xs.shape[0]
        ld, P = ops.sample_pitch(m), ops.num_partials(m)
        zT = torch.zeros((H, ld), device="cuda", dtype=torch.bfloat16); hT = torch.zeros_like(zT)
        p = torch.zeros((P, 3, H), device="cuda")
        ops.sl_up_fwd(xs, w, b, ai8, zT, hT, p)
        mom = torch.zeros((H, 5), device="cuda"); cnt = torch.zeros((), device="cuda", dtype=torch.int32)
        ops.stats_finalize(p, None, P, H, m, moments=mom, sample_count=cnt)
        torch.cuda.synchronize()
        return f32(mom), int(cnt.item())

    m_all, n_all = moments(x)
    m_a, n_a = moments(x[:512].contiguous()); m_b, n_b = moments(x[512:].contiguous())
    assert n_all == n_a + n_b == M
    np.testing.assert_allclose(m_all, m_a + m_b, rtol=1e-5, atol=1e-4)


# ------------------------------------------------------------------------------------------ module API
def test_module_contract_and_in_place_statistics(ops):
    from forde_b200.stateful_layer import StatefulLayer, STAT_NAMES
    assert STAT_NAMES == ("grad_gini", "grad_gdp", "act_gini", "act_gdp", "act_variance")
    dev = torch.device("cuda")
    layer = StatefulLayer(hidden_dim=512)
    x = torch.randn((4, 40, 128), device=dev).bfloat16().requires_grad_(True)
    v = layer.init(0, x)
    assert set(v) == {"params", "neuron_assignments", "stats_buffer"}
    assert set(v["params"]) == {"up_proj", "down_proj"} and set(v["params"]["up_proj"]) == {"kernel", "bias"}
    assert v["params"]["up_proj"]["kernel"].shape == (128, 512)       # Flax Dense kernel is (in, out)
    assert v["neuron_assignments"]["assignments"].dtype == torch.int32
    assert int(v["neuron_assignments"]["assignments"].abs().sum()) == 0  # D7: all generalists at init
    assert set(v["stats_buffer"]) == {"neuron_stats", "moments", "moments_lo", "sample_count", "step_count"}
    for grp in v["params"].values():
        for p in grp.values():
            p.requires_grad_(True)
    y0 = layer.apply(v, x)                      # immutable call: just the output
    assert y0.shape == x.shape
    for step in range(2):
        y, upd = layer.apply(v, x, mutable=["stats_buffer"])
        assert set(upd) == {"stats_buffer"}
        y.backward(torch.ones_like(y))
    torch.cuda.synchronize()
    sb = upd["stats_buffer"]
    assert int(sb["step_count"].item()) == 2 and int(sb["sample_count"].item()) == 2 * 160
    assert v["params"]["up_proj"]["kernel"].grad.shape == (128, 512)
    assert x.grad is not None and x.grad.shape == x.shape
    # two identical steps: accumulated per-step stats are exactly twice one step's
    one = f32(sb["neuron_stats"]) / 2
    z = f32(x).reshape(-1, 128) @ f32(v["params"]["up_proj"]["kernel"].bfloat16())
    np.testing.assert_allclose(one[:, 3], np.abs(O.bf16_round(z)).mean(0), rtol=1e-3, atol=1e-5)
    # the slow loop rewrites assignments in place -> the next forward routes differently
    v["neuron_assignments"]["assignments"].fill_(2)
    y2 = layer.apply(v, x)
    torch.cuda.synchronize()
    assert (y2.float() - y0.float()).abs().max() > 1e-2


def test_autograd_matches_oracle_backward(ops):
    from forde_b200.stateful_layer import StatefulLayer
    dev = torch.device("cuda")
    M, d, H = 300, 64, 256
    layer = StatefulLayer(hidden_dim=H)
    x = torch.randn((M, d), device=dev).bfloat16().requires_grad_(True)
    v = layer.init(1, x)
    a = np.arange(H) % 3
    v["neuron_assignments"]["assignments"].copy_(torch.as_tensor(a))
    ps = [v["params"]["up_proj"]["kernel"], v["params"]["up_proj"]["bias"], v["params"]["down_proj"]["kernel"],
          v["params"]["down_proj"]["bias"]]
    ps[1].normal_(0, 0.2); ps[3].normal_(0, 0.2)
    for p in ps:
        p.requires_grad_(True)
    y = layer.apply(v, x)
    dy = torch.randn_like(y)
    y.backward(dy)
    torch.cuda.synchronize()
    wu, wd = f32(ps[0].bfloat16()), f32(ps[2].bfloat16())
    fw = O.stateful_layer_fwd(f32(x), wu, f32(ps[1]), wd, f32(ps[3]), a)
    bw = O.stateful_layer_bwd(f32(x), wu, wd, a, fw["z"], fw["h"], f32(dy))
    assert relerr(f32(y), fw["y"]) < 1e-2
    assert relerr(f32(x.grad), bw["dx"]) < 1e-2
    assert relerr(f32(ps[0].grad), bw["dw_up"]) < 1e-2
    assert relerr(f32(ps[2].grad), bw["dw_down"]) < 1e-2
    assert relerr(f32(ps[1].grad), bw["db_up"]) < 1e-2
    assert relerr(f32(ps[3].grad), bw["db_down"]) < 1e-2


# ------------------------------------------------------------------------------------------ CLIP loss
@pytest.mark.parametrize("B,E", [(32, 128), (256, 512), (100, 72)])
def test_clip_loss_forward_backward_parity(ops, B, E):
    from forde_b200.clip_loss import clip_contrastive_loss
    g = torch.Generator(device="cuda").manual_seed(B)
    img = torch.randn((B, E), device="cuda", generator=g).bfloat16().float().requires_grad_(True)
    txt = (torch.randn((B, E), device="cuda", generator=g) + 0.5 * img.detach()).bfloat16().float().requires_grad_(True)
    lt = torch.tensor(float(np.log(1 / 0.07)), device="cuda", requires_grad=True)
    loss = clip_contrastive_loss(img, txt, lt)
    loss.backward()
    torch.cuda.synchronize()
    ref_loss = float(O.clip_loss_fwd(f32(img), f32(txt), float(lt.item()))[0])
    assert abs(float(loss.item()) - ref_loss) <= 1e-2 * max(1.0, abs(ref_loss))   # u, v are stored in bf16
    dimg, dtxt, dlt, _ = O.clip_loss_bwd(f32(img), f32(txt), float(lt.item()))
    assert relerr(f32(img.grad), dimg) < 3e-2
    assert relerr(f32(txt.grad), dtxt) < 3e-2
    assert abs(float(lt.grad.item()) - float(dlt)) < 3e-2 * max(1.0, abs(float(dlt)))


# ------------------------------------------------------------------------------------------ full-size properties
def test_c2_vision_layer_full_size_properties(ops):
    """BASELINE config 2, vision tower layer: M = 256 x 197 = 50432, d = 768, H = 3072 (too big for the NumPy
    oracle to be pleasant: checked through size-independent properties, with torch on the GPU as the checker)."""
    M, d, H = 50432, 768, 3072
    a = np.random.RandomState(0).randint(0, 3, H)
    t = run_layer(ops, M, d, H, a)
    at = torch.as_tensor(a, device="cuda")
    z = t["z"].float()
    z_ref = (t["x"].float() @ t["w_up"].float() + t["b_up"]).bfloat16().float()
    assert ((z - z_ref).norm() / z_ref.norm()).item() < 5e-3
    # multiplex + gate recomputed from the stored z: bit-exact on relu / step neurons
    gate = torch.where(at == 2, 0.1, 1.0).float()
    F = torch.where(at == 0, torch.relu(z), torch.where(at == 1, torch.tanh(z), (z > 0).float()))
    h_ref = (gate.double() * z.double() + F.double()).float().bfloat16()
    h = t["h"]
    exact = (h == h_ref)
    assert bool(exact[:, at != 1].all())
    assert (h.float() - h_ref.float())[:, at == 1].abs().max().item() <= 2.0 ** -7 * 2.0
    # bias gradient == column sums of dz ; dx linear in dz
    dz = t["dz"].float()
    torch.testing.assert_close(t["db_up"], dz.sum(0), rtol=2e-3, atol=2e-2)
    dx_ref = (dz @ t["w_up"].float().t()).bfloat16().float()
    assert ((t["dx"].float() - dx_ref).norm() / dx_ref.norm()).item() < 5e-3
    dw_ref = t["x"].float().t() @ dz
    assert ((t["dw_up"] - dw_ref).norm() / dw_ref.norm()).item() < 2e-3
    # statistics: reference formulae evaluated with torch on the same stored tensors
    n = float(M)
    l1, l2 = z.abs().sum(0), z.pow(2).sum(0).sqrt()
    act_gini = (n ** 0.5 - l1 / l2) / (n ** 0.5 - 1)
    st = t["stats"]
    torch.testing.assert_close(st[:, 2], act_gini, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(st[:, 3], l1 / n, rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(st[:, 4], z.var(0, unbiased=False), rtol=2e-3, atol=1e-5)
    gl1, gl2 = dz.abs().sum(0), dz.pow(2).sum(0).sqrt()
    torch.testing.assert_close(st[:, 0], (n ** 0.5 - gl1 / gl2) / (n ** 0.5 - 1), rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(st[:, 1], gl1 / n, rtol=1e-3, atol=1e-6)


# ------------------------------------------------------------------------------------------ add + LayerNorm
@pytest.mark.parametrize("rows,cols", [(160, 128), (1000, 768), (77, 512), (3, 1024), (50432, 768)])
def test_add_layernorm_forward_backward(ops, rows, cols):
    """The block's add/norm (README.md:42-43): fp32 torch LayerNorm of the fp32 sum as the checker.
    Tolerance: outputs / ds are bf16 (2^-8 relative rounding); d_gamma / d_beta fp32 sums, rtol 2e-3."""
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = torch.randn((rows, cols), device="cuda", generator=g).bfloat16()
    r = (torch.randn((rows, cols), device="cuda", generator=g) * 0.5 + 0.3).bfloat16()
    gamma = torch.rand(cols, device="cuda", generator=g) + 0.5
    beta = torch.randn(cols, device="cuda", generator=g) * 0.1
    dy = torch.randn((rows, cols), device="cuda", generator=g).bfloat16()
    out, mean, rstd = ops.add_layernorm_fwd(x, r, gamma, beta, 1e-6)
    ds, dg, db = ops.add_layernorm_bwd(dy, x, r, gamma, mean, rstd)
    torch.cuda.synchronize()
    s = (x.float() + r.float()).requires_grad_(True)
    gm = gamma.clone().requires_grad_(True); bt = beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(s, (cols,), gm, bt, 1e-6)
    ref.backward(dy.float())
    assert_close_bf16(f32(out), f32(ref), "ln out", fro=5e-3)
    assert_close_bf16(f32(ds), f32(s.grad), "ln ds", fro=5e-3)
    np.testing.assert_allclose(f32(dg), f32(gm.grad), rtol=2e-3, atol=2e-3 * float(gm.grad.abs().mean()) + 1e-4)
    np.testing.assert_allclose(f32(db), f32(bt.grad), rtol=2e-3, atol=2e-3 * float(bt.grad.abs().mean()) + 1e-4)
    np.testing.assert_allclose(f32(mean), f32(s.detach().mean(-1)), rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------ top-k sparsification (D8)
@pytest.mark.parametrize("M,H,k", [(100, 256, 16), (513, 3072, 64), (40, 128, 128), (33, 2048, 1)])
def test_topk_sparsify_matches_lax_top_k_semantics(ops, M, H, k):
    """Indices bit-exact against the oracle's lax.top_k restatement (incl. ties: bf16 values collide often)."""
    g = torch.Generator(device="cuda").manual_seed(M)
    ld = ops.sample_pitch(M)
    hT = torch.zeros((H, ld), device="cuda", dtype=torch.bfloat16)
    hT[:, :M] = (torch.randn((H, M), device="cuda", generator=g) * 2).round().div(2).bfloat16()   # many exact ties
    dense = f32(hT[:, :M].t())                                                                    # (M, H)
    idx, hoyer, keep = ops.sl_topk_sparsify(hT, M, k)
    torch.cuda.synchronize()
    ref_sparse, ref_idx = O.topk_sparsify(dense, k)
    np.testing.assert_array_equal(idx.cpu().numpy(), ref_idx)
    np.testing.assert_array_equal(f32(hT[:, :M].t()), ref_sparse)
    np.testing.assert_allclose(f32(hoyer), O.hoyer_sparsity(dense), rtol=1e-4, atol=1e-6)
    kb = keep.cpu().numpy().view(np.uint32)
    kept = ((kb[np.arange(H)[:, None] >> 5, np.arange(M)[None, :]] >> (np.arange(H)[:, None] & 31)) & 1).astype(bool)
    ref_kept = np.zeros((M, H), bool); np.put_along_axis(ref_kept, ref_idx, True, axis=1)
    np.testing.assert_array_equal(kept.T, ref_kept)


@pytest.mark.parametrize("use_signs", [False, True])
@pytest.mark.parametrize("shape", [(300, 64, 256, 32), (1000, 256, 640, 64), (2048, 512, 2048, 128)])
def test_dz_epilogue_applies_the_topk_mask_bit_exactly(ops, shape, use_signs):
    """forde_sl_dz_bwd_topk: dL/dz == (dL/dz of the dense backward) where the forward's top-k kept h, exactly 0
    elsewhere -- bit for bit -- and the gradient partial sums are those of the masked tensor (no separate masking pass)."""
    M, d, H, k = shape
    a = np.random.RandomState(5).randint(0, 3, H)
    t = run_layer(ops, M, d, H, a, seed=8)                       # dense forward + backward, every z box stored
    hT = t["hT"].clone()
    idx, hoyer, keep = ops.sl_topk_sparsify(hT, M, k)
    kb = keep.cpu().numpy().view(np.uint32)
    kept = ((kb[np.arange(H)[:, None] >> 5, np.arange(M)[None, :]] >> (np.arange(H)[:, None] & 31)) & 1).astype(bool)  # (H, M)
    assert (kept.sum(0) == k).all()
    np.testing.assert_array_equal(f32(hT[:, :M]), f32(t["hT"][:, :M]) * kept)     # masked tile written back whole
    signs = None
    zT = t["zT"]
    if use_signs:
        signs = torch.zeros(ops.sign_shape(M, H), device="cuda", dtype=torch.int32)
        zT2 = torch.full_like(t["zT"], float("nan")); h2 = torch.zeros_like(hT); ap = torch.zeros_like(t["actp"])
        ops.sl_up_fwd(t["x"], t["w_up"], t["b_up"], t["assign"], zT2, h2, ap, save_z=False, sign_bits=signs)
        zT = zT2
    dzT = torch.zeros_like(t["dzT"]); gp = torch.zeros_like(t["gradp"])
    ops.sl_dz_bwd(t["dy"], t["w_down"], zT, hT, t["assign"], dzT, gp, sign_bits=signs, keep_bits=keep)
    torch.cuda.synchronize()
    want = f32(t["dzT"][:, :M]) * kept
    np.testing.assert_array_equal(f32(dzT[:, :M]), want)
    sums = f32(gp).sum(0)                                                          # (3, H): sum|dz|, sum dz, sum dz^2
    np.testing.assert_allclose(sums[0], np.abs(want).astype(np.float64).sum(1), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(sums[1], want.astype(np.float64).sum(1), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(sums[2], (want.astype(np.float64) ** 2).sum(1), rtol=1e-4, atol=1e-5)


def test_stateful_layer_with_topk_masks_gradients_and_statistics(ops):
    from forde_b200.stateful_layer import StatefulLayer
    M, d, H, k = 300, 64, 256, 32
    layer = StatefulLayer(hidden_dim=H, top_k=k)
    x = torch.randn((M, d), device="cuda").bfloat16().requires_grad_(True)
    v = layer.init(2, x)
    a = np.arange(H) % 3
    v["neuron_assignments"]["assignments"].copy_(torch.as_tensor(a))
    ps = [v["params"]["up_proj"]["kernel"], v["params"]["down_proj"]["kernel"]]
    for p_ in ps:
        p_.requires_grad_(True)
    y, upd = layer.apply(v, x, mutable=["stats_buffer"])
    dy = torch.randn_like(y)
    y.backward(dy)
    torch.cuda.synchronize()
    wu, wd = f32(ps[0].bfloat16()), f32(ps[1].bfloat16())
    z = O.bf16_round(f32(x) @ wu)
    h = O.bf16_round(O.gated_residual(z, a))
    hs, ref_idx = O.topk_sparsify(h, k)
    np.testing.assert_array_equal(layer.last_topk_indices.cpu().numpy(), ref_idx)
    assert relerr(f32(y), O.bf16_round(hs @ wd)) < 1e-2
    mask = np.zeros_like(h, bool); np.put_along_axis(mask, ref_idx, True, axis=1)
    dz = O.bf16_round((f32(dy) @ wd.T) * (O.multiplex_grad(z, a) + O.gate_of(a)) * mask)
    assert relerr(f32(x.grad), O.bf16_round(dz @ wu.T)) < 1e-2
    assert relerr(f32(ps[0].grad), f32(x).T @ dz) < 1e-2
    st = f32(upd["stats_buffer"]["neuron_stats"])
    ref = O.calculate_neuron_stats(z, dz)
    np.testing.assert_allclose(st[:, 1], ref[:, 1], rtol=2e-2, atol=1e-5)      # grad_gdp of the MASKED dL/dz
    np.testing.assert_allclose(st[:, 3], ref[:, 3], rtol=2e-2, atol=1e-5)


@pytest.mark.parametrize("pattern", ["blocks64", "zeros", "no_tanh_mixed"])
def test_backward_reads_h_where_z_was_not_saved(ops, pattern):
    """save_z=0 (the module's default): z^T is only written where a 32-neuron group contains a tanh neuron; elsewhere
    the backward takes the relu mask from sign(h).  Gradients and statistics must equal the oracle's all the same."""
    from forde_b200.stateful_layer import StatefulLayer
    M, d, H = 700, 128, 512
    a = {"blocks64": (np.arange(H) // 64) % 3, "zeros": np.zeros(H, int),
         "no_tanh_mixed": np.where(np.arange(H) % 2 == 0, 0, 2)}[pattern]
    layer = StatefulLayer(hidden_dim=H)
    assert layer.save_z is False
    x = torch.randn((M, d), device="cuda").bfloat16().requires_grad_(True)
    v = layer.init(4, x)
    v["neuron_assignments"]["assignments"].copy_(torch.as_tensor(a))
    ps = [v["params"]["up_proj"]["kernel"], v["params"]["up_proj"]["bias"], v["params"]["down_proj"]["kernel"]]
    ps[1].normal_(0, 0.3)
    for p_ in ps:
        p_.requires_grad_(True)
    # poison the workspace's z^T so a read of an unsaved box cannot go unnoticed
    y0 = layer.apply(v, x)
    for ws in layer._ws[(M, H, str(x.device))]:
        ws.zT.fill_(float("nan"))
    y, upd = layer.apply(v, x, mutable=["stats_buffer"])
    dy = torch.randn_like(y)
    y.backward(dy)
    torch.cuda.synchronize()
    wu, wd = f32(ps[0].bfloat16()), f32(ps[2].bfloat16())
    z = O.bf16_round(f32(x) @ wu + f32(ps[1]))
    h = O.bf16_round(O.gated_residual(z, a))
    dz = O.bf16_round((f32(dy) @ wd.T) * (O.multiplex_grad(z, a) + O.gate_of(a)))
    assert np.isfinite(f32(x.grad)).all()
    assert relerr(f32(x.grad), O.bf16_round(dz @ wu.T)) < 1e-2
    assert relerr(f32(ps[0].grad), f32(x).T @ dz) < 1e-2
    assert relerr(f32(ps[2].grad), h.T @ f32(dy)) < 1e-2
    st = f32(upd["stats_buffer"]["neuron_stats"])
    np.testing.assert_allclose(st, O.calculate_neuron_stats(z, dz), rtol=2e-2, atol=2e-4)


def _pack_sign_bits(zT_bool: torch.Tensor, words: int, layout: int = 0) -> torch.Tensor:
    """[H, M] bool -> [H, words] int32; layout 0: bit j of word w = sample 32 w + j; layout 1 (builds with packed
    bf16x2 epilogues): sample 2k -> bit k, sample 2k + 1 -> bit 16 + k."""
    H, M = zT_bool.shape
    pad = torch.zeros((H, words * 32), dtype=torch.int64, device=zT_bool.device)
    pad[:, :M] = zT_bool.to(torch.int64)
    j = torch.arange(32, device=pad.device)
    pos = j if layout == 0 else (j >> 1) + 16 * (j & 1)
    w = (pad.view(H, words, 32) << pos).sum(-1)
    return torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)


@pytest.mark.parametrize("shape", [(160, 128, 512), (1000, 256, 640), (7, 64, 136), (2048, 768, 3072), (2500, 256, 3000)])
@pytest.mark.parametrize("pattern", ["zeros", "blocks", "mixed", "no_tanh_mixed"])
def test_sign_bit_path_is_bit_identical_to_the_h_path(ops, shape, pattern):
    """forde_sl_up_fwd_signs / forde_sl_dz_bwd_signs: where z is not stored the forward records (z > 0) bits and the
    backward reads those instead of h^T.  Integer/bit work: the bits equal (z > 0) of the stored-z run exactly, and
    h^T, dz^T and both partial-sum buffers are bit-identical to the default path (single-CTA and CTA-pair tiles,
    ragged M and H, neuron groups with and without tanh neurons)."""
    M, d, H = shape
    a = {"zeros": np.zeros(H, int), "blocks": np.arange(H) * 3 // H, "mixed": np.random.RandomState(3).randint(0, 3, H),
         "no_tanh_mixed": np.where(np.arange(H) % 2 == 0, 0, 2)}[pattern]
    t = run_layer(ops, M, d, H, a, seed=5, bias_scale=0.3)     # default path, every z box stored
    words = ops.sign_words(M)
    assert words == 8 * ((M + 255) // 256)
    zT2 = torch.full_like(t["zT"], float("nan")); hT2 = torch.zeros_like(t["hT"])
    layout = ops.sign_layout()
    signs_buf = torch.full(ops.sign_shape(M, H), 0x5A5A5A5A, device="cuda", dtype=torch.int32)
    signs = signs_buf.t() if layout & 2 else signs_buf          # [H, words] view either way
    actp2 = torch.zeros_like(t["actp"]); gradp2 = torch.zeros_like(t["gradp"]); dzT2 = torch.zeros_like(t["dzT"])
    ops.sl_up_fwd(t["x"], t["w_up"], t["b_up"], t["assign"], zT2, hT2, actp2, save_z=False, sign_bits=signs_buf)
    ops.sl_dz_bwd(t["dy"], t["w_down"], zT2, None, t["assign"], dzT2, gradp2, sign_bits=signs_buf)
    torch.cuda.synchronize()
    assert torch.equal(hT2[:, :M], t["hT"][:, :M]) and torch.equal(actp2, t["actp"])
    has_tanh = torch.as_tensor(a == 1, device="cuda")
    grp = torch.arange(H, device="cuda") // 32
    grp_tanh = torch.zeros(int(grp.max()) + 1, dtype=torch.bool, device="cuda").index_put_((grp,), has_tanh, accumulate=True)[grp]
    want = _pack_sign_bits(t["zT"][:, :M] > 0, words, layout & 1)
    valid = torch.zeros((H, words * 32), dtype=torch.bool, device="cuda"); valid[:, :M] = True
    vmask = _pack_sign_bits(valid, words, layout & 1)                                 # samples beyond M carry no meaning
    rows = ~grp_tanh
    assert torch.equal(signs[rows] & vmask[rows], want[rows]), "sign bits differ from (z > 0)"
    if bool(grp_tanh.any()):                                              # tanh groups: z box stored, bits untouched
        assert torch.equal(zT2[grp_tanh][:, :M], t["zT"][grp_tanh][:, :M])
        assert bool((signs[grp_tanh] == 0x5A5A5A5A).all())
    assert torch.equal(dzT2[:, :M], t["dzT"][:, :M]), "dz differs between the sign-bit and h^T paths"
    assert torch.equal(gradp2, t["gradp"])


def test_module_with_sign_bits_matches_default_module(ops):
    """StatefulLayer(sign_bits=True) end to end through autograd: same outputs, gradients and statistics, bit for bit."""
    from forde_b200.stateful_layer import StatefulLayer
    M, d, H = 700, 128, 512
    a = (np.arange(H) // 64) % 3
    outs = []
    for sb in (False, True):
        layer = StatefulLayer(hidden_dim=H, sign_bits=sb)
        x = torch.randn((M, d), device="cuda", generator=torch.Generator(device="cuda").manual_seed(9)).bfloat16().requires_grad_(True)
        v = layer.init(4, x)
        v["neuron_assignments"]["assignments"].copy_(torch.as_tensor(a))
        ps = [v["params"]["up_proj"]["kernel"], v["params"]["up_proj"]["bias"], v["params"]["down_proj"]["kernel"]]
        ps[1].copy_(torch.linspace(-0.5, 0.5, H, device="cuda"))
        for p_ in ps:
            p_.requires_grad_(True)
        y, upd = layer.apply(v, x, mutable=["stats_buffer"])
        y.backward(torch.ones_like(y))
        torch.cuda.synchronize()
        ws = layer._ws[(M, H, str(x.device))][0]
        assert (ws.signs is not None) == sb
        outs.append([y.detach(), x.grad, ps[0].grad, ps[1].grad, ps[2].grad, upd["stats_buffer"]["neuron_stats"]])
    for p0, p1 in zip(*outs):
        assert torch.equal(p0, p1)


def test_workspace_is_released_without_a_backward_and_stale_backward_is_refused(ops):
    """A grad-enabled forward that never runs its backward must not pin z^T / h^T forever, more than
    max_workspaces pending forwards raise, and a second backward through a retained graph after the buffers were
    reused raises instead of computing from overwritten tensors."""
    from forde_b200.stateful_layer import StatefulLayer
    M, d, H = 64, 64, 128
    layer = StatefulLayer(hidden_dim=H)
    x = torch.randn((M, d), device="cuda").bfloat16().requires_grad_(True)
    v = layer.init(0, x)
    for _ in range(10):                                   # evaluation without no_grad: each graph dies right away
        y = layer.apply(v, x)
        del y
    assert len(layer._ws[(M, H, str(x.device))]) == 1
    held = [layer.apply(v, x) for _ in range(layer.max_workspaces)]      # genuinely pending forwards
    with pytest.raises(RuntimeError, match="waiting for their backward"):
        layer.apply(v, x)
    held[0].sum().backward(retain_graph=True)             # releases workspace 0 ...
    y_new = layer.apply(v, x)                             # ... which a new forward takes and overwrites
    with pytest.raises(RuntimeError, match="overwritten by a later forward"):
        held[0].sum().backward()
    del held, y_new


def test_module_edge_cases_empty_batch_and_bad_widths(ops):
    from forde_b200.stateful_layer import StatefulLayer
    layer = StatefulLayer(hidden_dim=256)
    x = torch.zeros((0, 7, 64), device="cuda", dtype=torch.bfloat16)
    v = layer.init(0, x)
    y, upd = layer.apply(v, x, mutable=["stats_buffer"])
    assert y.shape == (0, 7, 64) and int(upd["stats_buffer"]["step_count"].item()) == 0
    bad = StatefulLayer(hidden_dim=100)
    xb = torch.zeros((4, 64), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="multiples of 8"):
        bad.apply(bad.init(0, xb), xb)


def test_clip_loss_row_block_of_a_larger_global_batch(ops):
    """BASELINE config 3 shape in miniature: one rank's rows (its images x ALL texts) of the global logits; per-rank
    pieces combined the way clip_loss.py does must equal the oracle on the global batch."""
    Bg, E, world = 384, 64, 3
    Bl = Bg // world
    g = torch.Generator(device="cuda").manual_seed(11)
    img = torch.randn((Bg, E), device="cuda", generator=g).bfloat16()
    txt = (torch.randn((Bg, E), device="cuda", generator=g) + 0.7 * img.float()).bfloat16()
    lt = torch.tensor([1.9], device="cuda")
    pieces = [ops.clip_fwd(img[r * Bl:(r + 1) * Bl].contiguous(), txt, lt, r * Bl) for r in range(world)]
    gmax = torch.stack([p["col_max"] for p in pieces]).max(0).values          # all-reduce(MAX)
    csum = torch.zeros_like(gmax)
    for p in pieces:
        ops.clip_col_rescale(p["col_max"], gmax, p["col_sum"])
        csum += p["col_sum"]                                                   # all-reduce(SUM)
    loss = sum(ops.clip_loss(p, gmax, csum, r * Bl)[0] for r, p in enumerate(pieces))
    col_lse = ops.clip_loss(pieces[0], gmax, csum, 0)[1]
    dV = torch.zeros((Bg, E), device="cuda"); dU = []
    for r, p in enumerate(pieces):
        du, dv, _ = ops.clip_bwd(p, col_lse, r * Bl)
        dU.append(du); dV += dv                                                # reduce-scatter(SUM)
    torch.cuda.synchronize()
    ref_loss, S, u, v = O.clip_loss_fwd(f32(img), f32(txt), 1.9)
    assert abs(float(loss.item()) - float(ref_loss)) < 1e-2
    _, _, _, dS = O.clip_loss_bwd(f32(img), f32(txt), 1.9)
    scale = float(np.exp(1.9))
    assert relerr(f32(torch.cat(dU)), scale * dS @ v) < 3e-2
    assert relerr(f32(dV), scale * dS.T @ u) < 3e-2


def test_composite_abi_entry_points_match_the_stepwise_calls(ops):
    """forde_sl_fwd / forde_sl_bwd (the two calls an XLA-FFI handler makes) against the per-kernel entry
    points: bit-identical outputs."""
    from forde_b200 import _lib
    lib = _lib.load()
    M, d, H = 600, 128, 384
    a = (np.arange(H) // 32) % 3
    t = run_layer(ops, M, d, H, a)                      # step-wise reference (save_z=1)
    ld, P = ops.sample_pitch(M), ops.num_partials(M)
    st = torch.cuda.current_stream().cuda_stream
    new = lambda like: torch.zeros_like(like)
    zT, hT, y, actp = new(t["zT"]), new(t["hT"]), new(t["y"]), new(t["actp"])
    rc = lib.forde_sl_fwd(st, t["x"].data_ptr(), t["w_up"].data_ptr(), t["b_up"].data_ptr(), t["w_down"].data_ptr(),
                          t["b_down"].data_ptr(), t["assign"].data_ptr(), zT.data_ptr(), hT.data_ptr(), y.data_ptr(),
                          actp.data_ptr(), M, d, H, ld, 1)
    assert rc == 0
    dzT, dx, gradp = new(t["dzT"]), new(t["dx"]), new(t["gradp"])
    dw_up, db_up, dw_down, db_down = new(t["dw_up"]), new(t["db_up"]), new(t["dw_down"]), new(t["db_down"])
    rc = lib.forde_sl_bwd(st, t["x"].data_ptr(), t["dy"].data_ptr(), t["w_up"].data_ptr(), t["w_down"].data_ptr(),
                          zT.data_ptr(), hT.data_ptr(), t["assign"].data_ptr(), dzT.data_ptr(), dx.data_ptr(),
                          dw_up.data_ptr(), db_up.data_ptr(), dw_down.data_ptr(), db_down.data_ptr(), gradp.data_ptr(),
                          M, d, H, ld)
    assert rc == 0
    torch.cuda.synchronize()
    for name, got in (("zT", zT), ("hT", hT), ("y", y), ("actp", actp), ("dzT", dzT), ("dx", dx), ("gradp", gradp)):
        assert torch.equal(got, t[name]), name
    # split-K red.add order is not fixed: fp32 weight gradients agree to rounding, bias gradients too
    torch.testing.assert_close(dw_up, t["dw_up"], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(dw_down, t["dw_down"], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(db_up, t["db_up"], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(db_down, t["db_down"], rtol=1e-5, atol=1e-4)
    # statistics from already pooled moments == per-step finalisation (single step)
    mom = torch.zeros((H, 5), device="cuda"); cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
    ops.stats_finalize(actp, gradp, P, H, M, moments=mom, sample_count=cnt)
    pooled = ops.stats_from_moments(mom, cnt)
    torch.cuda.synchronize()
    torch.testing.assert_close(pooled, t["stats"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("shape", [(256, 512, 256), (1000, 264, 72), (50432, 768, 768)])
def test_gemm_bf16_accumulate_in_place(ops, shape):
    """EPI_STORE_ADD: D += A B^T + bias with the C tile read and overwritten in the epilogue."""
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(7)
    A = torch.randn((M, K), device="cuda", generator=g).bfloat16()
    B = torch.randn((N, K), device="cuda", generator=g).bfloat16()
    C = torch.randn((M, N), device="cuda", generator=g).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = (A.float() @ B.float().t() + bias + C.float()).bfloat16()
    out = ops.gemm_bf16(A, B, False, False, bias_col=bias, add_to=C)
    torch.cuda.synchronize()
    assert out.data_ptr() == C.data_ptr()
    err = (out.float() - ref.float()).norm() / ref.float().norm()
    assert err.item() < 5e-3


def test_fused_block_half_equals_layer_then_add_norm(ops):
    """apply(..., post_norm=ln) == LayerNorm(x + apply(x)) with autograd doing the skip-path add: same outputs and
    gradients (the fused form adds the skip gradient in the dx GEMM's epilogue)."""
    from forde_b200.stateful_layer import StatefulLayer
    M, d, H = 520, 128, 384
    a = (np.arange(H) // 48) % 3
    xs, res = [], []
    for fused in (False, True):
        layer = StatefulLayer(hidden_dim=H)
        x = torch.randn((M, d), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)).bfloat16().requires_grad_(True)
        v = layer.init(5, x)
        v["neuron_assignments"]["assignments"].copy_(torch.as_tensor(a))
        ps = [v["params"]["up_proj"]["kernel"], v["params"]["down_proj"]["kernel"], v["params"]["down_proj"]["bias"]]
        ln = {"scale": (torch.rand(d, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4)) + 0.5).requires_grad_(True),
              "bias": torch.zeros(d, device="cuda").requires_grad_(True)}
        for p_ in ps:
            p_.requires_grad_(True)
        if fused:
            out = layer.apply(v, x, post_norm=ln)
        else:
            y = layer.apply(v, x)
            out = torch.nn.functional.layer_norm(x.float() + y.float(), (d,), ln["scale"], ln["bias"], 1e-6).bfloat16()
        dout = torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(6)).bfloat16()
        out.backward(dout)
        torch.cuda.synchronize()
        res.append([f32(out), f32(x.grad), f32(ps[0].grad), f32(ps[1].grad), f32(ps[2].grad), f32(ln["scale"].grad)])
    for name, r0, r1 in zip(("out", "dx", "dw_up", "dw_down", "db_down", "d_gamma"), *res):
        assert relerr(r1, r0) < 1.5e-2, name
