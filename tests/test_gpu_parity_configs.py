"""Oracle parity at the BASELINE config shapes themselves (C2 vision / C2 text / C5 vision / C5 text
layers, the C3 contrastive block), the optimizer step, and the pooled statistics of a whole slow-loop interval.

The kernels run at FULL size; the NumPy oracle (oracle/forde_oracle.py) checks a seeded SAMPLE of them so it stays
fast: >= 2048 sample rows x ALL neurons for everything that is row-local (z, h, y, dz, dx), and 192 neuron columns x
ALL samples for everything that reduces over samples (dW, db, the five statistics).  Rows are drawn so that the first
tile, the ragged last tile and tile boundaries are always in the sample.  Tolerances are the ones of
tests/test_gpu_parity.py (assert_close_bf16: every element within 2^-7 |ref| + 1e-3 max(1, rms ref))."""
import numpy as np
import pytest
import torch

from oracle import forde_oracle as O
from test_gpu_parity import assert_close_bf16, f32, relerr, run_layer

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import __graft_entry__ as g
    g.build()
    from forde_b200 import ops as _ops
    return _ops


CONFIG_LAYERS = {
    "c2_vision": (256 * 197, 768, 3072),     # BASELINE configs[1], vision tower layer
    "c2_text": (256 * 77, 512, 2048),        # BASELINE configs[1], text tower layer (the 2.08-wave shapes)
    "c5_vision": (256 * 257, 1024, 4096),    # BASELINE configs[4], ViT-L/14 vision layer
    "c5_text": (256 * 77, 768, 3072),        # BASELINE configs[4], text layer
}


def _sample_rows(M, n, rng):
    must = [0, 1, 127, 128, 255, 256, 257, M // 2, M - 257, M - 256, M - 129, M - 128, M - 2, M - 1]
    rows = np.unique(np.concatenate([np.array([r for r in must if 0 <= r < M]), rng.choice(M, size=n, replace=False)]))
    return rows


@pytest.mark.parametrize("name", list(CONFIG_LAYERS))
@pytest.mark.parametrize("pattern", ["thirds", "iid"])
def test_layer_parity_at_config_shape_against_sampled_oracle(ops, name, pattern):
    M, d, H = CONFIG_LAYERS[name]
    if pattern == "iid" and name not in ("c2_text",):
        pytest.skip("the i.i.d. map (every warp mixed) is covered on the text shape; thirds = a smoothed brain map")
    rng = np.random.RandomState(17)
    assign = (np.arange(H) * 3 // H) if pattern == "thirds" else rng.randint(0, 3, H)
    t = run_layer(ops, M, d, H, assign, seed=11, bias_scale=0.2)
    rows = _sample_rows(M, 2048, rng)
    cols = np.unique(np.concatenate([[0, 31, 32, H // 3 - 1, H // 3, H - 1], rng.choice(H, size=192, replace=False)]))
    rows_t = torch.as_tensor(rows, device="cuda"); cols_t = torch.as_tensor(cols, device="cuda")
    take_r = lambda name_: f32(t[name_][rows_t])                  # (rows, :) of a token-major tensor
    take_rT = lambda name_: f32(t[name_][:, rows_t].t())          # (rows, H) of a neuron-major tensor
    x_r, dy_r = take_r("x"), take_r("dy")
    wu, wd, bu, bd = f32(t["w_up"]), f32(t["w_down"]), f32(t["b_up"]), f32(t["b_down"])

    # ---- row-local tensors on the sampled rows, all neurons
    fw = O.stateful_layer_fwd(x_r, wu, bu, wd, bd, assign)
    z_k, h_k = take_rT("zT"), take_rT("hT")
    assert_close_bf16(z_k, fw["z"], "z")
    h_ref = O.stateful_layer_fwd_from_z(z_k, assign)              # discrete decisions from the kernel's own stored z
    for a_val in (0, 2):
        m = assign == a_val
        np.testing.assert_array_equal(O.bf16_bits(h_k[:, m]), O.bf16_bits(h_ref[:, m]))
    m1 = assign == 1
    assert np.abs(h_k[:, m1] - h_ref[:, m1]).max() <= 2.0 ** -7 * np.abs(h_ref[:, m1]).max() + 1e-6
    assert_close_bf16(take_r("y"), O.bf16_round(h_k @ wd + bd), "y")
    bw = O.stateful_layer_bwd(x_r, wu, wd, assign, z_k, h_k, dy_r)
    dz_k = take_rT("dzT")
    assert_close_bf16(dz_k, bw["dz"], "dz")
    m2 = assign == 2
    assert_close_bf16(dz_k[:, m2], O.bf16_round(np.float32(1.1) * (dy_r @ wd.T)[:, m2]), "dz(STE)")
    assert_close_bf16(take_r("dx"), O.bf16_round(dz_k @ wu.T), "dx")

    # ---- sample-reducing outputs on the sampled neuron columns, all samples
    x_all, dy_all = f32(t["x"]), f32(t["dy"])
    z_c = f32(t["zT"][cols_t, :M].t()); h_c = f32(t["hT"][cols_t, :M].t()); dz_c = f32(t["dzT"][cols_t, :M].t())
    assert relerr(f32(t["dw_up"][:, cols_t]), x_all.T.astype(np.float64) @ dz_c.astype(np.float64)) < 1e-3
    assert relerr(f32(t["dw_down"][cols_t]), h_c.T.astype(np.float64) @ dy_all.astype(np.float64)) < 1e-3
    np.testing.assert_allclose(f32(t["db_up"][cols_t]), dz_c.astype(np.float64).sum(0), rtol=1e-3, atol=2e-3)
    np.testing.assert_allclose(f32(t["db_down"]), dy_all.astype(np.float64).sum(0), rtol=1e-3, atol=2e-2)
    # the five statistics (sensing.py:34-69) of the reference's separate pass over (z, dL/dz)
    np.testing.assert_allclose(f32(t["stats"][cols_t]), O.calculate_neuron_stats(z_c, dz_c), rtol=1e-3, atol=1e-5)


def test_forced_tail_cut_is_bit_identical_to_the_single_launch(ops):
    """C2 text shape through forde_sl_down_fwd / forde_sl_dx_bwd(_add): the sample rows split into two launches at a
    tile-row boundary (how a caller may pipeline a batch) must give the same bits as one launch -- row blocks are
    independent, no state leaks between launches, and offset base pointers keep TMA's alignment rules."""
    M, d, H = 256 * 77, 512, 2048
    t = run_layer(ops, M, d, H, np.arange(H) % 3, seed=3)
    from forde_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    cut = 74 * 256
    ld = t["hT"].stride(0)
    y2 = torch.zeros_like(t["y"]); dx2 = torch.zeros_like(t["dx"])
    res = torch.randn_like(t["dx"])
    dx_add1 = torch.empty_like(res); dx_add2 = torch.empty_like(res)
    for r0, rows in ((0, cut), (cut, M - cut)):
        assert lib.forde_sl_down_fwd(st, t["hT"].data_ptr() + 2 * r0, t["w_down"].data_ptr(), t["b_down"].data_ptr(),
                                     y2.data_ptr() + 2 * r0 * d, rows, d, H, ld) == 0
        assert lib.forde_sl_dx_bwd(st, t["dzT"].data_ptr() + 2 * r0, t["w_up"].data_ptr(), dx2.data_ptr() + 2 * r0 * d,
                                   rows, d, H, ld) == 0
        assert lib.forde_sl_dx_bwd_add(st, t["dzT"].data_ptr() + 2 * r0, t["w_up"].data_ptr(),
                                       res.data_ptr() + 2 * r0 * d, dx_add2.data_ptr() + 2 * r0 * d, rows, d, H, ld) == 0
    assert lib.forde_sl_dx_bwd_add(st, t["dzT"].data_ptr(), t["w_up"].data_ptr(), res.data_ptr(), dx_add1.data_ptr(),
                                   M, d, H, ld) == 0
    torch.cuda.synchronize()
    assert torch.equal(y2, t["y"]) and torch.equal(dx2, t["dx"]) and torch.equal(dx_add1, dx_add2)


def test_c3_contrastive_block_1024_local_of_8192_global(ops):
    """BASELINE configs[2]: global batch 8192 over 8 ranks = 1024 local image rows x 8192 gathered texts x E = 512.
    Eight emulated ranks run the per-rank kernels at full size and are combined exactly as clip_loss.py combines
    them over NCCL (MAX / SUM of the column pieces, SUM of dV); loss, dU, dV, d(log_temp) against the NumPy oracle
    on the whole 8192 x 8192 problem."""
    Bg, E, world = 8192, 512, 8
    Bl = Bg // world
    g = torch.Generator(device="cuda").manual_seed(5)
    img = torch.randn((Bg, E), device="cuda", generator=g).bfloat16()
    txt = (torch.randn((Bg, E), device="cuda", generator=g) + 0.7 * img.float()).bfloat16()
    lt_val = float(np.log(1 / 0.07))
    lt = torch.tensor([lt_val], device="cuda")
    pieces = [ops.clip_fwd(img[r * Bl:(r + 1) * Bl].contiguous(), txt, lt, r * Bl) for r in range(world)]
    gmax = torch.stack([p["col_max"] for p in pieces]).max(0).values          # all-reduce(MAX)
    csum = torch.zeros_like(gmax)
    for p in pieces:
        ops.clip_col_rescale(p["col_max"], gmax, p["col_sum"])
        csum += p["col_sum"]                                                   # all-reduce(SUM)
    loss = sum(ops.clip_loss(p, gmax, csum, r * Bl)[0] for r, p in enumerate(pieces))
    col_lse = ops.clip_loss(pieces[0], gmax, csum, 0)[1]
    dV = torch.zeros((Bg, E), device="cuda"); dU = []; dlt = torch.zeros(1, device="cuda")
    for r, p in enumerate(pieces):
        du, dv, dl = ops.clip_bwd(p, col_lse, r * Bl)
        dU.append(du); dV += dv; dlt += dl                                     # reduce-scatter(SUM) / all-reduce(SUM)
    torch.cuda.synchronize()
    i32, t32 = f32(img), f32(txt)
    ref_loss, S, u, v = O.clip_loss_fwd(i32, t32, lt_val)
    assert abs(float(loss.item()) - float(ref_loss)) < 1e-2, (float(loss.item()), float(ref_loss))
    # one rank's logits block, element-wise (fp32 S from bf16 u, v: 2^-8 relative on |u.v| <= 1, times exp(tau) = 14.3)
    S3 = f32(pieces[3]["S"])
    assert np.abs(S3 - S[3 * Bl:4 * Bl]).max() < 14.3 * 2.0 ** -7
    del S
    _, _, dlt_ref, dS = O.clip_loss_bwd(i32, t32, lt_val)
    scale = float(np.exp(lt_val))
    assert relerr(f32(torch.cat(dU)), scale * dS @ v) < 3e-2
    assert relerr(f32(dV), scale * dS.T @ u) < 3e-2
    assert abs(float(dlt.item()) - float(dlt_ref)) < 3e-2 * max(1.0, abs(float(dlt_ref)))


# ------------------------------------------------------------------------------------------ statistics of an interval
def test_pooled_moments_over_a_100_step_interval_with_bias_dominated_neurons(ops):
    """The slow loop senses from raw moments pooled over a whole interval (default 100 steps, train.py:219-224).
    Variance = sum z^2 / n - mean^2 cancels when |mean| >> std: bias-dominated neurons (mean up to 24, std ~ 0.5).
    100 forward+backward accumulations through forde_stats_finalize, then forde_stats_from_moments, against
    O.calculate_neuron_stats (two-pass variance, sensing.py:59) on the CONCATENATED samples of all 100 steps."""
    M, d, H, steps = 1536, 64, 256, 100
    g = torch.Generator(device="cuda").manual_seed(2)
    w_up = (torch.randn((d, H), device="cuda", generator=g) * (0.5 / d ** 0.5)).bfloat16()
    w_down = (torch.randn((H, d), device="cuda", generator=g) / H ** 0.5).bfloat16()
    b_up = torch.linspace(-24.0, 24.0, H, device="cuda")            # |mean| / std up to ~50
    a = np.arange(H) % 3
    ai8 = torch.as_tensor(a, device="cuda").to(torch.int8)
    ld, P = ops.sample_pitch(M), ops.num_partials(M)
    zT = torch.zeros((H, ld), device="cuda", dtype=torch.bfloat16); hT = torch.zeros_like(zT); dzT = torch.zeros_like(zT)
    actp = torch.zeros((P, 3, H), device="cuda"); gradp = torch.zeros((P, 3, H), device="cuda")
    from forde_b200.stateful_layer import new_stats_buffer
    sb = new_stats_buffer(H, "cuda")
    plain = torch.zeros((H, 5), device="cuda")                       # what a plain fp32 accumulator would hold
    zs, dzs = [], []
    for s in range(steps):
        x = torch.randn((M, d), device="cuda", generator=g).bfloat16()
        dy = torch.randn((M, d), device="cuda", generator=g).bfloat16()
        ops.sl_up_fwd(x, w_up, b_up, ai8, zT, hT, actp)
        ops.sl_dz_bwd(dy, w_down, zT, hT, ai8, dzT, gradp)
        ops.stats_finalize(actp, gradp, P, H, M, neuron_stats=sb["neuron_stats"], moments=sb["moments"],
                           moments_lo=sb["moments_lo"], sample_count=sb["sample_count"], step_count=sb["step_count"])
        ops.stats_finalize(actp, gradp, P, H, M, moments=plain)
        zs.append(zT[:, :M].t().float().cpu().numpy()); dzs.append(dzT[:, :M].t().float().cpu().numpy())
    torch.cuda.synchronize()
    assert int(sb["step_count"].item()) == steps and int(sb["sample_count"].item()) == steps * M
    z_all, dz_all = np.concatenate(zs), np.concatenate(dzs)
    ref = O.calculate_neuron_stats(z_all.astype(np.float64).astype(np.float32), dz_all)
    ref_var = z_all.astype(np.float64).var(0)                        # the two-pass variance in fp64 (ground truth)
    got = f32(ops.stats_from_moments(ops.pooled_moments(sb), sb["sample_count"].reshape(1)))
    np.testing.assert_allclose(got[:, 4], ref_var, rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(got[:, :4], ref[:, :4], rtol=1e-3, atol=1e-5)
    # the raw sums agree with fp64 sums to the rounding of the per-slot fp32 partials (128 samples each, ~1e-7
    # relative, not growing with the number of steps): accumulation over the interval adds nothing to it
    mom_ref, n = O.raw_moments(z_all, dz_all)
    np.testing.assert_allclose(ops.pooled_moments(sb).cpu().numpy(), mom_ref, rtol=2e-6)
    err_plain = np.abs(f32(ops.stats_from_moments(plain, sb["sample_count"].reshape(1)))[:, 4] - ref_var) / ref_var
    print(f"pooled variance: max rel err double-single {np.abs(got[:, 4] - ref_var).max():.2e} abs; "
          f"a plain fp32 accumulator would be off by up to {err_plain.max():.2e} relative")


# ------------------------------------------------------------------------------------------ optimizer step
@pytest.mark.parametrize("grad_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("gscale", [3.0, 1e-4])
def test_fused_clip_adamw_matches_the_optax_restatement(ops, grad_dtype, gscale):
    """forde_grad_sqnorm + forde_adamw_clip_step against O.optax_clip_adamw_step (optax.chain(clip_by_global_norm(1.0),
    adamw), src/training/train.py:74-77) over 4 steps, with gradients above (clipped) and below the threshold, fp32
    gradients and the bf16 gradients of the data-parallel exchange.  fp32 tolerances: params / mu / nu rtol 2e-6."""
    n = 64 * 9973                                                     # flat buffer, a multiple of the 64-element pad
    g = torch.Generator(device="cuda").manual_seed(1)
    p = torch.randn(n, device="cuda", generator=g)
    mu = torch.zeros(n, device="cuda"); nu = torch.zeros(n, device="cuda")
    p16 = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
    count = torch.zeros((), device="cuda", dtype=torch.int32)
    partials = torch.zeros(ops.optim_num_partials(n), device="cuda")
    gn = torch.zeros(1, device="cuda")
    rp, rmu, rnu, rc = f32(p), f32(mu), f32(nu), 0
    lr, wd = 1e-2, 0.01
    for step in range(4):
        grads = (torch.randn(n, device="cuda", generator=g) * gscale / n ** 0.5 * (1 + step)).to(grad_dtype)
        ops.grad_sqnorm(grads, partials, count)
        ops.adamw_clip_step(p, grads, mu, nu, p16, partials, count, gn, 1.0, lr, 0.9, 0.999, 1e-8, wd)
        torch.cuda.synchronize()
        rp, rmu, rnu, rc, rnorm = O.optax_clip_adamw_step(rp, f32(grads), rmu, rnu, rc, lr, wd)
        assert int(count.item()) == rc
        assert abs(float(gn.item()) - float(rnorm)) <= 1e-5 * float(rnorm)
        # (the kernel contracts b * m + (1 - b) * g into one FMA, NumPy rounds twice: ~1e-7 of the larger term, which
        # is a large RELATIVE error only where the two terms cancel)
        np.testing.assert_allclose(f32(mu), rmu, rtol=2e-6, atol=1e-6 * float(np.abs(rmu).max()))
        np.testing.assert_allclose(f32(nu), rnu, rtol=2e-6, atol=1e-6 * float(np.abs(rnu).max()))
        np.testing.assert_allclose(f32(p), rp, rtol=2e-6, atol=1e-6)
        assert torch.equal(p16, p.bfloat16())                         # the compute mirror is bf16(p) exactly
    # deterministic: the same inputs give the same bits
    q1, q2 = p.clone(), p.clone()
    m1, m2, n1, n2 = mu.clone(), mu.clone(), nu.clone(), nu.clone()
    c1, c2 = count.clone(), count.clone()
    for q, m_, n_, c_ in ((q1, m1, n1, c1), (q2, m2, n2, c2)):
        ops.grad_sqnorm(grads, partials, c_)
        ops.adamw_clip_step(q, grads, m_, n_, None, partials, c_, gn, 1.0, lr, 0.9, 0.999, 1e-8, wd)
    torch.cuda.synchronize()
    assert torch.equal(q1, q2)


def test_train_state_takes_one_optax_step_per_batch(ops):
    """train_step on the tiny dual encoder: the parameters after one step equal the optax restatement applied to the
    gradients the step produced (flat_g32 still holds them), and grad_norm is the pre-clip global norm."""
    from forde_b200 import dual_encoder, train
    cfg = dual_encoder.config_tiny()
    st = train.create_train_state(cfg, 0, learning_rate=1e-3, weight_decay=0.05, device="cuda")
    b = {k: v.cuda() for k, v in train.synthetic_batch(cfg, 32, pinned=False).items()}
    p0 = f32(st.flat_p32)
    _, m = train.train_step(st, b)
    torch.cuda.synchronize()
    g = f32(st.flat_g32)
    rp, rmu, rnu, rc, rnorm = O.optax_clip_adamw_step(p0, g, np.zeros_like(p0), np.zeros_like(p0), 0, 1e-3, 0.05)
    assert abs(float(m["grad_norm"].item()) - float(rnorm)) <= 1e-5 * float(rnorm)
    np.testing.assert_allclose(f32(st.flat_p32), rp, rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(f32(st.opt_state["mu"]), rmu, rtol=2e-6, atol=1e-6 * float(np.abs(rmu).max()))
    assert int(st.opt_state["count"].item()) == 1 and st.step == 1
    assert torch.equal(st.flat_p16, st.flat_p32.detach().bfloat16())


# ------------------------------------------------------------------------------------------ tower front ends (8(f)-3)
@pytest.mark.parametrize("B,S,p,w", [(4, 32, 16, 128), (3, 224, 16, 768), (2, 224, 14, 1024)])
def test_vision_front_end_matches_plain_restatement(ops, B, S, p, w):
    """patchify -> patch-embedding GEMM -> +CLS, +pos (README.md:30-34) against a plain torch restatement (reshape /
    permute / matmul / cat in fp32 on the same bf16 inputs), forward and backward.  patchify is pure data movement:
    bit-exact.  p = 14 exercises the 588-wide (not 16-byte) patch rows of ViT-L/14."""
    from forde_b200.dual_encoder import _Dense, _VisionAssemble
    g_ = torch.Generator(device="cuda").manual_seed(B * S)
    g, k = S // p, p * p * 3
    img = torch.randn((B, S, S, 3), device="cuda", generator=g_).bfloat16()
    patches = ops.patchify(img, p)
    ref_patches = img.view(B, g, p, g, p, 3).permute(0, 1, 3, 2, 4, 5).reshape(B * g * g, k)
    assert torch.equal(patches, ref_patches)
    wk = (torch.randn((k, w), device="cuda", generator=g_) / k ** 0.5).requires_grad_(True)
    bk = (torch.randn(w, device="cuda", generator=g_) * 0.1).requires_grad_(True)
    cls = (torch.randn(w, device="cuda", generator=g_) * 0.02).requires_grad_(True)
    pos = (torch.randn((g * g + 1, w), device="cuda", generator=g_) * 0.02).requires_grad_(True)
    tok = _Dense.apply(patches, wk, bk, wk.detach().bfloat16(), None)
    x = _VisionAssemble.apply(tok, cls, pos, cls.detach().bfloat16(), pos.detach().bfloat16(), B)
    dx = torch.randn(x.shape, device="cuda", generator=g_).bfloat16()
    x.backward(dx)
    torch.cuda.synchronize()
    # restatement
    wr = wk.detach().bfloat16().float().requires_grad_(True); br = bk.detach().clone().requires_grad_(True)
    cr = cls.detach().bfloat16().float().requires_grad_(True); pr = pos.detach().bfloat16().float().requires_grad_(True)
    t_ = ref_patches.float() @ wr + br
    tok_r = t_ + (t_.detach().bfloat16().float() - t_.detach())     # value: bf16-rounded like the stored tokens; grad: identity
    x_r = torch.cat([cr.expand(B, 1, w), tok_r.view(B, g * g, w)], 1) + pr
    x_r.backward(dx.float())
    assert_close_bf16(f32(x), f32(x_r.detach().bfloat16().float()), "x")
    assert relerr(f32(wk.grad), f32(wr.grad)) < 2e-3 and relerr(f32(bk.grad), f32(br.grad)) < 2e-3
    assert relerr(f32(cls.grad), f32(cr.grad)) < 1e-3 and relerr(f32(pos.grad), f32(pr.grad)) < 1e-3


def test_token_embedding_front_end_matches_plain_restatement(ops):
    """x = token_embed[ids] + pos_embed (README.md:30-34) and its scatter-add backward against torch's embedding."""
    from forde_b200.dual_encoder import _TokenEmbed
    B, T, w, V = 64, 77, 512, 30522
    g_ = torch.Generator(device="cuda").manual_seed(3)
    ids = torch.randint(0, V, (B, T), device="cuda", generator=g_, dtype=torch.int32)
    ids[:, 0] = 101                                            # a token shared by every caption ([CLS]): hot row
    table = (torch.randn((V, w), device="cuda", generator=g_) * 0.02).requires_grad_(True)
    pos = (torch.randn((T, w), device="cuda", generator=g_) * 0.01).requires_grad_(True)
    x = _TokenEmbed.apply(ids, table, pos, table.detach().bfloat16(), pos.detach().bfloat16())
    dx = torch.randn(x.shape, device="cuda", generator=g_).bfloat16()
    x.backward(dx)
    torch.cuda.synchronize()
    tr = table.detach().bfloat16().float().requires_grad_(True); pr = pos.detach().bfloat16().float().requires_grad_(True)
    x_r = torch.nn.functional.embedding(ids.long(), tr) + pr
    x_r.backward(dx.float())
    assert torch.equal(x, x_r.detach().bfloat16())             # one fp32 add, one rounding: bit-exact
    assert relerr(f32(table.grad), f32(tr.grad)) < 1e-5 and relerr(f32(pos.grad), f32(pr.grad)) < 1e-5


# ------------------------------------------------------------------------------------------ short-sequence attention
@pytest.mark.parametrize("B,T,heads,blocked", [
    (4, 77, 8, False), (3, 16, 2, False), (2, 128, 4, False), (5, 1, 1, False), (256, 77, 8, False), (2, 100, 12, False),
    # csrc/forde_attn2.cu (two 128-row blocks, operand ring): the ViT-B/16 vision tower's 197 tokens at toy and full size,
    # every block-count / padding corner, more items than CTAs, and the one-block case forced onto the blocked kernel
    (2, 197, 12, True), (3, 256, 2, True), (2, 129, 4, True), (1, 144, 1, True), (2, 250, 3, True), (40, 197, 12, True),
    (4, 77, 8, True), (5, 1, 1, True), (3, 128, 2, True), (256, 197, 12, True)])
def test_attention_forward_backward_against_plain_softmax_attention(ops, B, T, heads, blocked):
    """csrc/forde_attn.cu / forde_attn2.cu against plain fp32 softmax(q k^T / 8) v and its autograd gradients on the same
    bf16 inputs.
    Tolerance: the probabilities are staged in bf16 (2^-8 relative each), outputs and gradients are bf16: relative
    Frobenius error <= 1e-2, and every element within 2^-6 |ref| + 5e-2 rms(ref) (dS = p (dP - D) is rounded to bf16
    after a cancellation, so a handful of the 10^7 gradient elements of the full-size case sit at 2-3 % of the rms)."""
    w = heads * 64
    g_ = torch.Generator(device="cuda").manual_seed(T * 131 + B)
    q, k, v = ((torch.randn((B, T, w), device="cuda", generator=g_) * s).bfloat16() for s in (1.5, 1.5, 1.0))
    d_o = torch.randn((B, T, w), device="cuda", generator=g_).bfloat16()
    o, lse = ops.attn_fwd(q.view(B * T, w), k.view(B * T, w), v.view(B * T, w), B, T, heads, blocked=blocked)
    dq, dk, dv = ops.attn_bwd(q.view(B * T, w), k.view(B * T, w), v.view(B * T, w), d_o.view(B * T, w), lse, B, T, heads,
                              o=o, blocked=blocked)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().view(B, T, heads, 64).transpose(1, 2).detach().requires_grad_(True) for t in (q, k, v))
    s = (qf @ kf.transpose(-1, -2)) * 0.125
    ref_lse = torch.logsumexp(s, -1)
    ref = torch.softmax(s, -1) @ vf
    ref.backward(d_o.float().view(B, T, heads, 64).transpose(1, 2))
    back = lambda t: t.transpose(1, 2).reshape(B * T, w)

    def close(got, want, name):
        got, want = f32(got), f32(want)
        if not want.any():      # T = 1: dq = dk = 0 exactly; with D taken from the bf16-rounded output they are ~2^-9 |dO . v|
            assert np.abs(got).max() <= 2.0 ** -6, f"{name}: {np.abs(got).max():.3e} where the reference is exactly zero"
            return
        assert relerr(got, want) <= 1e-2, f"{name}: rel Frobenius {relerr(got, want):.3e}"
        rms = float(np.sqrt(np.mean(want.astype(np.float64) ** 2)))
        err, bound = np.abs(got - want), 2.0 ** -6 * np.abs(want) + 5e-2 * rms
        bad = err > bound
        if blocked:
            # the blocked backward takes D = rowsum(dO * O) from the bf16-ROUNDED forward output (as flash attention
            # does), the reference from exact fp32 probabilities: of the ~10^7 gradient elements a few per million sit
            # just outside the bound above.  Stated: at most 1e-5 of the elements, none beyond twice the bound.
            assert bad.mean() <= 1e-5 and not (err > 2 * bound).any(), f"{name}: {bad.sum()} of {bad.size} elements off"
        else:
            assert not bad.any(), f"{name}: {bad.sum()} of {bad.size} elements off"

    np.testing.assert_allclose(f32(lse), f32(ref_lse), rtol=1e-3, atol=2e-3)
    close(o, back(ref), "o")
    close(dq, back(qf.grad), "dq"); close(dk, back(kf.grad), "dk"); close(dv, back(vf.grad), "dv")
