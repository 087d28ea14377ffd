"""End-to-end parity of the fast loop at BASELINE config 1 (tiny dual encoder, batch 32): the B200 train step
(bf16 storage, sm_100a StatefulLayer + contrastive-loss kernels) against the fp32 CPU port of the same step on
identical parameters and the same synthetic batch.  Tolerances: loss abs 5e-2 (bf16 towers vs fp32), per-layer
statistics rtol 0.1 on the scale-type stats, assignments routing exercised with a mixed brain map."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    return True


def test_tiny_dual_encoder_train_step_matches_cpu_port(built):
    from forde_b200 import dual_encoder, train
    from oracle import dual_encoder_cpu as C

    cfg = dual_encoder.config_tiny()
    st = train.create_train_state(cfg, 0, device="cuda")
    # a mixed brain map so all three paths and both gates are exercised
    for tower in ("vision", "text"):
        for l in st.variables["neuron_assignments"][tower].values():
            a = l["stateful_layer"]["assignments"]
            a.copy_(torch.arange(a.numel(), device=a.device) % 3)
    cpu = C.CpuTrainer(cfg, st.variables)
    b = train.synthetic_batch(cfg, 32, pinned=False)
    losses_g, losses_c = [], []
    for i in range(3):
        _, m = train.train_step(st, b)
        lc, gn_c, stats_c = cpu.step(b["image"].float(), b["input_ids"])
        losses_g.append(float(m["loss"].item())); losses_c.append(lc)
        if i == 0:
            gn_g = float(m["grad_norm"].item())
            assert abs(gn_g - gn_c) <= 0.1 * gn_c, (gn_g, gn_c)
    assert abs(losses_g[0] - losses_c[0]) < 5e-2, (losses_g, losses_c)
    assert losses_g[-1] < losses_g[0] and losses_c[-1] < losses_c[0]        # both arms are learning
    assert abs(losses_g[-1] - losses_c[-1]) < 0.25, (losses_g, losses_c)
    sb = st.variables["stats_buffer"]["vision"]["layer_0"]["stateful_layer"]
    assert int(sb["step_count"].item()) == 3 and int(sb["sample_count"].item()) == 3 * 32 * 5
    sc = cpu.variables["stats_buffer"]["vision"]["layer_0"]["stateful_layer"]["neuron_stats"].numpy()
    sg = sb["neuron_stats"].cpu().numpy()
    # act_gdp / act_variance / grad_gdp are smooth in the inputs: compare across arms; the Hoyer columns too
    for col, tol in ((3, 0.05), (4, 0.1), (1, 0.15), (2, 0.05), (0, 0.05)):
        denom = np.abs(sc[:, col]).mean()
        assert np.abs(sg[:, col] - sc[:, col]).mean() <= tol * denom, (col, np.abs(sg[:, col] - sc[:, col]).mean(), denom)


def test_slow_loop_closes_the_loop_on_gpu(built):
    from forde_b200 import dual_encoder, slow_loop, train

    cfg = dual_encoder.config_tiny()
    st = train.create_train_state(cfg, 1, device="cuda")
    b = train.synthetic_batch(cfg, 32, pinned=False)
    for _ in range(4):
        train.train_step(st, b)
    mv = {"stats_buffer": st.variables["stats_buffer"], "neuron_assignments": st.variables["neuron_assignments"]}
    _, mv, diag = slow_loop.slow_loop_step(st.params, mv, slow_loop.SlowLoopConfig(), key=0, step=4)
    assert len(diag["layers"]) == 4
    for name, d in diag["layers"].items():
        assert sum(d["census"]) == 512 and d["features"].shape == (512, 5)
    for tower in mv["stats_buffer"].values():
        for l in tower.values():
            assert all(float(t.float().abs().sum()) == 0.0 for t in l["stateful_layer"].values())
    # the new brain map is live in the next fast step
    _, m = train.train_step(st, b)
    assert np.isfinite(float(m["loss"].item()))


def test_cuda_graph_replay_matches_eager_steps(built):
    """The captured step must be the same computation: identical losses and statistics as eager launches, and it
    must see the slow loop's in-place rewrite of neuron_assignments and reset of stats_buffer."""
    from forde_b200 import dual_encoder, train

    cfg = dual_encoder.config_tiny()
    b = {k: v.cuda() for k, v in train.synthetic_batch(cfg, 32, pinned=False).items()}
    warm = 3
    eager = train.create_train_state(cfg, 5, device="cuda")
    le = [float(train.train_step(eager, b)[1]["loss"].item()) for _ in range(warm + 3)]
    gs = train.create_train_state(cfg, 5, device="cuda")
    p0 = gs.flat_p32.detach().clone()
    g = train.GraphedTrainStep(gs, b, warmup=warm)          # warm-up + capture train on NOTHING (state restored)
    assert gs.step == 0 and torch.equal(gs.flat_p32.detach(), p0) and int(gs.opt_state["count"].item()) == 0
    assert float(gs.opt_state["mu"].abs().sum()) == 0.0
    sb_g = gs.variables["stats_buffer"]["text"]["layer_1"]["stateful_layer"]
    assert int(sb_g["step_count"].item()) == 0 and float(sb_g["moments"].abs().sum()) == 0.0
    lg = [float(g(b)[1]["loss"].item()) for _ in range(warm + 3)]
    assert gs.step == warm + 3
    np.testing.assert_allclose(lg, le, rtol=2e-3, atol=2e-3)     # one optimizer update per replayed batch
    sb_e = eager.variables["stats_buffer"]["text"]["layer_1"]["stateful_layer"]
    assert int(sb_g["step_count"].item()) == int(sb_e["step_count"].item()) == warm + 3
    np.testing.assert_allclose(sb_g["moments"].cpu().numpy(), sb_e["moments"].cpu().numpy(), rtol=2e-2, atol=1e-3)
    # slow-loop side effects between replays: new brain map + zeroed statistics
    for tower in gs.variables["neuron_assignments"].values():
        for l in tower.values():
            l["stateful_layer"]["assignments"].fill_(2)
    for tower in gs.variables["stats_buffer"].values():
        for l in tower.values():
            for t in l["stateful_layer"].values():
                t.zero_()
    l2 = float(g(b)[1]["loss"].item())
    assert int(sb_g["step_count"].item()) == 1
    assert abs(l2 - lg[-1]) > 1e-4 and np.isfinite(l2)       # routing changed, so the loss moved


def test_fit_runs_fast_and_slow_loops_together(built, tmp_path):
    """train.py:320-399 cadence: graph-replayed fast steps, slow loop every N steps reading the in-kernel statistics."""
    from forde_b200 import dual_encoder, train

    cfg = dual_encoder.config_tiny()
    st = train.create_train_state(cfg, 3, device="cuda")
    b = {k: v.cuda() for k, v in train.synthetic_batch(cfg, 32, pinned=False).items()}
    st, hist = train.fit(st, [b] * 7, max_steps=7, slow_loop_interval=3, key=1, plot_dir=str(tmp_path))
    assert len(hist["slow_loop"]) == 2                      # steps 3 and 6
    import glob, os
    assert len(glob.glob(os.path.join(str(tmp_path), "*", "brain_scan_epoch0_step*.png"))) == 8   # 4 layers x 2 loops
    assert len(glob.glob(os.path.join(str(tmp_path), "*", "feature_space_epoch0_step*.png"))) == 8
    for diag in hist["slow_loop"]:
        assert len(diag["layers"]) == 4 and all("census" in d for d in diag["layers"].values())
    assert np.isfinite(hist["loss"][-1][1])
    # statistics restarted after the last slow loop (step 6): nothing accumulated since
    sb = st.variables["stats_buffer"]["vision"]["layer_0"]["stateful_layer"]
    assert int(sb["step_count"].item()) == 0


def test_text_tower_on_its_own_stream_is_the_same_computation(built):
    """dual_encoder.FORDEDualEncoder(concurrent_towers=True) forks the text tower onto a side stream (forward and
    backward) and joins it before the loss / before the optimizer reads the gradients.  Same seed, same batches: the
    losses, gradient norms, parameters and sensing statistics after three steps must equal the one-stream run (up to
    the fp32 accumulation order of the split-K weight-gradient adds)."""
    from forde_b200 import dual_encoder, train

    cfg = dual_encoder.config_tiny()
    b = {k: v.cuda() for k, v in train.synthetic_batch(cfg, 32, pinned=False).items()}
    runs = {}
    for concurrent in (True, False):
        st = train.create_train_state(cfg, 11, device="cuda")
        st.model.concurrent_towers = concurrent
        out = []
        for _ in range(3):
            _, m = train.train_step(st, b)
            out.append((float(m["loss"].item()), float(m["grad_norm"].item())))
        torch.cuda.synchronize()
        sb = st.variables["stats_buffer"]["text"]["layer_1"]["stateful_layer"]
        runs[concurrent] = (out, st.flat_p32.detach().cpu().numpy(), sb["moments"].cpu().numpy(),
                            int(sb["step_count"].item()))
    (o1, p1, m1, c1), (o0, p0, m0, c0) = runs[True], runs[False]
    np.testing.assert_allclose(np.array(o1), np.array(o0), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(p1, p0, rtol=0, atol=2e-5)
    np.testing.assert_allclose(m1, m0, rtol=1e-3, atol=1e-4)
    assert c1 == c0 == 3
