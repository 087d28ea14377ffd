"""World-size-2 `gloo` tests of the N > 1 host logic (runs on CPU): the data-parallel formulation of the
contrastive loss (embedding all-gather, column log-sum-exp merged from per-rank (max, sum-exp) pairs,
reduce-scatter of d(txt)), the gradient SUM all-reduce, and the pooled-moment all-reduce of the slow loop.

The CUDA kernels cannot run here, so the ops layer is replaced by NumPy emulations written in this file (test
side only -- the product never falls back to them); what is under test is the collective plumbing and that the
2-rank result equals the single-process oracle on the concatenated batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import forde_oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


# ---- NumPy emulation of the clip kernels' C-ABI semantics (include/forde_b200.h) on CPU tensors
class _FakeOps:
    @staticmethod
    def clip_fwd(img, txt_all, log_temp, row_offset):
        i = img.float().numpy(); t = txt_all.float().numpy()
        rn_i = 1.0 / np.linalg.norm(i, axis=1); rn_t = 1.0 / np.linalg.norm(t, axis=1)
        u = i * rn_i[:, None]; v = t * rn_t[:, None]
        scale = np.exp(log_temp.numpy().astype(np.float64))
        S = (scale * (u @ v.T)).astype(np.float32)
        Bl = i.shape[0]
        m = S.max(1); row_lse = m + np.log(np.exp(S - m[:, None]).sum(1))
        cm = S.max(0); cs = np.exp(S - cm[None, :]).sum(0)
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        return {"u": f(u), "v": f(v), "rn_i": f(rn_i), "rn_t": f(rn_t), "scale": f(np.array([scale])), "S": f(S),
                "row_lse": f(row_lse), "col_max": f(cm), "col_sum": f(cs),
                "diag": f(S[np.arange(Bl), row_offset + np.arange(Bl)])}

    @staticmethod
    def clip_col_rescale(col_max_local, col_max_global, col_sum):
        col_sum.mul_(torch.exp(col_max_local - col_max_global))

    @staticmethod
    def clip_loss(fwd, col_max, col_sum, row_offset):
        col_lse = col_max + torch.log(col_sum)
        Bl, Bg = fwd["row_lse"].numel(), col_max.numel()
        d = fwd["diag"]
        share = ((fwd["row_lse"] - d).sum() + (col_lse[row_offset:row_offset + Bl] - d).sum()) * 0.5 / Bg
        return share.reshape(1), col_lse

    @staticmethod
    def clip_bwd(fwd, col_lse, row_offset):
        S = fwd["S"]; Bl, Bg = S.shape
        w = 0.5 / Bg
        dS = (torch.exp(S - fwd["row_lse"][:, None]) + torch.exp(S - col_lse[None, :])) * w
        dS[torch.arange(Bl), row_offset + torch.arange(Bl)] -= 2 * w
        scale = fwd["scale"]
        return scale * dS @ fwd["v"], scale * dS.t() @ fwd["u"], (dS * S).sum().reshape(1)

    @staticmethod
    def l2norm_bwd(u, inv_norm, dU):
        return inv_norm[:, None] * (dU - u * (u * dU).sum(1, keepdim=True))

    @staticmethod
    def pooled_moments(stats_leaf):
        return stats_leaf["moments"].double() + stats_leaf["moments_lo"].double()

    @staticmethod
    def stats_from_moments(moments, sample_count, out=None):
        assert moments.dtype == torch.float64          # the cross-rank SUM of the accumulators runs in fp64
        n = int(sample_count.reshape(-1)[0])
        res = torch.from_numpy(O.stats_from_moments(moments.numpy(), n)) if n > 0 else torch.zeros((moments.shape[0], 5))
        if out is not None:
            out.copy_(res)
            return out
        return res


def _clip_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import forde_b200  # noqa: F401
        from forde_b200 import clip_loss as CL
        CL.ops = _FakeOps                      # the kernels' stand-in (test side)
        rng = np.random.RandomState(0)
        B, E = 12, 16
        img = torch.from_numpy(rng.randn(B, E).astype(np.float32)).bfloat16().float()
        txt = torch.from_numpy((rng.randn(B, E) + 0.5 * img.numpy()).astype(np.float32)).bfloat16().float()
        Bl = B // world
        il = img[rank * Bl:(rank + 1) * Bl].clone().requires_grad_(True)
        tl = txt[rank * Bl:(rank + 1) * Bl].clone().requires_grad_(True)
        lt = torch.tensor(1.7, requires_grad=True)
        loss = CL.clip_contrastive_loss(il, tl, lt)
        loss.backward()
        dlt = lt.grad.clone()
        dist.all_reduce(dlt)                   # parameter gradients are SUM-reduced by the training loop
        q.put((rank, float(loss), il.grad.numpy(), tl.grad.numpy(), float(dlt)))
    finally:
        dist.destroy_process_group()


def test_two_rank_clip_loss_equals_global_batch_oracle():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_clip_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.RandomState(0)
    B, E = 12, 16
    img = O.bf16_round(rng.randn(B, E).astype(np.float32))
    txt = O.bf16_round((rng.randn(B, E) + 0.5 * img).astype(np.float32))
    ref_loss = float(O.clip_loss_fwd(img, txt, 1.7)[0])
    dimg, dtxt, dlt, _ = O.clip_loss_bwd(img, txt, 1.7)
    for rank, loss, gi, gt, g_lt in res:
        assert abs(loss - ref_loss) < 1e-5                      # every rank returns the GLOBAL loss
        np.testing.assert_allclose(gi, dimg[rank * 6:(rank + 1) * 6], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(gt, dtxt[rank * 6:(rank + 1) * 6], rtol=1e-4, atol=1e-6)
        assert abs(g_lt - float(dlt)) < 1e-5


def _grad_and_moments_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import forde_b200  # noqa: F401
        from forde_b200 import slow_loop as SL, train
        from forde_b200.stateful_layer import new_stats_buffer
        SL.ops = _FakeOps
        g = torch.full((10,), float(rank + 1))
        train.sync_gradients(g)                                  # SUM, not mean
        rng = np.random.RandomState(7)
        a = (rng.randn(64, 8) + 0.3).astype(np.float32); gr = (rng.randn(64, 8) * 1e-2).astype(np.float32)
        sl = slice(rank * 32, (rank + 1) * 32)                   # each rank sensed its own half of the samples
        m, n = O.raw_moments(a[sl], gr[sl])
        sb = new_stats_buffer(8, "cpu")
        hi = m.astype(np.float32)                                # double-single accumulator: high part + residue
        sb["moments"] += torch.from_numpy(hi); sb["moments_lo"] += torch.from_numpy((m - hi).astype(np.float32))
        sb["sample_count"] += n; sb["step_count"] += 1
        stats = SL.sense_layer(sb, pooled=True)
        q.put((rank, g.numpy(), stats))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_sum_and_pooled_moments():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_and_moments_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.RandomState(7)
    a = (rng.randn(64, 8) + 0.3).astype(np.float32); gr = (rng.randn(64, 8) * 1e-2).astype(np.float32)
    ref = O.calculate_neuron_stats(a, gr)                        # statistics of the POOLED samples of both ranks
    for rank, g, stats in res:
        np.testing.assert_array_equal(g, np.full(10, 3.0, np.float32))
        np.testing.assert_allclose(stats, ref, rtol=1e-4, atol=1e-6)


def _bucket_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import forde_b200  # noqa: F401
        from forde_b200 import train
        tree = {"vision": {"patch_embed": {"kernel": torch.ones(70)},
                           "layer_0": {"w": torch.ones(100)}, "layer_5": {"w": torch.ones(33)},
                           "proj": {"kernel": torch.ones(10)}},
                "text": {"token_embed": torch.ones(20), "layer_0": {"w": torch.ones(64)}},
                "log_temp": torch.ones(())}
        leaves = list(train._leaves(tree))
        offs, total = [], 0
        for _, t in leaves:
            offs.append(total); total += (t.numel() + 63) // 64 * 64
        flat_p = torch.zeros(total); flat_g = torch.zeros(total)
        params = []
        for (path, t), off in zip(leaves, offs):
            p = flat_p[off:off + t.numel()].view(t.shape)
            p.copy_(t); p.requires_grad_(True); p.grad = flat_g[off:off + t.numel()].view(t.shape)
            params.append(p)
        b = train._GradBuckets(leaves, offs, total, flat_g, None)
        b.attach(params)
        assert [r[0] for r in b.ranges] == [("vision", -1), ("vision", 0), ("vision", 1), ("vision", 10 ** 6),
                                            ("text", -1), ("text", 0), ("log_temp", 10 ** 6)]
        for step in range(2):                                   # counters must re-arm every step
            flat_g.zero_()
            b.begin()
            used = params[:-2] + params[-1:]                    # text/layer_0 gets no gradient: finish() must cover it
            # params[1] (vision/layer_0) takes the gradient-sink route: its gradient is written into the flat view by
            # "a kernel", announced by hand, and autograd's hook then fires once more for it (with a zero gradient
            # here) -- it must be counted once
            loss = sum((0.0 if i == 1 else (rank + 1) * (i + 1)) * p.sum() for i, p in enumerate(used))
            with torch.no_grad():
                params[1].grad.add_(float((rank + 1) * 2))
            params[1]._forde_ready(params[1])
            loss.backward()
            b.finish()
        q.put((rank, flat_g.clone().numpy(), [(o, t.numel()) for (_, t), o in zip(leaves, offs)]))
    finally:
        dist.destroy_process_group()


def test_two_rank_bucketed_gradient_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g, layout in res:
        used = layout[:-2] + layout[-1:]
        for i, (off, n) in enumerate(used):
            np.testing.assert_array_equal(g[off:off + n], np.full(n, 3.0 * (i + 1), np.float32))   # (1 + 2) * (i+1)
        off, n = layout[-2]
        np.testing.assert_array_equal(g[off:off + n], np.zeros(n, np.float32))


def _slow_loop_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import forde_b200  # noqa: F401
        from forde_b200 import slow_loop as SL
        from forde_b200.stateful_layer import new_stats_buffer
        SL.ops = _FakeOps
        if rank != 0:                      # with broadcast only rank 0 may cluster; anything else is a bug
            def boom(*a, **k):
                raise AssertionError("slow_loop_compute ran on a rank other than 0")
            SL.slow_loop_compute = boom
        rng = np.random.RandomState(3)
        H = 64
        z = (rng.randn(256, H) * np.linspace(0.2, 2.0, H)).astype(np.float32)
        g = (rng.randn(256, H) * np.where(np.arange(H) % 3 == 0, 1.0, 0.05)).astype(np.float32)
        sl = slice(rank * 128, (rank + 1) * 128)                 # each rank sensed its own half of the samples
        m, n = O.raw_moments(z[sl], g[sl])
        mv = {"stats_buffer": {"layer_0": {"stateful_layer": new_stats_buffer(H, "cpu")}},
              "neuron_assignments": {"layer_0": {"stateful_layer": {"assignments": torch.zeros(H, dtype=torch.int32)}}}}
        sb = mv["stats_buffer"]["layer_0"]["stateful_layer"]
        hi = m.astype(np.float32)
        sb["moments"] += torch.from_numpy(hi); sb["moments_lo"] += torch.from_numpy((m - hi).astype(np.float32))
        sb["sample_count"] += n; sb["step_count"] += 1
        _, mv, diag = SL.slow_loop_step({}, mv, SL.SlowLoopConfig(), key=5, step=100)
        a = mv["neuron_assignments"]["layer_0"]["stateful_layer"]["assignments"].numpy().copy()
        d = diag["layers"]["layer_0/stateful_layer"]
        q.put((rank, a, d["census"], d["raw_census"] is not None, float(sb["moments"].abs().sum())))
    finally:
        dist.destroy_process_group()


def test_two_rank_slow_loop_rank0_clusters_and_broadcasts_the_maps():
    """Sense pools both ranks' moments; only rank 0 runs Cluster + Smooth; every rank actuates rank 0's map; Reset."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slow_loop_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, a0, c0, raw0, m0), (_, a1, c1, raw1, m1) = res
    np.testing.assert_array_equal(a0, a1)
    assert c0 == c1 and sum(c0) == 64 and len(set(a0.tolist())) > 1      # a non-trivial map, identical on both ranks
    assert raw0 is True and raw1 is False                                # only rank 0 holds the GMM's own labels
    assert m0 == 0.0 and m1 == 0.0                                       # Reset on both
