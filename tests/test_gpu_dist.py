"""NCCL path on real GPUs (needs >= 2 devices; skipped on a single-GPU box): the two-rank contrastive loss and one
data-parallel train step must equal the single-process result on the concatenated batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import forde_oracle as O

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import forde_b200  # noqa: F401
        from forde_b200 import dual_encoder, train
        from forde_b200.clip_loss import clip_contrastive_loss
        dev = torch.device("cuda", rank)
        # ---- contrastive loss over the global batch
        rng = np.random.RandomState(0)
        B, E = 64, 128
        img = torch.from_numpy(rng.randn(B, E).astype(np.float32)).bfloat16().float()
        txt = torch.from_numpy((rng.randn(B, E) + 0.5 * img.numpy()).astype(np.float32)).bfloat16().float()
        Bl = B // world
        il = img[rank * Bl:(rank + 1) * Bl].to(dev).requires_grad_(True)
        tl = txt[rank * Bl:(rank + 1) * Bl].to(dev).requires_grad_(True)
        lt = torch.tensor(2.0, device=dev, requires_grad=True)
        loss = clip_contrastive_loss(il, tl, lt)
        loss.backward()
        dlt = lt.grad.clone(); dist.all_reduce(dlt)
        # ---- one data-parallel train step of the tiny dual encoder (global batch 32 split over the ranks)
        cfg = dual_encoder.config_tiny()
        st = train.create_train_state(cfg, 0, device=dev)
        b = train.synthetic_batch(cfg, 32, pinned=False)
        sl = slice(rank * 16, (rank + 1) * 16)
        _, m = train.train_step(st, {k: v[sl].to(dev) for k, v in b.items()})
        torch.cuda.synchronize()
        q.put((rank, float(loss.item()), il.grad.cpu().numpy(), tl.grad.cpu().numpy(), float(dlt.item()),
               float(m["loss"].item()), float(m["grad_norm"].item()), st.flat_p32[:4096].cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_nccl_clip_loss_and_train_step():
    import __graft_entry__ as g
    g.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rng = np.random.RandomState(0)
    B, E = 64, 128
    img = O.bf16_round(rng.randn(B, E).astype(np.float32))
    txt = O.bf16_round((rng.randn(B, E) + 0.5 * img).astype(np.float32))
    ref_loss = float(O.clip_loss_fwd(img, txt, 2.0)[0])
    dimg, dtxt, dlt, _ = O.clip_loss_bwd(img, txt, 2.0)
    for rank, loss, gi, gt, g_lt, *_ in res:
        assert abs(loss - ref_loss) < 1e-2
        sl = slice(rank * 32, (rank + 1) * 32)
        assert np.linalg.norm(gi - dimg[sl]) / np.linalg.norm(dimg[sl]) < 3e-2
        assert np.linalg.norm(gt - dtxt[sl]) / np.linalg.norm(dtxt[sl]) < 3e-2
        assert abs(g_lt - float(dlt)) < 3e-2 * max(1.0, abs(float(dlt)))
    # both ranks took the same optimizer step: identical loss, gradient norm and parameters
    assert res[0][5] == res[1][5] and abs(res[0][6] - res[1][6]) < 1e-4 * res[0][6]
    np.testing.assert_array_equal(res[0][7], res[1][7])
    # and it is the single-process step on the whole batch of 32 (same init, same data)
    from forde_b200 import dual_encoder, train
    cfg = dual_encoder.config_tiny()
    st = train.create_train_state(cfg, 0, device="cuda")
    b = {k: v.cuda() for k, v in train.synthetic_batch(cfg, 32, pinned=False).items()}
    _, m = train.train_step(st, b)
    assert abs(float(m["loss"].item()) - res[0][5]) < 2e-2
    assert abs(float(m["grad_norm"].item()) - res[0][6]) < 5e-2 * res[0][6]
