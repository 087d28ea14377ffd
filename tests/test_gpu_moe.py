"""MoE variant of the FFN slot (BASELINE config 4) against the oracle's restatement of src/forde/moe.py.
Bit-exact: top-k indices (given the kernel's own logits, jax.lax.top_k order), bincount, the routing plan.
Tolerances: logits/probs fp32 1e-4; outputs and activation gradients bf16 (rel Frobenius 2e-2); weight gradients
rel Frobenius 2e-2; aux loss 1e-5."""
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


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / (np.linalg.norm(b) + 1e-30))


def make(M, d, H, E, seed=0, router_scale=0.5):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((M, d), device="cuda", generator=g).bfloat16()
    w_r = (torch.randn((d, E), device="cuda", generator=g) * router_scale / d ** 0.5).bfloat16()
    b_r = torch.randn(E, device="cuda", generator=g) * 0.1
    w_up = (torch.randn((E, d, H), device="cuda", generator=g) / d ** 0.5).bfloat16()
    b_up = torch.randn((E, H), device="cuda", generator=g) * 0.1
    w_down = (torch.randn((E, H, d), device="cuda", generator=g) / H ** 0.5).bfloat16()
    b_down = torch.randn((E, d), device="cuda", generator=g) * 0.1
    return x, w_r, b_r, w_up, b_up, w_down, b_down


@pytest.mark.parametrize("M,d,E,k", [(64, 64, 4, 2), (1000, 128, 8, 2), (8192, 512, 8, 2), (300, 64, 16, 4), (5, 64, 4, 1)])
def test_router_topk_histogram_and_plan(ops, M, d, E, k):
    x, w_r, b_r, *_ = make(M, d, 64, E, seed=M)
    usage = torch.zeros(E, device="cuda"); steps = torch.zeros((), device="cuda", dtype=torch.int32)
    r = ops.moe_route(x, w_r, b_r, k, 0.01, expert_usage=usage, step_count=steps)
    torch.cuda.synchronize()
    logits = f32(r["logits"])
    ref_logits = f32(x) @ f32(w_r) + f32(b_r)
    np.testing.assert_allclose(logits, ref_logits, rtol=1e-4, atol=1e-4)
    # top-k: bit-exact against lax.top_k semantics applied to the kernel's own logits
    vals, idx = O.top_k(logits, k)
    np.testing.assert_array_equal(r["topk_idx"].cpu().numpy(), idx)
    np.testing.assert_allclose(f32(r["topk_w"]), O.softmax(vals), rtol=1e-5, atol=1e-6)
    probs = f32(r["probs"])
    np.testing.assert_allclose(probs, O.softmax(logits), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(probs.sum(-1), 1.0, atol=1e-5)                      # tests/test_moe.py:38
    # histogram: bit-exact bincount ; aux loss ; usage accumulation
    aux_ref, counts_ref = O.moe_aux_loss(probs, idx, E, k, 0.01)
    np.testing.assert_array_equal(r["counts"].cpu().numpy(), counts_ref)
    assert abs(float(r["aux_loss"].item()) - float(aux_ref)) < 1e-6
    np.testing.assert_allclose(f32(usage), probs.mean(0), rtol=1e-5, atol=1e-7)
    assert int(steps.item()) == 1
    # plan: 128-aligned segments in expert order, stable (token-ordered) slots, inverse map, block -> expert
    seg = r["seg_offsets"].cpu().numpy()
    pad = (counts_ref + 127) // 128 * 128
    np.testing.assert_array_equal(seg, np.concatenate([[0], np.cumsum(pad)]))
    assert int(r["n_blocks"].item()) == seg[-1] // 128
    slot_of = r["slot_of"].cpu().numpy(); src = r["src_of_slot"].cpu().numpy()
    flat = idx.reshape(-1)
    for e in range(E):
        entries = np.nonzero(flat == e)[0]                       # ascending entry index = token order
        np.testing.assert_array_equal(slot_of[entries], seg[e] + np.arange(len(entries)))
    valid = src >= 0
    assert valid.sum() == M * k
    np.testing.assert_array_equal(slot_of[src[valid]], np.nonzero(valid)[0])
    be = r["block_expert"].cpu().numpy()
    for e in range(E):
        assert np.all(be[seg[e] // 128: seg[e + 1] // 128] == e)
    assert np.all(be[seg[-1] // 128:] == -1)


def test_router_tie_rule_matches_lax_top_k(ops):
    # identical logits everywhere (zero input, zero router weights, equal bias): lower index wins
    E, k = 8, 2
    x = torch.zeros((40, 64), device="cuda", dtype=torch.bfloat16)
    r = ops.moe_route(x, torch.zeros((64, E), device="cuda", dtype=torch.bfloat16), torch.ones(E, device="cuda"), k, 0.01)
    idx = r["topk_idx"].cpu().numpy()
    assert np.all(idx[:, 0] == 0) and np.all(idx[:, 1] == 1)
    # tests/test_moe.py:46-70: hand-made logits [10,2,5,8] -> {0,3}, [1,9,3,4] -> {1,3} (router = identity rows)
    x2 = torch.zeros((2, 64), device="cuda"); x2[0, :4] = torch.tensor([10., 2., 5., 8.]); x2[1, :4] = torch.tensor([1., 9., 3., 4.])
    w = torch.zeros((64, 4), device="cuda"); w[:4, :4] = torch.eye(4)
    r2 = ops.moe_route(x2.bfloat16(), w.bfloat16(), torch.zeros(4, device="cuda"), 2, 0.01)
    i2 = r2["topk_idx"].cpu().numpy()
    assert set(i2[0].tolist()) == {0, 3} and set(i2[1].tolist()) == {1, 3}
    assert i2[0].tolist() == [0, 3] and i2[1].tolist() == [1, 3]          # descending by value


@pytest.mark.parametrize("M,d,H,E,k", [(256, 64, 128, 4, 2), (1000, 128, 256, 8, 2), (2048, 512, 2048, 8, 2)])
def test_moe_layer_forward_backward_parity(ops, M, d, H, E, k):
    from forde_b200.moe_layer import MoEStatefulLayer
    x, w_r, b_r, w_up, b_up, w_down, b_down = make(M, d, H, E, seed=1)
    layer = MoEStatefulLayer(num_experts=E, top_k=k, expert_hidden_dim=H, d_model=d, aux_loss_weight=0.01)
    v = layer.init(0, x)
    p = v["params"]
    p["router"]["router_linear"]["kernel"] = w_r.float().requires_grad_(True)
    p["router"]["router_linear"]["bias"] = b_r.clone().requires_grad_(True)
    p["experts"]["up_proj"]["kernel"] = w_up.float().requires_grad_(True)
    p["experts"]["up_proj"]["bias"] = b_up.clone().requires_grad_(True)
    p["experts"]["down_proj"]["kernel"] = w_down.float().requires_grad_(True)
    p["experts"]["down_proj"]["bias"] = b_down.clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    (out, aux, routing), upd = layer.apply(v, xr, mutable=["stats_buffer"], return_routing=True)
    dout = torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9)).bfloat16()
    (out.float() * dout.float()).sum().add(aux * 3.0).backward()
    torch.cuda.synchronize()
    ref = O.moe_layer_torch(f32(x), f32(w_r), f32(b_r), f32(w_up), f32(b_up), f32(w_down), f32(b_down), k, 0.01,
                            dout=f32(dout), d_aux=3.0)
    # routing must agree for the comparison to be meaningful (fp32 vs fp64 logits: no near-ties at these seeds)
    np.testing.assert_array_equal(routing["top_k_indices"].cpu().numpy(), ref["idx"])
    np.testing.assert_array_equal(routing["expert_counts"].cpu().numpy(), ref["counts"])
    assert abs(float(aux.item()) - ref["aux"]) < 1e-5
    assert rel(f32(out), ref["out"]) < 2e-2
    assert rel(f32(xr.grad), ref["dx"]) < 2e-2
    ex = p["experts"]
    assert rel(f32(ex["up_proj"]["kernel"].grad), ref["dw_up"]) < 2e-2
    assert rel(f32(ex["down_proj"]["kernel"].grad), ref["dw_down"]) < 2e-2
    assert rel(f32(ex["up_proj"]["bias"].grad), ref["db_up"]) < 2e-2
    assert rel(f32(ex["down_proj"]["bias"].grad), ref["db_down"]) < 2e-2
    assert rel(f32(p["router"]["router_linear"]["kernel"].grad), ref["dw_r"]) < 3e-2
    assert rel(f32(p["router"]["router_linear"]["bias"].grad), ref["db_r"]) < 3e-2
    sb = upd["stats_buffer"]
    assert int(sb["step_count"].item()) == 1
    np.testing.assert_allclose(f32(sb["expert_usage"]), ref["probs"].mean(0), rtol=1e-4, atol=1e-6)


def test_moe_expert_without_tokens_gets_zero_gradient(ops):
    """An expert nobody routes to: its segment is empty, the grouped-K dW GEMM must write zeros for it."""
    from forde_b200.moe_layer import MoEStatefulLayer
    M, d, H, E, k = 200, 64, 128, 4, 1
    x, w_r, b_r, w_up, b_up, w_down, b_down = make(M, d, H, E, seed=3)
    b_r = torch.tensor([0.0, 0.0, -1e4, 0.0], device="cuda")        # expert 2 is never selected
    layer = MoEStatefulLayer(num_experts=E, top_k=k, expert_hidden_dim=H, d_model=d)
    v = layer.init(0, x)
    p = v["params"]
    p["router"]["router_linear"]["kernel"] = w_r.float(); p["router"]["router_linear"]["bias"] = b_r
    p["experts"]["up_proj"]["kernel"] = w_up.float().requires_grad_(True)
    p["experts"]["down_proj"]["kernel"] = w_down.float().requires_grad_(True)
    out, aux, routing = layer.apply(v, x, return_routing=True)
    out.float().sum().backward()
    torch.cuda.synchronize()
    assert int(routing["expert_counts"][2].item()) == 0
    assert float(p["experts"]["up_proj"]["kernel"].grad[2].abs().max()) == 0.0
    assert float(p["experts"]["down_proj"]["kernel"].grad[2].abs().max()) == 0.0
    assert float(p["experts"]["up_proj"]["kernel"].grad[0].abs().max()) > 0.0


def test_c4_full_size_sampled_tokens(ops):
    """BASELINE config 4 dims (8 x 1024 tokens, d=512, H=2048, E=8, top-2): the routed-only result for 48 sampled tokens
    against the reference formulation evaluated per token in float64."""
    from forde_b200.moe_layer import MoEStatefulLayer
    M, d, H, E, k = 8192, 512, 2048, 8, 2
    x, w_r, b_r, w_up, b_up, w_down, b_down = make(M, d, H, E, seed=5)
    layer = MoEStatefulLayer(num_experts=E, top_k=k, expert_hidden_dim=H, d_model=d)
    v = layer.init(0, x)
    p = v["params"]
    p["router"]["router_linear"]["kernel"] = w_r.float(); p["router"]["router_linear"]["bias"] = b_r
    p["experts"]["up_proj"]["kernel"] = w_up.float(); p["experts"]["up_proj"]["bias"] = b_up
    p["experts"]["down_proj"]["kernel"] = w_down.float(); p["experts"]["down_proj"]["bias"] = b_down
    out, aux, routing = layer.apply(v, x.view(8, 1024, d), return_routing=True)
    torch.cuda.synchronize()
    assert out.shape == (8, 1024, d)
    idx = routing["top_k_indices"].reshape(M, k).cpu().numpy()
    counts = routing["expert_counts"].cpu().numpy()
    np.testing.assert_array_equal(counts, np.bincount(idx.reshape(-1), minlength=E))           # histogram exact
    assert counts.sum() == M * k
    probs = f32(routing["router_probs"]).reshape(M, E)
    aux_ref, _ = O.moe_aux_loss(probs, idx, E, k, 0.01)
    assert abs(float(aux.item()) - float(aux_ref)) < 1e-6
    xs, wu, wd = f32(x).astype(np.float64), f32(w_up).astype(np.float64), f32(w_down).astype(np.float64)
    logits = xs @ f32(w_r).astype(np.float64) + f32(b_r)
    o = f32(out).reshape(M, d)
    for m in np.random.RandomState(0).choice(M, 48, replace=False):
        sel = np.argsort(-logits[m], kind="stable")[:k]
        np.testing.assert_array_equal(idx[m], sel)
        w = np.exp(logits[m, sel] - logits[m, sel].max()); w /= w.sum()
        ref = sum(w[j] * (O.gelu_tanh((xs[m] @ wu[e] + f32(b_up)[e]).astype(np.float32)).astype(np.float64) @ wd[e] + f32(b_down)[e])
                  for j, e in enumerate(sel))
        assert rel(o[m], ref) < 2e-2


def test_moe_slow_loop_closes_the_loop_on_gpu(ops):
    """Kernel-side expert_usage / step_count feed moe_slow_loop_step; its in-place router-bias nudges are what the next
    forward routes with (moe_slow_loop.py:418-439), and the statistics restart from zero (:441-449)."""
    from forde_b200.moe_layer import MoEStatefulLayer
    from forde_b200 import slow_loop as SL
    M, d, H, E, k = 512, 64, 128, 4, 2
    x, *_ = make(M, d, H, E, seed=8)
    layer = MoEStatefulLayer(num_experts=E, top_k=k, expert_hidden_dim=H, d_model=d)
    v = layer.init(0, x)
    v["params"]["router"]["router_linear"]["bias"].copy_(torch.tensor([1.0, 0.0, -0.5, 0.0], device="cuda"))
    params = {"params": {"layer_0": {"moe": v["params"]["router"]}}}
    mv = {"stats_buffer": {"layer_0": {"moe": v["stats_buffer"]}}}
    for _ in range(3):
        layer.apply(v, x, mutable=["stats_buffer"])
    torch.cuda.synchronize()
    assert int(v["stats_buffer"]["step_count"].item()) == 3
    usage = f32(v["stats_buffer"]["expert_usage"]) / 3

    class Cfg:
        num_layers, num_experts = 1, E
    bias_before = v["params"]["router"]["router_linear"]["bias"].clone()
    _, _, diag = SL.moe_slow_loop_step(params, mv, Cfg, key=3, step=100)
    np.testing.assert_allclose(diag["expert_usage"], usage, rtol=1e-5)
    delta = f32(v["params"]["router"]["router_linear"]["bias"] - bias_before)
    np.testing.assert_allclose(delta, diag["adjustments"], rtol=1e-5, atol=1e-7)
    assert delta[0] < 0 and delta[2] > 0                       # over-used expert pushed down, under-used pushed up
    assert int(v["stats_buffer"]["step_count"].item()) == 0 and float(v["stats_buffer"]["expert_usage"].abs().sum()) == 0.0


def test_decoder_lm_with_moe_slot_trains_and_feeds_the_slow_loop(ops):
    """Config-4 shaped caller in miniature: decoder LM (causal attention + MoEStatefulLayer in the FFN slot), a few
    AdamW steps on synthetic tokens (src/data/dataset.py:27 style), first-step loss against an fp32 dense-all-experts
    restatement, then moe_slow_loop_step on the model's own variable trees."""
    from forde_b200 import llm_harness as LM, slow_loop as SL
    cfg = LM.LLMConfig(vocab_size=512, d_model=64, num_layers=2, num_heads=2, head_dim=32, max_seq_len=64,
                       num_experts=4, top_k_experts=2, expert_hidden_dim=128)
    model = LM.FORDEDecoderLM(cfg)
    v = model.init(0, "cuda")
    ids = torch.from_numpy(np.random.RandomState(42).randint(0, cfg.vocab_size, (4, 64)).astype(np.int32)).cuda()
    leaves = list(LM.param_leaves(v["params"]))
    opt = torch.optim.AdamW(leaves, lr=3e-3)

    # fp32 restatement of the same forward (every expert on every token, moe.py:163) for the first-step loss
    def ref_loss():
        p = v["params"]
        f = lambda t: t.detach().float()
        B, S = ids.shape
        x = torch.nn.functional.embedding(ids.long(), f(p["token_embed"])) + f(p["pos_embed"])[:S]
        aux_t = 0.0
        for i in range(cfg.num_layers):
            lp = p[f"layer_{i}"]
            h = torch.nn.functional.layer_norm(x, (64,), f(lp["attn_norm"]["scale"]), f(lp["attn_norm"]["bias"]), 1e-6)
            q, k, vv = (h @ f(lp["attn"]["qkv"])).view(B, S, 3, 2, 32).permute(2, 0, 3, 1, 4)
            a = torch.nn.functional.scaled_dot_product_attention(q, k, vv, is_causal=True).transpose(1, 2).reshape(B, S, -1)
            x = x + a @ f(lp["attn"]["out"])
            h = torch.nn.functional.layer_norm(x, (64,), f(lp["ffn_norm"]["scale"]), f(lp["ffn_norm"]["bias"]), 1e-6)
            mp = lp["moe"]
            r = O.moe_layer_torch(h.reshape(-1, 64).cpu().numpy(), f(mp["router"]["router_linear"]["kernel"]).cpu().numpy(),
                                  f(mp["router"]["router_linear"]["bias"]).cpu().numpy(), f(mp["experts"]["up_proj"]["kernel"]).cpu().numpy(),
                                  f(mp["experts"]["up_proj"]["bias"]).cpu().numpy(), f(mp["experts"]["down_proj"]["kernel"]).cpu().numpy(),
                                  f(mp["experts"]["down_proj"]["bias"]).cpu().numpy(), 2, 0.01)
            x = x + torch.from_numpy(r["out"]).float().cuda().view(B, S, 64)
            aux_t += r["aux"]
        x = torch.nn.functional.layer_norm(x, (64,), f(p["final_norm"]["scale"]), f(p["final_norm"]["bias"]), 1e-6)
        logits = x @ f(p["lm_head"]["kernel"]) + f(p["lm_head"]["bias"])
        return float(LM.lm_loss(logits, ids, aux_t)[0])

    expected = ref_loss()
    losses = []
    for step in range(6):
        (logits, aux), _ = model.apply(v, ids, mutable=["stats_buffer"])
        loss, lm = LM.lm_loss(logits, ids, aux)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(float(loss.item()))
    assert abs(losses[0] - expected) < 5e-2, (losses[0], expected)
    assert losses[-1] < losses[0] - 0.3                                        # it learns the fixed batch
    sb = v["stats_buffer"]["layer_1"]["moe"]
    assert int(sb["step_count"].item()) == 6 and abs(float(sb["expert_usage"].sum().item()) - 6.0) < 1e-3

    class Cfg:
        num_layers, num_experts = cfg.num_layers, cfg.num_experts
    bias0 = v["params"]["layer_0"]["moe"]["router"]["router_linear"]["bias"].detach().clone()
    _, _, diag = SL.moe_slow_loop_step({"params": v["params"]}, {"stats_buffer": v["stats_buffer"]}, Cfg, key=1, step=6)
    assert diag["router_biases_updated"] == cfg.num_layers and diag["step_count"] == 6
    assert int(sb["step_count"].item()) == 0
    delta = v["params"]["layer_0"]["moe"]["router"]["router_linear"]["bias"].detach() - bias0
    np.testing.assert_allclose(delta.cpu().numpy(), diag["adjustments"], rtol=1e-5, atol=1e-7)
