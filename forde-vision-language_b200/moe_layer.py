"""MoEStatefulLayer -- the LLM-variant occupant of the FFN slot (BASELINE config 4).

Reference: src/forde/moe.py:225-281 (MoEStatefulLayer wrapping MoELayer, moe.py:57-222).  Interface kept:
  * ``module.apply(variables, x, mutable=["stats_buffer"]) -> ((output, aux_loss), updated)``; without ``mutable`` it
    returns ``(output, aux_loss)`` (moe.py:281; call site src/forde/model.py:165-173);
  * collections: ``params/router/router_linear/{kernel (d,E), bias (E)}`` (moe.py:49-53: normal(0.02) kernel), expert
    weights, and ``stats_buffer/{expert_usage (E,) f32, step_count i32}`` (moe.py:265-279) which the slow loop finds by
    key substring and zeroes (src/forde/moe_slow_loop.py:145-170,441-449); its router-bias nudges
    (moe_slow_loop.py:418-439) write ``router_linear/bias`` in place.
B200 layout difference: the E experts' weights are STACKED (``params/experts/up_proj/kernel`` is (E, d, H), etc.) so one
grouped tcgen05 GEMM serves all experts; ``expert_view(params, i)`` gives the reference-shaped ``expert_{i}`` dict of
views.  Only the routed (token, expert) pairs are computed -- the reference evaluates every expert on every token and
gathers (moe.py:163), which is result-equivalent.

Kernels (csrc/forde_moe.cu + the grouped modes of csrc/forde_gemm.cuh): router / top-k / histogram / aux loss /
usage statistics / routing plan, gather, grouped up-projection with fused bias + tanh-GELU, grouped down-projection,
weighted combine; backward: scatter, gate gradient, grouped dz (fused GELU'), dx, per-expert dW (grouped-K) and db.
router backward (softmax Jacobians, dx += dlogits W_r^T, dW_r, db_r).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Iterable, Union

import torch

from . import ops


class _MoEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, w_r, b_r, w_up, b_up, w_down, b_down, layer, stats):
        k, aux_w = layer.top_k, layer.aux_loss_weight
        w_r16 = w_r.detach().to(torch.bfloat16).contiguous()
        w_up16 = w_up.detach().to(torch.bfloat16).contiguous()
        w_down16 = w_down.detach().to(torch.bfloat16).contiguous()
        f = lambda t: t.detach().float().contiguous()
        route = ops.moe_route(x2d, w_r16, f(b_r), k, aux_w,
                              expert_usage=stats["expert_usage"] if stats is not None else None,
                              step_count=stats["step_count"] if stats is not None else None)
        xg = ops.moe_gather(x2d, route)
        zg, hg, yg = ops.moe_experts_fwd(xg, w_up16, f(b_up), w_down16, f(b_down), route)
        out = ops.moe_combine(yg, route, weighted=True)
        ctx.route, ctx.layer = route, layer
        ctx.save_for_backward(x2d, w_r16, w_up16, w_down16, xg, zg, hg, yg)
        ctx.dtypes = tuple(t.dtype for t in (w_r, b_r, w_up, b_up, w_down, b_down))
        ctx.mark_non_differentiable(route["topk_idx"], route["counts"])
        return out, route["aux_loss"].reshape(()), route["probs"], route["topk_idx"], route["counts"]

    @staticmethod
    def backward(ctx, dout, d_aux, d_probs, _di, _dc):
        x2d, w_r16, w_up16, w_down16, xg, zg, hg, yg = ctx.saved_tensors
        r = ctx.route
        M, E, k = r["M"], r["E"], r["k"]
        dout = dout.contiguous().to(torch.bfloat16)
        dyg = ops.moe_scatter(dout, r)
        d_w = ops.moe_gate_grad(dout, yg, r)
        dxg, dw_up, db_up, dw_down, db_down = ops.moe_experts_bwd(xg, dyg, w_up16, w_down16, zg, hg, r)
        dx = ops.moe_combine(dxg, r, weighted=False)
        # ---- router backward kernel: top-k softmax Jacobian + the aux-loss path through probs (moe.py:212-222);
        #      adds dlogits . W_r^T into dx in place
        da = d_aux.detach().float().reshape(1).contiguous() if d_aux is not None else None
        dp = d_probs.detach().float().contiguous() if d_probs is not None else None
        dw_r, db_r = ops.moe_router_bwd(x2d, w_r16, r, d_w, da, dp, ctx.layer.aux_loss_weight, dx)
        dt = ctx.dtypes
        return (dx, dw_r.to(dt[0]), db_r.to(dt[1]), dw_up.to(dt[2]), db_up.to(dt[3]), dw_down.to(dt[4]),
                db_down.to(dt[5]), None, None)


@dataclass
class MoEStatefulLayer:
    """Drop-in for src/forde/moe.py:225 (same dataclass fields and defaults, moe.py:237-241)."""

    num_experts: int = 8
    top_k: int = 2
    expert_hidden_dim: int = 2048
    d_model: int = 512
    aux_loss_weight: float = 0.01
    name: str = "moe"

    def init(self, key: Union[int, torch.Generator], x: torch.Tensor) -> Dict[str, dict]:
        d, E, H = x.shape[-1], self.num_experts, self.expert_hidden_dim
        dev = x.device
        gen = key if isinstance(key, torch.Generator) else torch.Generator(device="cpu").manual_seed(int(key))

        def lecun(fan_in, shape):
            w = torch.empty(shape, dtype=torch.float32)
            torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
            return (w * (math.sqrt(1.0 / fan_in) / 0.87962566103423978)).to(dev)

        return {
            "params": {
                "router": {"router_linear": {"kernel": (torch.randn((d, E), generator=gen) * 0.02).to(dev),
                                             "bias": torch.zeros(E, device=dev)}},
                "experts": {"up_proj": {"kernel": lecun(d, (E, d, H)), "bias": torch.zeros((E, H), device=dev)},
                            "down_proj": {"kernel": lecun(H, (E, H, d)), "bias": torch.zeros((E, d), device=dev)}},
            },
            "stats_buffer": {"expert_usage": torch.zeros(E, dtype=torch.float32, device=dev),
                             "step_count": torch.zeros((), dtype=torch.int32, device=dev)},
        }

    @staticmethod
    def expert_view(params: Dict[str, dict], i: int) -> Dict[str, dict]:
        """The reference-shaped parameters of expert i (views of the stacked arrays): expert_{i}/{up,down}_proj."""
        e = params["experts"]
        return {"up_proj": {"kernel": e["up_proj"]["kernel"][i], "bias": e["up_proj"]["bias"][i]},
                "down_proj": {"kernel": e["down_proj"]["kernel"][i], "bias": e["down_proj"]["bias"][i]}}

    @staticmethod
    def from_reference_params(tree) -> Dict[str, dict]:
        """Convert the reference's Flax parameter tree of one MoEStatefulLayer into this module's ``params``.

        The reference nests ``MoELayer_0/{MoERouter_0/router_linear, expert_0 .. expert_{E-1}}/...`` with one
        ``{up_proj, down_proj}/{kernel, bias}`` per expert (src/forde/moe.py:17-29,88,96-104,254-260); here the experts
        are stacked along a leading axis.  Any nesting above ``router_linear`` / ``expert_{i}`` is accepted; leaves may
        be torch tensors or array-likes (NumPy / JAX arrays of a loaded checkpoint)."""
        found = {"router": None, "experts": {}}

        def walk(node):
            if not isinstance(node, dict):
                return
            for k, v in node.items():
                if k == "router_linear":
                    found["router"] = v
                elif k.startswith("expert_") and k[7:].isdigit():
                    found["experts"][int(k[7:])] = v
                else:
                    walk(v)

        walk(tree)
        E = len(found["experts"])
        if found["router"] is None or E == 0 or sorted(found["experts"]) != list(range(E)):
            raise ValueError("not a MoEStatefulLayer parameter tree: need router_linear and expert_0 .. expert_{E-1}")
        t = lambda a: a if torch.is_tensor(a) else torch.as_tensor(__import__("numpy").asarray(a))
        stack = lambda proj, leaf: torch.stack([t(found["experts"][i][proj][leaf]) for i in range(E)])
        return {"router": {"router_linear": {"kernel": t(found["router"]["kernel"]), "bias": t(found["router"]["bias"])}},
                "experts": {"up_proj": {"kernel": stack("up_proj", "kernel"), "bias": stack("up_proj", "bias")},
                            "down_proj": {"kernel": stack("down_proj", "kernel"), "bias": stack("down_proj", "bias")}}}

    @staticmethod
    def to_reference_params(params: Dict[str, dict]) -> Dict[str, dict]:
        """The inverse: this module's ``params`` as the reference's nested tree (views of the stacked arrays)."""
        E = params["experts"]["up_proj"]["kernel"].shape[0]
        inner = {"MoERouter_0": {"router_linear": dict(params["router"]["router_linear"])}}
        for i in range(E):
            inner[f"expert_{i}"] = MoEStatefulLayer.expert_view(params, i)
        return {"MoELayer_0": inner}

    def apply(self, variables, x: torch.Tensor, mutable: Union[bool, Iterable[str]] = False, return_routing=False):
        p = variables["params"]
        mut = set(mutable) if not isinstance(mutable, bool) else (set(variables) - {"params"} if mutable else set())
        stats = variables.get("stats_buffer") if "stats_buffer" in mut else None
        lead = x.shape[:-1]
        x2d = x.reshape(-1, x.shape[-1])
        if x2d.dtype != torch.bfloat16:
            x2d = x2d.to(torch.bfloat16)
        x2d = x2d.contiguous()
        rl, ex = p["router"]["router_linear"], p["experts"]
        out, aux, probs, idx, counts = _MoEFn.apply(x2d, rl["kernel"], rl["bias"], ex["up_proj"]["kernel"],
                                                    ex["up_proj"]["bias"], ex["down_proj"]["kernel"],
                                                    ex["down_proj"]["bias"], self, stats)
        res = (out.reshape(*lead, out.shape[-1]), aux)
        if return_routing:
            res = res + ({"router_probs": probs.reshape(*lead, -1), "top_k_indices": idx.reshape(*lead, -1),
                          "expert_counts": counts},)
        if isinstance(mutable, bool) and not mutable:
            return res
        return res, {k: variables[k] for k in mut if k in variables}

    __call__ = apply
