"""Decoder-only LM harness around the MoE variant of the FFN slot (BASELINE config 4) -- a CALLER, not the product.

The reference's live model (src/forde/model.py:202-294, FORDEDecoderLM) with its optional subsystems switched off
exactly as its CLI allows (``--no_nsa --no_mhc``, src/training/train.py:228-234): token + learned position embedding
(model.py:235-248) -> L x [LayerNorm -> causal self-attention -> residual -> LayerNorm -> MoEStatefulLayer -> residual]
(model.py:120-197 with plain residuals) -> final LayerNorm -> lm_head (model.py:276-292); loss = next-token cross
entropy + the summed MoE auxiliary losses (train.py:141-153).  NSA attention, the mHC hyper-connections and dropout are
outside the hot path and are not built.

Only the FFN slot runs on this repository's kernels (forde_b200.moe_layer); attention is torch SDPA, the rest torch ops.
Variables: ``params``, and ``stats_buffer/layer_{i}/moe/{expert_usage, step_count}`` -- the tree shape
``moe_slow_loop_step`` walks (src/forde/moe_slow_loop.py:145-170; tests/verify_reset.py:17-35).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F

from .moe_layer import MoEStatefulLayer


@dataclass
class LLMConfig:
    """The fields of src/forde/model.py:39-69 that this harness uses (same names and defaults)."""
    vocab_size: int = 32000
    d_model: int = 512
    num_layers: int = 12
    num_heads: int = 8
    head_dim: int = 64
    max_seq_len: int = 2048
    num_experts: int = 8
    top_k_experts: int = 2
    expert_hidden_dim: int = 2048
    moe_aux_loss_weight: float = 0.01


class FORDEDecoderLM:
    def __init__(self, config: LLMConfig):
        self.config = config
        self.moe = [MoEStatefulLayer(num_experts=config.num_experts, top_k=config.top_k_experts,
                                     expert_hidden_dim=config.expert_hidden_dim, d_model=config.d_model,
                                     aux_loss_weight=config.moe_aux_loss_weight, name=f"layer_{i}/moe")
                    for i in range(config.num_layers)]

    def init(self, key: int, device) -> Dict[str, dict]:
        c = self.config
        gen = torch.Generator(device="cpu").manual_seed(int(key))
        dev = torch.device(device)
        w = c.num_heads * c.head_dim

        def normal(shape, std):
            return (torch.randn(shape, generator=gen) * std).to(dev).requires_grad_(True)

        def zeros(n):
            return torch.zeros(n, device=dev, requires_grad=True)

        def ones(n):
            return torch.ones(n, device=dev, requires_grad=True)

        params = {"token_embed": normal((c.vocab_size, c.d_model), 0.02), "pos_embed": normal((c.max_seq_len, c.d_model), 0.02)}
        stats = {}
        for i, moe in enumerate(self.moe):
            v = moe.init(int(torch.randint(0, 2 ** 31 - 1, (), generator=gen)), torch.empty((1, c.d_model), device=dev))
            mp = v["params"]
            for grp in (mp["router"]["router_linear"], mp["experts"]["up_proj"], mp["experts"]["down_proj"]):
                for t in grp.values():
                    t.requires_grad_(True)
            params[f"layer_{i}"] = {
                "attn_norm": {"scale": ones(c.d_model), "bias": zeros(c.d_model)},
                "attn": {"qkv": normal((c.d_model, 3 * w), 1 / math.sqrt(c.d_model)), "out": normal((w, c.d_model), 1 / math.sqrt(w))},
                "ffn_norm": {"scale": ones(c.d_model), "bias": zeros(c.d_model)},
                "moe": mp,
            }
            stats[f"layer_{i}"] = {"moe": v["stats_buffer"]}
        params["final_norm"] = {"scale": ones(c.d_model), "bias": zeros(c.d_model)}
        params["lm_head"] = {"kernel": normal((c.d_model, c.vocab_size), 0.02), "bias": zeros(c.vocab_size)}
        return {"params": params, "stats_buffer": stats}

    def apply(self, variables, input_ids, mutable=False):
        """Returns (logits (B,S,V) bf16, total_aux_loss) -- model.py:212-294 -- and the mutable collections if asked."""
        c, p = self.config, variables["params"]
        mut = set(mutable) if not isinstance(mutable, bool) else (set(variables) - {"params"} if mutable else set())
        B, S = input_ids.shape
        bf = torch.bfloat16
        x = (F.embedding(input_ids.long(), p["token_embed"]) + p["pos_embed"][:S]).to(bf)
        aux_total = torch.zeros((), device=x.device)
        for i, moe in enumerate(self.moe):
            lp = p[f"layer_{i}"]
            h = F.layer_norm(x, (c.d_model,), lp["attn_norm"]["scale"].to(bf), lp["attn_norm"]["bias"].to(bf), 1e-6)
            qkv = h @ lp["attn"]["qkv"].to(bf)
            q, k, v = qkv.view(B, S, 3, c.num_heads, c.head_dim).permute(2, 0, 3, 1, 4)
            a = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, S, -1)
            x = x + a @ lp["attn"]["out"].to(bf)
            h = F.layer_norm(x, (c.d_model,), lp["ffn_norm"]["scale"].to(bf), lp["ffn_norm"]["bias"].to(bf), 1e-6)
            mv = {"params": lp["moe"], "stats_buffer": variables["stats_buffer"][f"layer_{i}"]["moe"]}
            if "stats_buffer" in mut:
                (y, aux), _ = moe.apply(mv, h, mutable=["stats_buffer"])
            else:
                y, aux = moe.apply(mv, h)
            x = x + y
            aux_total = aux_total + aux
        x = F.layer_norm(x, (c.d_model,), p["final_norm"]["scale"].to(bf), p["final_norm"]["bias"].to(bf), 1e-6)
        logits = x @ p["lm_head"]["kernel"].to(bf) + p["lm_head"]["bias"].to(bf)
        if isinstance(mutable, bool) and not mutable:
            return logits, aux_total
        return (logits, aux_total), {k_: variables[k_] for k_ in mut if k_ in variables}


def lm_loss(logits, input_ids, aux_loss):
    """Next-token cross entropy + aux (src/training/train.py:141-153)."""
    lm = F.cross_entropy(logits[:, :-1].float().reshape(-1, logits.shape[-1]), input_ids[:, 1:].long().reshape(-1))
    return lm + aux_loss, lm


def param_leaves(tree):
    for k in tree:
        v = tree[k]
        if isinstance(v, dict):
            yield from param_leaves(v)
        else:
            yield v
