"""CPU restatement of the FORDE dual-encoder train step (fast loop).  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

The reference's own fast loop cannot run here (JAX/Flax/optax are not installable and the dual encoder survives
only as README prose), so this is the "port" baseline: fp32 PyTorch-CPU math that follows
  README.md:26-56      two towers of FORDETransformerBlocks, [CLS] -> Dense projection, StatefulLayer, STE, gates
  README.md:35         cosine-similarity contrastive loss (symmetric cross entropy)
  sensing.py:34-69     calculate_neuron_stats as a SEPARATE pass over z and dL/dz (what the reference does),
                       with dL/dz captured at the Dense output (tests/reproduce_grad_sink.py:16-17)
  train.py:74-77,159-165   value_and_grad, clip_by_global_norm(1.0) -> adamw
PARITY UNPINNED for everything but the sensing formulae (see oracle/forde_oracle.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


class _BinaryStepSTE(torch.autograd.Function):
    """README.md:54-55: forward (z > 0), backward identity."""

    @staticmethod
    def forward(ctx, z):
        return (z > 0).to(z.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def hoyer_sparsity(x: torch.Tensor) -> torch.Tensor:
    """sensing.py:5-32 along the last axis."""
    n = x.shape[-1]
    l1 = x.abs().sum(-1)
    l2 = x.pow(2).sum(-1).sqrt()
    safe = torch.where(l2 == 0, torch.ones_like(l2), l2)
    denom = 1.0 if n == 1 else math.sqrt(n) - 1.0
    s = (math.sqrt(n) - l1 / safe) / denom
    s = torch.where(l2 == 0, torch.zeros_like(s), s)
    return torch.zeros_like(s) if n == 1 else s


def calculate_neuron_stats(activations: torch.Tensor, gradients: torch.Tensor) -> torch.Tensor:
    """sensing.py:34-69 -> (F,5) [grad_gini, grad_gdp, act_gini, act_gdp, act_variance]."""
    f = activations.shape[-1]
    a = activations.reshape(-1, f).float()
    g = gradients.reshape(-1, f).float()
    return torch.stack([hoyer_sparsity(g.t()), g.abs().mean(0), hoyer_sparsity(a.t()), a.abs().mean(0),
                        a.var(0, unbiased=False)], dim=-1)


def stateful_layer(x, p, assign, sink):
    """README.md:47-56 (D1-D5).  `sink` collects z (with retain_grad) for the sensing pass."""
    z = x @ p["up_proj"]["kernel"] + p["up_proj"]["bias"]
    if sink is not None:
        z.retain_grad()
        sink.append(z)
    a = assign.to(torch.long)
    f = torch.where(a == 0, torch.relu(z), torch.where(a == 1, torch.tanh(z), _BinaryStepSTE.apply(z)))
    gate = torch.where((a == 0) | (a == 1), 1.0, 0.1).to(z.dtype)
    h = f + gate * z
    return h @ p["down_proj"]["kernel"] + p["down_proj"]["bias"]


def _block(x, p, heads, assign, sink):
    B, T, w = x.shape
    q, k, v = ((x @ p["attn"][n]["kernel"] + p["attn"][n]["bias"]).view(B, T, heads, w // heads).transpose(1, 2)
               for n in ("query", "key", "value"))
    a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, T, w)
    a = a @ p["attn"]["out"]["kernel"] + p["attn"]["out"]["bias"]
    x = F.layer_norm(x + a, (w,), p["ln_1"]["scale"], p["ln_1"]["bias"], 1e-6)
    m = stateful_layer(x, p["stateful_layer"], assign, sink)
    return F.layer_norm(x + m, (w,), p["ln_2"]["scale"], p["ln_2"]["bias"], 1e-6)


def forward(cfg, variables, images, input_ids, sinks=None):
    """Global-batch contrastive loss of the dual encoder; fills sinks[tower][i] with z when given."""
    P, A = variables["params"], variables["neuron_assignments"]
    pv = P["vision"]
    B = images.shape[0]
    g, ps = cfg.image_size // cfg.patch_size, cfg.patch_size
    x = images.float().view(B, g, ps, g, ps, 3).permute(0, 1, 3, 2, 4, 5).reshape(B, g * g, ps * ps * 3)
    x = x @ pv["patch_embed"]["kernel"] + pv["patch_embed"]["bias"]
    x = torch.cat([pv["cls"].expand(B, 1, -1), x], 1) + pv["pos_embed"]
    for i in range(cfg.vision.layers):
        x = _block(x, pv[f"layer_{i}"], cfg.vision.heads, A["vision"][f"layer_{i}"]["stateful_layer"]["assignments"],
                   None if sinks is None else sinks["vision"])
    img = x[:, 0] @ pv["proj"]["kernel"]
    pt = P["text"]
    x = F.embedding(input_ids.long(), pt["token_embed"]) + pt["pos_embed"][: input_ids.shape[1]]
    for i in range(cfg.text.layers):
        x = _block(x, pt[f"layer_{i}"], cfg.text.heads, A["text"][f"layer_{i}"]["stateful_layer"]["assignments"],
                   None if sinks is None else sinks["text"])
    txt = x[:, 0] @ pt["proj"]["kernel"]
    u = img / img.norm(dim=-1, keepdim=True)
    v = txt / txt.norm(dim=-1, keepdim=True)
    s = P["log_temp"].exp() * (u @ v.t())
    tgt = torch.arange(B)
    return 0.5 * (F.cross_entropy(s, tgt) + F.cross_entropy(s.t(), tgt))


def to_cpu_fp32(tree):
    if isinstance(tree, dict):
        return {k: to_cpu_fp32(v) for k, v in tree.items()}
    t = tree.detach().cpu()
    return t.float().clone() if t.is_floating_point() else t.clone()


def _param_leaves(tree):
    for k in sorted(tree):
        v = tree[k]
        if isinstance(v, dict):
            yield from _param_leaves(v)
        else:
            yield v


class CpuTrainer:
    """fast-loop step: value_and_grad -> sensing pass -> clip(1.0) -> AdamW -> stats_buffer accumulation."""

    def __init__(self, cfg, variables, lr=1e-4, wd=0.01):
        self.cfg = cfg
        self.variables = to_cpu_fp32(variables)
        self.leaves = list(_param_leaves(self.variables["params"]))
        for p in self.leaves:
            p.requires_grad_(True)
        self.opt = torch.optim.AdamW(self.leaves, lr=lr, weight_decay=wd)

    def step(self, images, input_ids, sense=True):
        sinks = {"vision": [], "text": []} if sense else None
        self.opt.zero_grad(set_to_none=True)
        loss = forward(self.cfg, self.variables, images, input_ids, sinks)
        loss.backward()
        step_stats = {}
        if sense:
            for tower in ("vision", "text"):
                for i, z in enumerate(sinks[tower]):
                    st = calculate_neuron_stats(z.detach(), z.grad)       # the separate pass over z and dL/dz
                    sb = self.variables["stats_buffer"][tower][f"layer_{i}"]["stateful_layer"]
                    sb["neuron_stats"] += st
                    sb["step_count"] += 1
                    sb["sample_count"] += z.shape[0] * z.shape[1]
                    step_stats[f"{tower}/layer_{i}"] = st
        gnorm = torch.nn.utils.clip_grad_norm_(self.leaves, 1.0)
        self.opt.step()
        return float(loss.detach()), float(gnorm), step_stats
