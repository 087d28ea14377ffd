"""FORDE CLIP-style dual encoder: the CALLER of the hot path (README.md:26-43).

This is the benchmark / integration harness around the StatefulLayer, not the product: the towers are the
plain macro-architecture the README describes (patchify or token embedding -> L x FORDETransformerBlock ->
[CLS] -> Dense projection).  Everything a train step launches runs on this repository's sm_100a kernels: the FFN slot
of every block -- ``StatefulLayer`` --, the `add/norm` either side of it, the q/k/v/out projections, patch embedding and
heads (tcgen05 GEMM engine), the front ends (csrc/forde_frontend.cu), the attention core for head_dim 64 and up to 256
tokens (csrc/forde_attn.cu: T <= 128, the text tower; csrc/forde_attn2.cu: T <= 256, the ViT-B/16 vision tower) and the
contrastive loss.  Longer sequences (ViT-L/14: 257 tokens) fall back to stock torch SDPA (cuDNN flash, "library code"
for this task).  FORDE block = self-attention -> add/norm -> StatefulLayer -> add/norm (README.md:42-43).

Variables follow the Flax collection layout the reference's training loop threads through ``train_step``
(src/training/train.py:122-171): ``params`` (differentiated), ``neuron_assignments`` and ``stats_buffer``
(mutable, keyed ``{tower}/layer_{i}/stateful_layer`` so the slow loop's walker finds them,
src/forde/moe_slow_loop.py:145-170).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ops
from .clip_loss import INIT_LOG_TEMP, clip_contrastive_loss
from .stateful_layer import StatefulLayer, grad_delivered, grad_sink, new_stats_buffer


@dataclass
class TowerConfig:
    width: int
    layers: int
    heads: int
    hidden: int                      # StatefulLayer H (4 x width, src/training/train.py:250)


@dataclass
class DualEncoderConfig:
    """SURVEY.md section 8 config table (C1, C2/C3, C5); CLIP-standard dims where the README is silent."""
    image_size: int = 224
    patch_size: int = 16
    vision: TowerConfig = field(default_factory=lambda: TowerConfig(768, 12, 12, 3072))
    text: TowerConfig = field(default_factory=lambda: TowerConfig(512, 12, 8, 2048))
    vocab_size: int = 30522          # bert-base-uncased (src/data/download_to_drive.py:122)
    text_len: int = 77
    embed_dim: int = 512
    name: str = "vit_b16"

    @property
    def vision_tokens(self) -> int:
        return (self.image_size // self.patch_size) ** 2 + 1


def config_tiny() -> DualEncoderConfig:      # C1
    return DualEncoderConfig(32, 16, TowerConfig(128, 2, 4, 512), TowerConfig(128, 2, 4, 512), 30522, 16, 128, "tiny")


def config_vit_b16() -> DualEncoderConfig:   # C2 / C3
    return DualEncoderConfig()


def config_vit_l14() -> DualEncoderConfig:   # C5
    return DualEncoderConfig(224, 14, TowerConfig(1024, 24, 16, 4096), TowerConfig(768, 12, 12, 3072), 30522, 77, 768,
                             "vit_l14")


CONFIGS = {"tiny": config_tiny, "vit_b16": config_vit_b16, "vit_l14": config_vit_l14}


class _Bf16Mirror(torch.autograd.Function):
    """Use the pre-cast bf16 mirror of an fp32 master parameter; route the gradient back to the master."""

    @staticmethod
    def forward(ctx, master, mirror):
        return mirror.view_as(mirror)

    @staticmethod
    def backward(ctx, g):
        return g.float(), None


def _colsum(dy2d):
    out = torch.zeros(dy2d.shape[1], dtype=torch.float32, device=dy2d.device)
    ops.colsum_bf16(dy2d, out)
    return out


def _deliver_dense_grads(w, b, x2t, dy2):
    """dW = x^T dy and db = colsum(dy) for a Dense whose masters are (w, b).  With gradient sinks attached (train
    state) they are accumulated in place and (None, None) is returned to autograd; otherwise fresh tensors."""
    sw, sb = grad_sink(w), grad_sink(b)
    x2 = x2t.t()   # (M, in) row-major, the layout the weight-gradient GEMM reads in place
    if sw is not None and sb is not None:
        ops.gemm_f32(x2, dy2, True, True, out=sw, accumulate=True)     # dW += x^T dy, split-K red.add into the sink
        ops.colsum_bf16(dy2, sb)
        grad_delivered(w); grad_delivered(b)
        return None, None
    dw = torch.zeros(w.shape, dtype=torch.float32, device=dy2.device)
    ops.gemm_f32(x2, dy2, True, True, out=dw, accumulate=True)
    return dw, _colsum(dy2)


class _Dense(torch.autograd.Function):
    """y = x W (+ b) for bf16 activations.  Forward, input gradient and weight gradient run on this repository's tcgen05
    GEMM engine (bias fused in the epilogue; dW accumulated in fp32 straight into the gradient sink), the bias
    gradient on the column-sum kernel.  `w16` is the bf16 mirror (in, out); `b` may be None (projection heads)."""

    @staticmethod
    def forward(ctx, x, w, b, w16, b16, residual=None, skip_holder=None):
        """``residual`` (shape of the output): y = x W + b + residual, added in the GEMM epilogue; its gradient is the
        output gradient itself.  It is returned to autograd, or -- with ``skip_holder``, a dict shared with the
        _QKVDense of the same block -- left in the holder so that _QKVDense.backward accumulates its GEMMs on top of it
        in their epilogues instead of autograd adding it with a separate pass."""
        x2 = x.reshape(-1, x.shape[-1])
        if x2.stride(1) != 1 or x2.stride(0) % 8 != 0 or x2.data_ptr() % 16 != 0:
            x2 = x2.contiguous()        # (a row-pitched view such as the padded patch rows is fine for TMA as it is)
        r2 = None
        if residual is not None:
            r2 = residual.reshape(-1, residual.shape[-1])
            r2 = r2 if r2.is_contiguous() else r2.contiguous()
        y = ops.gemm_bf16(x2, w16, False, True, bias_col=None if b is None else b.detach().float(), residual=r2)
        ctx.save_for_backward(x2, w16)
        ctx.lead = x.shape[:-1]
        ctx.masters = (w, b)
        ctx.needs_dx = ctx.needs_input_grad[0]
        ctx.has_residual = residual is not None
        ctx.skip_holder = skip_holder if residual is not None else None
        ctx.res_shape = residual.shape if residual is not None else None
        return y.reshape(*x.shape[:-1], w16.shape[1])

    @staticmethod
    def backward(ctx, dy):
        x2, w16 = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        if dy2.dtype != torch.bfloat16:
            dy2 = dy2.to(torch.bfloat16)
        dx = None
        if ctx.needs_dx:                                                             # dx = dy W^T
            dx = ops.gemm_bf16(dy2, w16, False, False).reshape(*ctx.lead, w16.shape[0])
        w, b = ctx.masters
        d_res = None
        if ctx.has_residual:
            if ctx.skip_holder is not None:
                ctx.skip_holder["ds"] = dy2.view(ctx.res_shape)      # picked up (and accumulated onto) by _QKVDense
            else:
                d_res = dy2.view(ctx.res_shape)
        if b is None:
            sw = grad_sink(w)
            dw = sw if sw is not None else torch.zeros(w.shape, dtype=torch.float32, device=dy2.device)
            ops.gemm_f32(x2, dy2, True, True, out=dw, accumulate=True)               # dW += x^T dy
            if sw is not None:
                grad_delivered(w)
                dw = None
            return dx, dw, None, None, None, d_res, None
        dw, db = _deliver_dense_grads(w, b, x2.t(), dy2)
        return dx, dw, db, None, None, d_res, None


class _SelfAttention(torch.autograd.Function):
    """softmax(q k^T / sqrt(d_head)) v for sequences of at most ops.attn_max_tokens() tokens and d_head = 64 on this
    repository's tcgen05 kernels: csrc/forde_attn.cu (T <= 128, the text tower: one tile per (batch, head)) and
    csrc/forde_attn2.cu (T <= 256, the ViT-B/16 vision tower: two 128-row blocks per item, operands ring-buffered).
    q, k, v: (B, T, w) bf16, heads packed along w."""

    @staticmethod
    def forward(ctx, q, k, v, heads):
        B, T, w = q.shape
        q2, k2, v2 = (t.reshape(B * T, w) for t in (q, k, v))
        q2, k2, v2 = (t if t.is_contiguous() else t.contiguous() for t in (q2, k2, v2))
        o, lse = ops.attn_fwd(q2, k2, v2, B, T, heads)
        ctx.save_for_backward(q2, k2, v2, lse, o)
        ctx.dims = (B, T, heads, w)
        return o.view(B, T, w)

    @staticmethod
    def backward(ctx, d_o):
        q2, k2, v2, lse, o = ctx.saved_tensors
        B, T, heads, w = ctx.dims
        d2 = d_o.reshape(B * T, w)
        if d2.dtype != torch.bfloat16:
            d2 = d2.to(torch.bfloat16)
        if not d2.is_contiguous():
            d2 = d2.contiguous()
        dq, dk, dv = ops.attn_bwd(q2, k2, v2, d2, lse, B, T, heads, o=o)
        return dq.view(B, T, w), dk.view(B, T, w), dv.view(B, T, w), None


class _VisionAssemble(torch.autograd.Function):
    """x[b] = [cls + pos[0] ; patch tokens + pos[1:]] (README.md:30-34) on the front-end kernels of csrc/forde_frontend.cu;
    d(cls), d(pos_embed) are accumulated in fp32 straight into their gradient sinks."""

    @staticmethod
    def forward(ctx, tok, cls, pos, cls16, pos16, B):
        ctx.masters = (cls, pos)
        return ops.vision_assemble_fwd(tok.contiguous(), cls16.contiguous(), pos16.contiguous(), B)

    @staticmethod
    def backward(ctx, dx):
        cls, pos = ctx.masters
        sc, sp = grad_sink(cls), grad_sink(pos)
        direct = sc is not None and sp is not None
        d_cls = sc if direct else torch.zeros(cls.shape, dtype=torch.float32, device=dx.device)
        d_pos = sp if direct else torch.zeros(pos.shape, dtype=torch.float32, device=dx.device)
        d_tok = ops.vision_assemble_bwd(dx.contiguous(), d_cls, d_pos)
        if direct:
            grad_delivered(cls); grad_delivered(pos)
            return d_tok, None, None, None, None, None
        return d_tok, d_cls, d_pos, None, None, None


class _TokenEmbed(torch.autograd.Function):
    """x[b][t] = token_embed[ids[b][t]] + pos_embed[t] and its scatter-add backward (fp32 red.add into the sinks)."""

    @staticmethod
    def forward(ctx, ids, table, pos, table16, pos16):
        ids = ids.to(torch.int32).contiguous()
        ctx.masters = (table, pos)
        ctx.save_for_backward(ids)
        return ops.embed_tokens_fwd(ids, table16.contiguous(), pos16.contiguous())

    @staticmethod
    def backward(ctx, dx):
        (ids,) = ctx.saved_tensors
        table, pos = ctx.masters
        st, sp = grad_sink(table), grad_sink(pos)
        direct = st is not None and sp is not None
        d_table = st if direct else torch.zeros(table.shape, dtype=torch.float32, device=dx.device)
        d_pos = sp if direct else torch.zeros(pos.shape, dtype=torch.float32, device=dx.device)
        ops.embed_tokens_bwd(ids, dx.contiguous(), d_table, d_pos)
        if direct:
            grad_delivered(table); grad_delivered(pos)
            return None, None, None, None, None
        return None, d_table, d_pos, None, None


class _QKVDense(torch.autograd.Function):
    """The three attention input projections sharing one input: three forward GEMMs; the backward accumulates dx
    inside the GEMM epilogues (EPI_STORE_ADD: no separate add passes over the (M, w) gradient)."""

    @staticmethod
    def forward(ctx, x, wq, bq, wk, bk, wv, bv, wq16, bq16, wk16, bk16, wv16, bv16, skip_holder=None):
        ctx.skip_holder = skip_holder
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        ctx.save_for_backward(x2, wq16, wk16, wv16)
        ctx.shape = x.shape
        ctx.masters = ((wq, bq), (wk, bk), (wv, bv))
        return tuple(ops.gemm_bf16(x2, w16, False, True, bias_col=b.detach().float()).view(x.shape)
                     for w16, b in ((wq16, bq), (wk16, bk), (wv16, bv)))

    @staticmethod
    def backward(ctx, dq, dk, dv):
        x2, wq16, wk16, wv16 = ctx.saved_tensors
        d = [t.reshape(-1, t.shape[-1]) for t in (dq, dk, dv)]
        d = [t if t.is_contiguous() else t.contiguous() for t in d]
        base = ctx.skip_holder.pop("ds", None) if ctx.skip_holder is not None else None
        if base is not None:
            # the block's skip path (x -> add/norm) left its gradient here: dx = d(skip) + dq Wq^T + dk Wk^T + dv Wv^T
            dx = base.reshape(-1, base.shape[-1])
            ops.gemm_bf16(d[0], wq16, False, False, add_to=dx)
        else:
            dx = ops.gemm_bf16(d[0], wq16, False, False)             # dx = dq Wq^T
            if not dx.is_contiguous():
                dx = dx.contiguous()
        ops.gemm_bf16(d[1], wk16, False, False, add_to=dx)           #    += dk Wk^T  (accumulated in the epilogue)
        ops.gemm_bf16(d[2], wv16, False, False, add_to=dx)           #    += dv Wv^T
        xt = x2.t()
        grads = []
        for t, (w, b) in zip(d, ctx.masters):
            grads += list(_deliver_dense_grads(w, b, xt, t))
        return (dx.view(ctx.shape), *grads, None, None, None, None, None, None, None)


class _AddLayerNorm(torch.autograd.Function):
    """LayerNorm(x + r) on the fused sm_100a kernel (csrc/forde_norm.cu); gamma / beta are fp32 masters.  With
    ``r = None`` the input is already the sum (formed in the producing GEMM's epilogue)."""

    @staticmethod
    def forward(ctx, x, r, gamma, beta, skip_holder=None):
        # skip_holder: a dict shared with the _QKVDense of the same block.  The gradient of the skip input x is then
        # not returned to autograd (which would add it to the projections' input gradient with a separate pass) but
        # left in the holder, and _QKVDense.backward accumulates its GEMMs on top of it.
        ctx.skip_holder = skip_holder
        x = x.contiguous()
        r = r.contiguous() if r is not None else None
        out, mean, rstd = ops.add_layernorm_fwd(x, r, gamma, beta, 1e-6)
        ctx.save_for_backward(x, r, gamma, mean, rstd)
        ctx.masters = (gamma, beta)
        ctx.has_r = r is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, r, gamma, mean, rstd = ctx.saved_tensors
        sg, sb = grad_sink(ctx.masters[0]), grad_sink(ctx.masters[1])
        direct = sg is not None and sb is not None
        ds, dg, db = ops.add_layernorm_bwd(dy.contiguous(), x, r, gamma, mean, rstd, d_gamma=sg if direct else None,
                                           d_beta=sb if direct else None)
        dx_skip = ds
        if ctx.skip_holder is not None:
            ctx.skip_holder["ds"] = ds
            dx_skip = None
        dr = ds if ctx.has_r else None
        if direct:
            grad_delivered(ctx.masters[0]); grad_delivered(ctx.masters[1])
            return dx_skip, dr, None, None, None
        return dx_skip, dr, dg, db, None


class _JoinTowerStreams(torch.autograd.Function):
    """Identity on the text embedding, applied on the caller's stream after the forward join.  Its backward is the
    first node of the text tower's backward pass: it queues an engine callback that runs when the whole backward has
    been enqueued and makes the stream ``backward()`` was called on wait for the text tower's stream (that tower's
    parameter gradients are accumulated in place by its kernels, so autograd itself has nothing to synchronise on)."""

    @staticmethod
    def forward(ctx, t, model, device):
        ctx.model, ctx.device = model, device
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        model, device = ctx.model, ctx.device
        torch.autograd.Variable._execution_engine.queue_callback(lambda: model.join_streams(device))
        return g, None, None


def _flatten(tree, prefix=""):
    out = {}
    for k in sorted(tree):
        v = tree[k]
        name = f"{prefix}/{k}" if prefix else k
        if isinstance(v, dict):
            out.update(_flatten(v, name))
        else:
            out[name] = v
    return out


class FORDEDualEncoder:
    def __init__(self, cfg: DualEncoderConfig, concurrent_towers: Optional[bool] = None):
        """``concurrent_towers`` (default on, env FORDE_TOWER_STREAMS=0 turns it off): the two towers are independent
        until the contrastive loss, so the text tower is enqueued on its own CUDA stream, forward and backward (autograd
        runs a node's backward on its forward's stream).  Its GEMMs have partial last waves (154 tiles on 74 CTA pairs at
        C2) and the vision tower's tiles fill those SMs; inside a captured step this is a fork/join of the graph."""
        self.cfg = cfg
        self.concurrent_towers = (os.environ.get("FORDE_TOWER_STREAMS", "1") != "0") if concurrent_towers is None \
            else bool(concurrent_towers)
        self._tower_streams: Dict[str, "torch.cuda.Stream"] = {}
        self.layers: Dict[str, StatefulLayer] = {}
        for tower, tc in (("vision", cfg.vision), ("text", cfg.text)):
            for i in range(tc.layers):
                self.layers[f"{tower}/layer_{i}"] = StatefulLayer(hidden_dim=tc.hidden, name=f"{tower}/layer_{i}")

    # ------------------------------------------------------------------ init
    def init(self, key: int, device) -> Dict[str, dict]:
        cfg = self.cfg
        gen = torch.Generator(device="cpu").manual_seed(int(key))
        dev = torch.device(device)

        def normal(shape, std):
            return (torch.randn(shape, generator=gen) * std).to(dev)

        def dense(i, o, bias=True):
            p = {"kernel": normal((i, o), 1.0 / math.sqrt(i))}
            if bias:
                p["bias"] = torch.zeros(o, device=dev)
            return p

        def ln(w):
            return {"scale": torch.ones(w, device=dev), "bias": torch.zeros(w, device=dev)}

        params, assigns, stats = {}, {}, {}
        for tower, tc in (("vision", cfg.vision), ("text", cfg.text)):
            tp, ta, ts = {}, {}, {}
            w = tc.width
            if tower == "vision":
                tp["patch_embed"] = dense(cfg.patch_size * cfg.patch_size * 3, w)
                tp["cls"] = normal((w,), 0.02)
                tp["pos_embed"] = normal((cfg.vision_tokens, w), 0.02)
            else:
                tp["token_embed"] = normal((cfg.vocab_size, w), 0.02)
                tp["pos_embed"] = normal((cfg.text_len, w), 0.01)
            for i in range(tc.layers):
                sl = self.layers[f"{tower}/layer_{i}"]
                v = sl.init(int(torch.randint(0, 2 ** 31 - 1, (), generator=gen)), torch.empty((1, w), device=dev))
                tp[f"layer_{i}"] = {"attn": {"query": dense(w, w), "key": dense(w, w), "value": dense(w, w),
                                             "out": dense(w, w)}, "ln_1": ln(w), "ln_2": ln(w),
                                    "stateful_layer": v["params"]}
                ta[f"layer_{i}"] = {"stateful_layer": v["neuron_assignments"]}
                ts[f"layer_{i}"] = {"stateful_layer": new_stats_buffer(tc.hidden, dev)}
            tp["proj"] = dense(w, cfg.embed_dim, bias=False)
            params[tower], assigns[tower], stats[tower] = tp, ta, ts
        params["log_temp"] = torch.tensor(INIT_LOG_TEMP, device=dev)
        return {"params": params, "neuron_assignments": assigns, "stats_buffer": stats}

    # ------------------------------------------------------------------ streams
    def tower_stream(self, device) -> "torch.cuda.Stream":
        """The text tower's stream on ``device`` (created on first use)."""
        key = str(torch.device(device))
        s = self._tower_streams.get(key)
        if s is None:
            s = self._tower_streams[key] = torch.cuda.Stream(device=device)
        return s

    def join_streams(self, device) -> None:
        """Make the current stream wait for the text tower's stream (call after ``loss.backward()``: the text tower's
        backward ran there, and its gradients were accumulated in place, not handed to autograd)."""
        s = self._tower_streams.get(str(torch.device(device)))
        if s is not None:
            torch.cuda.current_stream(device).wait_stream(s)

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _w(p, p16, name):
        """bf16 compute weight of a master parameter (uses the pre-cast mirror when the train state keeps one)."""
        if p16 is not None:
            return _Bf16Mirror.apply(p[name], p16[name])
        return p[name].to(torch.bfloat16)

    def _block(self, key, x, p, p16, heads, variables, sense):
        B, T, w = x.shape
        pa, pa16 = p["attn"], (p16["attn"] if p16 is not None else None)

        def mirror(name, leaf):
            return pa16[name][leaf] if pa16 is not None else pa[name][leaf].to(torch.bfloat16)

        # Flax MultiHeadDotProductAttention layout: separate query / key / value / out projections.  Heads are a
        # strided view of the (B, T, w) projections, so nothing is re-packed on the way in or out of the kernel.
        holder = {} if x.requires_grad else None      # carries d(skip) from the add/norm to the q/k/v input gradient
        q, k, v = _QKVDense.apply(x, pa["query"]["kernel"], pa["query"]["bias"], pa["key"]["kernel"], pa["key"]["bias"],
                                  pa["value"]["kernel"], pa["value"]["bias"], mirror("query", "kernel"),
                                  mirror("query", "bias"), mirror("key", "kernel"), mirror("key", "bias"),
                                  mirror("value", "kernel"), mirror("value", "bias"), holder)
        if w // heads == 64 and q.is_cuda and T <= int(os.environ.get("FORDE_ATTN_MAX_TOKENS", ops.attn_max_tokens())):
            a = _SelfAttention.apply(q, k, v, heads)           # both towers at C2: this repository's tcgen05 kernels
        else:                                                 # longer sequences (ViT-L/14: 257 tokens): cuDNN flash
            q, k, v = (t.view(B, T, heads, w // heads).transpose(1, 2) for t in (q, k, v))
            a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, T, w)
        # x + attention(x): the skip connection is added in the out-projection's epilogue, the add/norm streams the sum
        s_ = _Dense.apply(a, pa["out"]["kernel"], pa["out"]["bias"], mirror("out", "kernel"), mirror("out", "bias"), x,
                          holder)
        x = _AddLayerNorm.apply(s_, None, p["ln_1"]["scale"], p["ln_1"]["bias"], None)
        tower, lname = key.split("/")
        sl_vars = {
            "params": p["stateful_layer"],
            "neuron_assignments": variables["neuron_assignments"][tower][lname]["stateful_layer"],
            "stats_buffer": variables["stats_buffer"][tower][lname]["stateful_layer"],
        }
        if p16 is not None:
            sl_vars["params_bf16"] = p16["stateful_layer"]
        layer = self.layers[key]
        # second half of the FORDE block in one call: LayerNorm(x + StatefulLayer(x))
        if sense:
            out, _ = layer.apply(sl_vars, x, mutable=["stats_buffer"], post_norm=p["ln_2"])
        else:
            out = layer.apply(sl_vars, x, post_norm=p["ln_2"])
        return out

    def encode_image(self, variables, images, sense=True, params_bf16=None):
        """images (B, S, S, 3) NHWC (src/data/download_to_drive.py:173), any float dtype."""
        cfg = self.cfg
        p = variables["params"]["vision"]
        p16 = params_bf16["vision"] if params_bf16 is not None else None
        B = images.shape[0]

        def m16(*path):      # bf16 compute mirror of a master leaf (the train state's pre-cast one when there is one)
            t, t16 = p, p16
            for k in path:
                t = t[k]
                t16 = t16[k] if t16 is not None else None
            return t16 if t16 is not None else t.detach().to(torch.bfloat16)

        # patchify (layout only) -> patch-embedding GEMM on the tcgen05 engine -> +CLS, +pos (README.md:30-34)
        patches = ops.patchify(images.to(torch.bfloat16).contiguous(), cfg.patch_size)
        tok = _Dense.apply(patches, p["patch_embed"]["kernel"], p["patch_embed"]["bias"], m16("patch_embed", "kernel"),
                           None)
        x = _VisionAssemble.apply(tok, p["cls"], p["pos_embed"], m16("cls"), m16("pos_embed"), B)
        for i in range(cfg.vision.layers):
            x = self._block(f"vision/layer_{i}", x, p[f"layer_{i}"], p16 and p16[f"layer_{i}"], cfg.vision.heads,
                            variables, sense)
        return _Dense.apply(x[:, 0], p["proj"]["kernel"], None, m16("proj", "kernel"), None)     # [CLS] -> embed_dim

    def encode_text(self, variables, input_ids, sense=True, params_bf16=None):
        """input_ids (B, S) int (src/data/download_to_drive.py:106-108); token 0 is [CLS]."""
        cfg = self.cfg
        p = variables["params"]["text"]
        p16 = params_bf16["text"] if params_bf16 is not None else None
        def m16(*path):
            t, t16 = p, p16
            for k in path:
                t = t[k]
                t16 = t16[k] if t16 is not None else None
            return t16 if t16 is not None else t.detach().to(torch.bfloat16)

        x = _TokenEmbed.apply(input_ids, p["token_embed"], p["pos_embed"], m16("token_embed"), m16("pos_embed"))
        for i in range(cfg.text.layers):
            x = self._block(f"text/layer_{i}", x, p[f"layer_{i}"], p16 and p16[f"layer_{i}"], cfg.text.heads,
                            variables, sense)
        return _Dense.apply(x[:, 0], p["proj"]["kernel"], None, m16("proj", "kernel"), None)     # token 0 = [CLS]

    def apply(self, variables, images, input_ids, mutable=False, params_bf16=None, group=None):
        """Returns the global-batch contrastive loss (and the mutable collections when asked, Flax style)."""
        mut = set(mutable) if not isinstance(mutable, bool) else (set(variables) - {"params"} if mutable else set())
        sense = "stats_buffer" in mut
        if self.concurrent_towers and images.is_cuda:
            cur = torch.cuda.current_stream(images.device)
            side = self.tower_stream(images.device)
            side.wait_stream(cur)                     # fork: the text tower sees everything enqueued so far
            with torch.cuda.stream(side):
                txt = self.encode_text(variables, input_ids, sense, params_bf16)
            img = self.encode_image(variables, images, sense, params_bf16)
            cur.wait_stream(side)                     # join before the loss reads both embeddings
            txt.record_stream(cur)
            if txt.requires_grad:
                txt = _JoinTowerStreams.apply(txt, self, images.device)
        else:
            img = self.encode_image(variables, images, sense, params_bf16)
            txt = self.encode_text(variables, input_ids, sense, params_bf16)
        loss = clip_contrastive_loss(img, txt, variables["params"]["log_temp"], group)
        if isinstance(mutable, bool) and not mutable:
            return loss
        return loss, {k: variables[k] for k in mut if k in variables}


def stateful_layer_flops(cfg: DualEncoderConfig, batch: int) -> float:
    """Algorithmic FLOPs of the hot path per train step: 12 M d H per layer (six GEMMs), SURVEY.md 8d."""
    v = 12.0 * batch * cfg.vision_tokens * cfg.vision.width * cfg.vision.hidden * cfg.vision.layers
    t = 12.0 * batch * cfg.text_len * cfg.text.width * cfg.text.hidden * cfg.text.layers
    return v + t
