"""StatefulLayer -- host-side mirror of the reference's Flax module contract for the FFN slot.

Reference interface this keeps:
  * Flax-linen style ``module.init(key, x) -> variables`` and
    ``module.apply(variables, x, mutable=[...]) -> (out, updated_collections)``
    (src/forde/moe.py:225-281; call site src/training/train.py:135-139; tests/verify_init.py:55-57).
  * variable collections, names verbatim:
      params/{up_proj,down_proj}/{kernel,bias}      Flax Dense layout, kernel is (in, out)
      neuron_assignments/assignments                (H,) int32, read in the fast loop, written by the slow loop
                                                    (README.md:15,20,50); initialised to 0 = generalist (D7)
      stats_buffer/{neuron_stats (H,5) f32, moments (H,5) f32, moments_lo (H,5) f32, sample_count i32,
                    step_count i32}
                                                    (pattern of moe.py:265-279; D6)
  * the per-step statistics leave through the mutable collection, not the return value (moe.py:281).

The math (README.md:47-56): z = x W_up + b_up ; F = relu / tanh / binary_step by assignment ;
h = F(z) + gate * z (gate 0.1 for specialists, 1.0 otherwise) ; y = h W_down + b_down ; backward uses the
straight-through estimator for binary_step.  All of it runs in the sm_100a kernels of csrc/ through the C ABI;
torch only owns memory, the stream and the autograd tape.

Difference from Flax that a user must know: JAX is functional, torch autograd is not.  The gradient half of
the statistics only exists once backward has run, so the ``stats_buffer`` leaves returned by ``apply`` are
updated IN PLACE when the backward pass of that call finishes (forward-only calls update nothing unless
``sense_forward_only=True``).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, Iterable, Optional, Tuple, Union

import torch

from . import ops

Collections = Dict[str, Dict[str, torch.Tensor]]

STAT_NAMES = ("grad_gini", "grad_gdp", "act_gini", "act_gdp", "act_variance")  # sensing.py:65-67


class _Workspace:
    """Per-(M, H, device) scratch that is reused every step: the neuron-major residual buffers and partials."""

    def __init__(self, M: int, H: int, device, sign_bits: bool = False):
        ld = ops.sample_pitch(M)
        P = ops.num_partials(M)
        self.M, self.H, self.ld, self.P = M, H, ld, P
        self.zT = torch.empty((H, ld), dtype=torch.bfloat16, device=device)
        self.hT = torch.empty((H, ld), dtype=torch.bfloat16, device=device)
        self.act_partials = torch.empty((P, 3, H), dtype=torch.float32, device=device)
        self.grad_partials = torch.empty((P, 3, H), dtype=torch.float32, device=device)
        # (z > 0) bit per neuron and sample for the neuron groups whose z box the forward does not store
        self.signs = torch.empty(ops.sign_shape(M, H), dtype=torch.int32, device=device) if sign_bits else None
        self.in_flight = False  # forward done, backward pending: zT / hT must not be overwritten
        self.generation = 0     # bumped by every forward that takes this workspace (stale-backward detection)


class _WorkspaceLease:
    """Lives on the autograd ctx of one forward.  When that ctx dies without its backward having run (a grad-enabled
    forward used for evaluation, a skipped step, an exception) the workspace is handed back instead of staying
    in flight forever."""

    __slots__ = ("ws", "generation")

    def __init__(self, ws: "_Workspace"):
        self.ws, self.generation = ws, ws.generation

    def __del__(self):
        ws = self.ws
        if ws.in_flight and ws.generation == self.generation:
            ws.in_flight = False


_DZ_POOL: Dict[Tuple[int, int, str, str], torch.Tensor] = {}


def _dz_scratch(H: int, ld: int, device, group: str = "") -> torch.Tensor:
    """dz^T is consumed inside one layer's backward, so every layer of the same shape AND the same group (the tower:
    towers may run concurrently on different streams) shares one buffer."""
    key = (H, ld, str(device), group)
    t = _DZ_POOL.get(key)
    if t is None:
        t = torch.empty((H, ld), dtype=torch.bfloat16, device=device)
        _DZ_POOL[key] = t
    return t


def grad_sink(p: torch.Tensor):
    """A train state may attach to a master parameter the fp32 buffer its gradient must end up in (a view of the flat
    gradient buffer, ``p._forde_grad``) and a callback to run once it is there (``p._forde_ready``, the bucketed
    all-reduce trigger).  Kernels that ACCUMULATE (red.add dW GEMMs, column sums, the finalisation's db_up) then
    write straight into it and the autograd function returns None for that input: no temporary, no zero fill, no
    AccumulateGrad add."""
    return getattr(p, "_forde_grad", None)


def grad_delivered(p: torch.Tensor) -> None:
    cb = getattr(p, "_forde_ready", None)
    if cb is not None:
        cb(p)


class _StatefulLayerFn(torch.autograd.Function):
    """fwd: two GEMM launches (up + fused epilogue, down).  bwd: dz (fused), dx, dW_up, dW_down, db, sensing.
    With LayerNorm parameters (ln_g, ln_b) the function is the second half of a FORDE block, LayerNorm(x + layer(x))
    (README.md:43): the add/norm kernels run inside, and in the backward the skip path's gradient is added in the dx
    GEMM's epilogue instead of by autograd."""

    @staticmethod
    def forward(ctx, x2d, w_up, b_up, w_down, b_down, ln_g, ln_b, layer, stats, w_up16, w_down16):
        M, d = x2d.shape
        H = w_up.shape[1]
        ws = layer._workspace(M, H, x2d.device)
        w_up16 = layer._bf16(w_up, "up") if w_up16 is None else w_up16
        w_down16 = layer._bf16(w_down, "down") if w_down16 is None else w_down16
        b_up32 = b_up.float() if b_up.dtype != torch.float32 else b_up
        b_down32 = b_down.float() if b_down.dtype != torch.float32 else b_down
        assign_i8 = layer._assign_i8
        y = torch.empty((M, w_down.shape[1]), dtype=torch.bfloat16, device=x2d.device)
        ops.sl_up_fwd(x2d, w_up16, b_up32, assign_i8, ws.zT, ws.hT, ws.act_partials, save_z=layer.save_z,
                      sign_bits=ws.signs)
        ctx.keep_bits = None
        if layer.top_k is not None:  # D8: keep the k largest |h| per token, zero the rest (in place)
            layer.last_topk_indices, layer.last_row_hoyer, ctx.keep_bits = ops.sl_topk_sparsify(ws.hT, M, layer.top_k)
        ctx.with_norm = ln_g is not None
        if ctx.with_norm:
            # y = x + layer(x): the skip connection is added in the down projection's epilogue, so the add/norm
            # (forward and backward) streams one tensor instead of two
            ops.sl_down_fwd(ws.hT, w_down16, b_down32, y, M, residual=x2d)
            out, mean, rstd = ops.add_layernorm_fwd(y, None, ln_g, ln_b, 1e-6)
            ctx.save_for_backward(x2d, w_up16, w_down16, y, ln_g, mean, rstd)
            ctx.ln_masters = (ln_g, ln_b)
        else:
            ops.sl_down_fwd(ws.hT, w_down16, b_down32, y, M)
            out = y
            ctx.save_for_backward(x2d, w_up16, w_down16)
        ctx.layer, ctx.ws, ctx.stats, ctx.assign_i8 = layer, ws, stats, assign_i8
        ctx.param_dtypes = (w_up.dtype, b_up.dtype, w_down.dtype, b_down.dtype)
        ctx.masters = (w_up, b_up, w_down, b_down)
        ctx.needs_dx = ctx.needs_input_grad[0]
        will_backward = any(ctx.needs_input_grad)
        ws.generation += 1
        ws.in_flight = will_backward
        ctx.lease = _WorkspaceLease(ws) if will_backward else None
        if stats is not None and not will_backward and layer.sense_forward_only:
            layer._finalize(ws, stats, with_grad=False)
        return out

    @staticmethod
    def backward(ctx, dy):
        layer, ws, stats, assign_i8 = ctx.layer, ctx.ws, ctx.stats, ctx.assign_i8
        if ctx.lease is None or ctx.lease.generation != ws.generation:
            raise RuntimeError(
                "StatefulLayer backward: the saved z^T / h^T of this forward were overwritten by a later forward of the "
                "same layer (a second backward through a retained graph after the workspace was released); run the "
                "backward before the next forward, or use separate layer objects")
        dy = dy.contiguous()
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        d_ln = (None, None)
        skip = None
        if ctx.with_norm:
            x2d, w_up16, w_down16, y, ln_g, mean, rstd = ctx.saved_tensors
            sg, sb = grad_sink(ctx.ln_masters[0]), grad_sink(ctx.ln_masters[1])
            direct_ln = sg is not None and sb is not None
            # d(x + y): the gradient of the layer's output AND of the skip path
            dy, dg, db = ops.add_layernorm_bwd(dy, y, None, ln_g, mean, rstd, d_gamma=sg if direct_ln else None,
                                               d_beta=sb if direct_ln else None)
            skip = dy
            if direct_ln:
                grad_delivered(ctx.ln_masters[0]); grad_delivered(ctx.ln_masters[1])
            else:
                d_ln = (dg, db)
        else:
            x2d, w_up16, w_down16 = ctx.saved_tensors
        M, d = x2d.shape
        H = ws.H
        dev = x2d.device
        dzT = _dz_scratch(H, ws.ld, dev, layer.name.split('/')[0])
        # top-k path (keep_bits): dL/dz is zero where h was dropped, applied in the same epilogue
        ops.sl_dz_bwd(dy, w_down16, ws.zT, ws.hT, assign_i8, dzT, ws.grad_partials, sign_bits=ws.signs,
                      keep_bits=ctx.keep_bits)
        d_out = dy.shape[1]
        sinks = [grad_sink(m) for m in ctx.masters]
        direct = all(s is not None for s in sinks)      # accumulate straight into the train state's gradient views
        if direct:
            dw_up, db_up, dw_down, db_down = sinks
        else:
            dw_up = torch.zeros((d, H), dtype=torch.float32, device=dev)
            dw_down = torch.zeros((H, d_out), dtype=torch.float32, device=dev)
            db_up = torch.zeros(H, dtype=torch.float32, device=dev)
            db_down = torch.zeros(d_out, dtype=torch.float32, device=dev)
        ops.sl_dw_up(x2d, dzT, dw_up)
        ops.sl_dw_down(ws.hT, dy, dw_down)
        ops.colsum_bf16(dy, db_down)
        if stats is not None:
            layer._finalize(ws, stats, with_grad=True, db_up=db_up)
        else:
            ops.stats_finalize(None, ws.grad_partials, ws.P, H, M, db_up=db_up)
        # dx last: with the fused add/norm it is accumulated over the skip gradient IN PLACE, i.e. over `dy`, which the
        # dW_down GEMM and the db_down column sum above still had to read
        dx = None
        if ctx.needs_dx:
            if skip is not None:
                dx = skip
                ops.sl_dx_bwd(dzT, w_up16, dx, M, residual=skip)
            else:
                dx = torch.empty((M, d), dtype=torch.bfloat16, device=dev)
                ops.sl_dx_bwd(dzT, w_up16, dx, M)
        ws.in_flight = False
        if direct:
            for m in ctx.masters:
                grad_delivered(m)
            return dx, None, None, None, None, d_ln[0], d_ln[1], None, None, None, None
        return dx, dw_up, db_up, dw_down, db_down, d_ln[0], d_ln[1], None, None, None, None


@dataclass
class StatefulLayer:
    """Drop-in for the reference's FFN-slot module (Flax dataclass-field construction, moe.py:225-243)."""

    hidden_dim: int
    out_dim: Optional[int] = None          # defaults to the input feature size (D2: d -> H -> d)
    top_k: Optional[int] = None            # D8 activation sparsification; None = off (vision/text towers)
    param_dtype: torch.dtype = torch.float32
    sense_forward_only: bool = False       # also accumulate activation-only stats in no-grad forwards
    save_z: bool = False                   # keep all of z^T (debugging); by default only what the backward reads
    # record (z > 0) bits where z is not stored and let the backward read those instead of h^T (1/16 of the bytes);
    # same results bit for bit, dz_bwd 191 -> 163 us at C2 vision dims.  On unless FORDE_SIGN_BITS=0 (or set
    # explicitly); ignored with save_z.
    sign_bits: bool = field(default_factory=lambda: os.environ.get("FORDE_SIGN_BITS", "1") == "1")
    name: str = "stateful_layer"
    max_workspaces: int = 4                # forwards of one shape that may await their backward at the same time

    def __post_init__(self):
        self._ws: Dict[Tuple[int, int, str], list] = {}
        self._bf16_cache: Dict[str, Tuple[int, int, torch.Tensor]] = {}
        self._assign_i8: Optional[torch.Tensor] = None
        self._assign_src: Optional[Tuple[int, int]] = None
        self.last_topk_indices: Optional[torch.Tensor] = None   # (M, k) i32, lax.top_k order (top_k mode)
        self.last_row_hoyer: Optional[torch.Tensor] = None      # (M,) f32 Hoyer sparsity of each token's dense h

    # ------------------------------------------------------------------ Flax-style API
    def init(self, key: Union[int, torch.Generator], x: torch.Tensor) -> Collections:
        """Create the variable collections for inputs shaped like ``x`` (..., d)."""
        d = x.shape[-1]
        H = self.hidden_dim
        out = self.out_dim or d
        dev = x.device
        gen = key if isinstance(key, torch.Generator) else torch.Generator(device="cpu").manual_seed(int(key))

        def lecun(fan_in, shape):  # flax nn.Dense default kernel_init (lecun_normal), zero bias
            w = torch.empty(shape, dtype=torch.float32)
            torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
            return (w * (math.sqrt(1.0 / fan_in) / 0.87962566103423978)).to(self.param_dtype).to(dev)

        return {
            "params": {
                "up_proj": {"kernel": lecun(d, (d, H)), "bias": torch.zeros(H, dtype=self.param_dtype, device=dev)},
                "down_proj": {"kernel": lecun(H, (H, out)), "bias": torch.zeros(out, dtype=self.param_dtype, device=dev)},
            },
            "neuron_assignments": {"assignments": torch.zeros(H, dtype=torch.int32, device=dev)},
            "stats_buffer": new_stats_buffer(H, dev),
        }

    def apply(self, variables: Collections, x: torch.Tensor, mutable: Union[bool, Iterable[str]] = False,
              post_norm: Optional[Dict[str, torch.Tensor]] = None):
        """Forward pass.  With ``mutable=["stats_buffer"]`` returns ``(y, {"stats_buffer": ...})`` and the
        sensing statistics of this step are accumulated into those leaves (see module docstring).
        ``post_norm={"scale": ..., "bias": ...}`` (fp32) makes the call the whole second half of a FORDE block,
        ``LayerNorm(x + layer(x))`` (README.md:43), with the add/norm fused around the layer."""
        p = variables["params"]
        mut = set(mutable) if not isinstance(mutable, bool) else (set(variables) - {"params"} if mutable else set())
        stats = variables.get("stats_buffer") if "stats_buffer" in mut else None
        self._load_assignments(variables["neuron_assignments"]["assignments"])
        lead = x.shape[:-1]
        x2d = x.reshape(-1, x.shape[-1])
        if x2d.dtype != torch.bfloat16:
            x2d = x2d.to(torch.bfloat16)
        x2d = x2d.contiguous()
        if x2d.shape[0] == 0:  # empty batch: nothing to launch, nothing to sense
            y = x2d.new_zeros((*lead, p["down_proj"]["kernel"].shape[1]))
            return y if (isinstance(mutable, bool) and not mutable) else (y, {k: variables[k] for k in mut if k in variables})
        if x2d.shape[1] % 8 != 0 or self.hidden_dim % 8 != 0:
            raise ValueError("StatefulLayer needs feature sizes that are multiples of 8 (16-byte TMA rows): "
                             f"got d={x2d.shape[1]}, hidden_dim={self.hidden_dim}")
        # optional pre-cast bf16 mirror of the fp32 master weights (kept by the train state, refreshed once per
        # optimizer step) so the forward does not re-cast them
        p16 = variables.get("params_bf16")
        w_up16 = p16["up_proj"]["kernel"] if p16 is not None else None
        w_down16 = p16["down_proj"]["kernel"] if p16 is not None else None
        ln_g = post_norm["scale"] if post_norm is not None else None
        ln_b = post_norm["bias"] if post_norm is not None else None
        if post_norm is not None and p["down_proj"]["kernel"].shape[1] != x2d.shape[1]:
            raise ValueError("post_norm needs out_dim == input width (residual connection)")
        y = _StatefulLayerFn.apply(x2d, p["up_proj"]["kernel"], p["up_proj"]["bias"], p["down_proj"]["kernel"],
                                   p["down_proj"]["bias"], ln_g, ln_b, self, stats, w_up16, w_down16)
        y = y.reshape(*lead, y.shape[-1])
        if isinstance(mutable, bool) and not mutable:
            return y
        return y, {k: variables[k] for k in mut if k in variables}

    __call__ = apply

    # ------------------------------------------------------------------ internals
    def _workspace(self, M, H, device) -> _Workspace:
        pool = self._ws.setdefault((M, H, str(device)), [])
        for ws in pool:
            if not ws.in_flight:
                return ws
        if len(pool) >= self.max_workspaces:
            raise RuntimeError(
                f"StatefulLayer {self.name!r}: {len(pool)} forwards of shape M={M}, H={H} are waiting for their backward "
                f"(each holds z^T and h^T, {2 * 2 * H * ops.sample_pitch(M) / 2 ** 20:.0f} MiB).  Run inference under "
                "torch.no_grad(), or raise max_workspaces if the module really is applied that often per step")
        ws = _Workspace(M, H, device, sign_bits=self.sign_bits and not self.save_z)  # the module is applied again before its previous backward ran
        pool.append(ws)
        return ws

    def _bf16(self, w: torch.Tensor, tag: str) -> torch.Tensor:
        if w.dtype == torch.bfloat16 and w.is_contiguous():
            return w.detach()
        c = self._bf16_cache.get(tag)
        if c is not None and c[0] == w.data_ptr() and c[1] == w._version:
            return c[2]
        w16 = w.detach().to(torch.bfloat16).contiguous()
        self._bf16_cache[tag] = (w.data_ptr(), w._version, w16)
        return w16

    def _load_assignments(self, a: torch.Tensor) -> None:
        """int8 mirror of the (H,) int32 brain map for the kernels.  Refreshed when the collection leaf changed
        (the slow loop writes it in place); while a CUDA graph is being captured the conversion is always
        recorded, so replays pick up whatever the slow loop wrote since."""
        src = (a.data_ptr(), a._version)
        if self._assign_i8 is None or self._assign_i8.numel() != a.numel() or self._assign_i8.device != a.device:
            self._assign_i8 = torch.empty(a.numel(), dtype=torch.int8, device=a.device)
            self._assign_src = None
        if self._assign_src != src or (a.is_cuda and torch.cuda.is_current_stream_capturing()):
            # D5: values outside {0,1} act as specialists; clamp into int8 range before the narrowing cast
            self._assign_i8.copy_(a.clamp(-128, 127))
            self._assign_src = src

    def _finalize(self, ws: _Workspace, stats: Dict[str, torch.Tensor], with_grad: bool, db_up=None) -> None:
        ops.stats_finalize(ws.act_partials, ws.grad_partials if with_grad else None, ws.P, ws.H, ws.M,
                           step_stats=stats.get("last_step_stats"), neuron_stats=stats["neuron_stats"],
                           moments=stats["moments"], moments_lo=stats.get("moments_lo"),
                           sample_count=stats["sample_count"], step_count=stats["step_count"], db_up=db_up)


def new_stats_buffer(H: int, device) -> Dict[str, torch.Tensor]:
    """Zero-initialised ``stats_buffer`` leaves for one layer (every leaf is a plain tensor so the slow loop's
    ``zeros_like`` reset, moe_slow_loop.py:441-449, works unchanged)."""
    return {
        "neuron_stats": torch.zeros((H, 5), dtype=torch.float32, device=device),
        "moments": torch.zeros((H, 5), dtype=torch.float32, device=device),
        # rounding residue of `moments` (double-single accumulation over a slow-loop interval, forde_stats_finalize)
        "moments_lo": torch.zeros((H, 5), dtype=torch.float32, device=device),
        "sample_count": torch.zeros((), dtype=torch.int32, device=device),
        "step_count": torch.zeros((), dtype=torch.int32, device=device),
    }
