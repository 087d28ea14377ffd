"""CLIP-style contrastive loss over the global batch (reference spec: README.md:35).

loss = 1/2 (CE(S, arange) + CE(S^T, arange)),  S = exp(tau) * normalise(img) . normalise(txt)^T.

Data-parallel formulation (one process per GPU, batch sharded):
  * all-gather of the text embeddings, each rank computes its own ROWS of S (its images x all texts):
    the row log-sum-exp is then complete locally;
  * the column log-sum-exp is combined across ranks from per-rank (max, sum-exp) pairs: one MAX and one SUM
    all-reduce over B_global floats;
  * backward: d(img) is complete locally; d(txt) needs the transpose of the all-gather, a reduce-scatter(SUM).
All math runs in csrc/ kernels (normalise, tcgen05 GEMMs for S / dU / dV, row+column log-softmax, dS);
torch.distributed (NCCL) only moves bytes.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.distributed as dist

from . import ops

INIT_LOG_TEMP = math.log(1.0 / 0.07)  # D9


class _ClipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, txt, log_temp, group):
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank(group) if world > 1 else 0
        img16 = img.detach().to(torch.bfloat16).contiguous()
        txt16 = txt.detach().to(torch.bfloat16).contiguous()
        Bl = img16.shape[0]
        if world > 1:
            txt_all = torch.empty((Bl * world, txt16.shape[1]), dtype=torch.bfloat16, device=txt16.device)
            dist.all_gather_into_tensor(txt_all, txt16, group=group)
        else:
            txt_all = txt16
        row_offset = rank * Bl
        lt = log_temp.detach().reshape(1).float().contiguous()
        f = ops.clip_fwd(img16, txt_all, lt, row_offset)
        col_max, col_sum = f["col_max"], f["col_sum"]
        if world > 1:
            gmax = col_max.clone()
            dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
            ops.clip_col_rescale(col_max, gmax, col_sum)
            dist.all_reduce(col_sum, op=dist.ReduceOp.SUM, group=group)
            col_max = gmax
        loss, col_lse = ops.clip_loss(f, col_max, col_sum, row_offset)
        if world > 1:
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
        ctx.f, ctx.col_lse, ctx.row_offset, ctx.world, ctx.group = f, col_lse, row_offset, world, group
        ctx.Bl = Bl
        ctx.in_dtypes = (img.dtype, txt.dtype, log_temp.dtype, log_temp.shape)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        f = ctx.f
        dU, dV, dlt = ops.clip_bwd(f, ctx.col_lse, ctx.row_offset)
        if ctx.world > 1:
            # The returned loss is the GLOBAL loss; backward yields its exact gradient w.r.t. this rank's LOCAL
            # embeddings: d(img) is complete locally, d(txt) is the reduce-scatter(SUM) of the per-rank dV
            # (transpose of the forward all-gather).  d(log_temp) is this rank's share, like every replicated
            # parameter's gradient: the training loop all-reduces parameter gradients with SUM.
            dV_loc = torch.empty((ctx.Bl, dV.shape[1]), dtype=torch.float32, device=dV.device)
            dist.reduce_scatter_tensor(dV_loc, dV, op=dist.ReduceOp.SUM, group=ctx.group)
            v_loc = f["v"][ctx.row_offset:ctx.row_offset + ctx.Bl]
            rn_t = f["rn_t"][ctx.row_offset:ctx.row_offset + ctx.Bl]
        else:
            dV_loc, v_loc, rn_t = dV, f["v"], f["rn_t"]
        d_img = ops.l2norm_bwd(f["u"], f["rn_i"], dU)
        d_txt = ops.l2norm_bwd(v_loc.contiguous(), rn_t.contiguous(), dV_loc)
        g = gout.float()
        idt, tdt, ldt, lshape = ctx.in_dtypes
        return (d_img * g).to(idt), (d_txt * g).to(tdt), (dlt * g).reshape(lshape).to(ldt), None


def clip_contrastive_loss(img: torch.Tensor, txt: torch.Tensor, log_temp: torch.Tensor,
                          group: Optional["dist.ProcessGroup"] = None) -> torch.Tensor:
    """img (B_local, E), txt (B_local, E) embeddings (any float dtype, computed in bf16), log_temp scalar tensor.
    Returns the global-batch loss (identical on every rank)."""
    return _ClipLossFn.apply(img, txt, log_temp, group)
