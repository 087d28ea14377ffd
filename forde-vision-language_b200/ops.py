"""Thin torch-tensor wrappers over the C ABI (include/forde_b200.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every computation below is one of the
hand-written sm_100a kernels in csrc/.  Tensors must live on a CUDA device; nothing falls back to torch math.
"""
from __future__ import annotations

import os

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise _lib.FordeError(f"{name} must be a CUDA tensor (the FORDE hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.FordeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.FordeError(f"{name} must be contiguous")


def _req_rows(t: torch.Tensor, dtype, name: str) -> None:
    """A 2-D GEMM operand: unit column stride and a 16-byte aligned row pitch (TMA); rows need not be packed."""
    if not t.is_cuda:
        raise _lib.FordeError(f"{name} must be a CUDA tensor (the FORDE hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.FordeError(f"{name} must be {dtype}, got {t.dtype}")
    if t.dim() != 2 or t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        raise _lib.FordeError(f"{name} must be a 2-D tensor with contiguous rows")


# ---- optional per-kernel CUDA-event timing (bench.py's live roofline measurement); off by default -----------
_TIMING = None


def enable_timing(on: bool = True) -> None:
    """Record a CUDA-event pair on the launching stream around every StatefulLayer kernel launch."""
    global _TIMING
    _TIMING = {} if on else None


def timing_summary():
    """{kernel family: {"launches", "ms_total", "flops_total"}}; synchronises.  Call after the timed region."""
    out = {}
    if _TIMING is None:
        return out
    torch.cuda.synchronize()
    for name, recs in _TIMING.items():
        ms = sum(a.elapsed_time(b) for a, b, _ in recs)
        out[name] = {"launches": len(recs), "ms_total": ms, "flops_total": float(sum(f for _, _, f in recs))}
    return out


class _Timed:
    __slots__ = ("name", "flops", "a")

    def __init__(self, name, flops):
        self.name, self.flops = name, flops

    def __enter__(self):
        if _TIMING is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if _TIMING is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            _TIMING.setdefault(self.name, []).append((self.a, b, self.flops))
        return False


def launch_count() -> int:
    return int(_lib.load().forde_launch_count())


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def num_partials(M: int) -> int:
    return _lib.load().forde_sl_num_partials(M)


def sign_words(M: int) -> int:
    """Row pitch (u32 words) of the optional sign-bit buffer [H][sign_words(M)] for M samples."""
    return _lib.load().forde_sl_sign_words(M)


def sign_layout() -> int:
    """Bit mask.  Bit 0: 0 = bit j of a sign word is sample 32 w + j, 1 = interleaved (sample 2k -> bit k, 2k + 1 ->
    bit 16 + k).  Bit 1: 0 = buffer is [H, words], 1 = [words, H].  This build returns 3."""
    return _lib.load().forde_sl_sign_layout()


def sign_shape(M: int, H: int):
    """Shape of the int32 sign-bit buffer for M samples and H neurons in this build."""
    return (sign_words(M), H) if sign_layout() & 2 else (H, sign_words(M))


def sample_pitch(M: int) -> int:
    """Row pitch (in elements) of the neuron-major zT / hT / dzT buffers for M samples."""
    return round_up(M, 8)


# ------------------------------------------------------------------------------------------ generic GEMM
def gemm_bf16(A: torch.Tensor, B: torch.Tensor, a_mn_major: bool, b_mn_major: bool, bias_col=None, add_to=None,
              residual=None) -> torch.Tensor:
    """D[M,N] = sum_k A(m,k) B(n,k) (+ bias_col) -> bf16.  K-major operands are [rows, K]; MN-major operands are
    [K, rows].  With ``add_to`` (a contiguous bf16 [M, N] tensor, N % 8 == 0) the product is accumulated into it in
    place and it is returned; with ``residual`` (same shape) a NEW tensor D = product + bias + residual is returned."""
    _req_rows(A, torch.bfloat16, "A"); _req_rows(B, torch.bfloat16, "B")
    M, K = (A.shape[1], A.shape[0]) if a_mn_major else (A.shape[0], A.shape[1])
    N, Kb = (B.shape[1], B.shape[0]) if b_mn_major else (B.shape[0], B.shape[1])
    assert K == Kb, "reduction extents differ"
    if bias_col is not None:
        _req(bias_col, torch.float32, "bias_col")
    if add_to is not None:
        _req(add_to, torch.bfloat16, "add_to")
        assert tuple(add_to.shape) == (M, N)
        rc = _lib.load().forde_gemm_bf16_add(_stream(), A.data_ptr(), int(a_mn_major), A.stride(0), B.data_ptr(),
                                             int(b_mn_major), B.stride(0), add_to.data_ptr(), add_to.stride(0),
                                             add_to.data_ptr(), _ptr(bias_col), M, N, K)
        _lib.check(rc, "forde_gemm_bf16_add")
        return add_to
    ldd = round_up(N, 8)
    D = torch.empty((M, ldd), dtype=torch.bfloat16, device=A.device)
    if residual is not None:
        _req(residual, torch.bfloat16, "residual")
        assert tuple(residual.shape) == (M, N) and ldd == N
        rc = _lib.load().forde_gemm_bf16_add(_stream(), A.data_ptr(), int(a_mn_major), A.stride(0), B.data_ptr(),
                                             int(b_mn_major), B.stride(0), D.data_ptr(), ldd, residual.data_ptr(),
                                             _ptr(bias_col), M, N, K)
        _lib.check(rc, "forde_gemm_bf16_add")
        return D
    rc = _lib.load().forde_gemm_bf16(_stream(), A.data_ptr(), int(a_mn_major), A.stride(0), B.data_ptr(),
                                     int(b_mn_major), B.stride(0), D.data_ptr(), ldd, _ptr(bias_col), M, N, K)
    _lib.check(rc, "forde_gemm_bf16")
    return D[:, :N]


def gemm_f32(A, B, a_mn_major, b_mn_major, out=None, scale=1.0, accumulate=False, row_scale=None, col_scale=None):
    _req_rows(A, torch.bfloat16, "A"); _req_rows(B, torch.bfloat16, "B")
    M, K = (A.shape[1], A.shape[0]) if a_mn_major else (A.shape[0], A.shape[1])
    N, Kb = (B.shape[1], B.shape[0]) if b_mn_major else (B.shape[0], B.shape[1])
    assert K == Kb, "reduction extents differ"
    if out is None:
        out = torch.zeros((M, N), dtype=torch.float32, device=A.device)
    _req(out, torch.float32, "out")
    rc = _lib.load().forde_gemm_f32(_stream(), A.data_ptr(), int(a_mn_major), A.stride(0), B.data_ptr(),
                                    int(b_mn_major), B.stride(0), out.data_ptr(), out.stride(0), _ptr(row_scale),
                                    _ptr(col_scale), float(scale), int(accumulate), M, N, K)
    _lib.check(rc, "forde_gemm_f32")
    return out


# ------------------------------------------------------------------------------------------ StatefulLayer
def sl_up_fwd(x, w_up, b_up, assign_i8, zT, hT, act_partials, save_z: bool = True, sign_bits=None):
    """With ``sign_bits`` (int32, ``sign_shape(M, H)``) the save_z = 0 forward also records (z > 0) wherever it does not
    store the z box, for ``sl_dz_bwd(..., sign_bits=...)``."""
    M, d = x.shape
    H = w_up.shape[1]
    _req(x, torch.bfloat16, "x"); _req(w_up, torch.bfloat16, "w_up"); _req(assign_i8, torch.int8, "assign")
    _req(zT, torch.bfloat16, "zT"); _req(hT, torch.bfloat16, "hT")
    if b_up is not None:
        _req(b_up, torch.float32, "b_up")
    if sign_bits is not None:
        _req(sign_bits, torch.int32, "sign_bits")
        assert not save_z, "sign bits replace the z boxes the save_z = 0 forward skips"
        assert tuple(sign_bits.shape) == sign_shape(M, H) and sign_bits.is_contiguous()
        with _Timed(f"sl_up_fwd[{M}x{d}x{H}]", 2.0 * M * d * H):
            rc = _lib.load().forde_sl_up_fwd_signs(_stream(), x.data_ptr(), w_up.data_ptr(), _ptr(b_up),
                                                   assign_i8.data_ptr(), zT.data_ptr(), hT.data_ptr(),
                                                   sign_bits.data_ptr(), _ptr(act_partials), M, d, H, zT.stride(0))
        _lib.check(rc, "forde_sl_up_fwd_signs")
        return
    with _Timed(f"sl_up_fwd[{M}x{d}x{H}]", 2.0 * M * d * H):
        rc = _lib.load().forde_sl_up_fwd(_stream(), x.data_ptr(), w_up.data_ptr(), _ptr(b_up), assign_i8.data_ptr(),
                                         zT.data_ptr(), hT.data_ptr(), _ptr(act_partials), M, d, H, zT.stride(0),
                                         int(save_z))
    _lib.check(rc, "forde_sl_up_fwd")


def sl_down_fwd(hT, w_down, b_down, y, M, residual=None):
    """y = h . w_down + b_down (+ residual: the block's skip connection, added in the epilogue)."""
    H, d = w_down.shape
    _req(hT, torch.bfloat16, "hT"); _req(w_down, torch.bfloat16, "w_down"); _req(y, torch.bfloat16, "y")
    if b_down is not None:
        _req(b_down, torch.float32, "b_down")
    with _Timed(f"sl_down_fwd[{M}x{d}x{H}]", 2.0 * M * d * H):
        if residual is None:
            rc = _lib.load().forde_sl_down_fwd(_stream(), hT.data_ptr(), w_down.data_ptr(), _ptr(b_down), y.data_ptr(),
                                               M, d, H, hT.stride(0))
        else:
            _req(residual, torch.bfloat16, "residual")
            rc = _lib.load().forde_sl_down_fwd_add(_stream(), hT.data_ptr(), w_down.data_ptr(), _ptr(b_down),
                                                   residual.data_ptr(), y.data_ptr(), M, d, H, hT.stride(0))
    _lib.check(rc, "forde_sl_down_fwd")


def sl_dz_bwd(dy, w_down, zT, hT, assign_i8, dzT, grad_partials, sign_bits=None, keep_bits=None):
    """With ``sign_bits`` (written by ``sl_up_fwd(..., sign_bits=...)``) h^T is not read and may be None.
    With ``keep_bits`` (from ``sl_topk_sparsify``) dL/dz is zeroed where the forward's top-k dropped h, in the same
    epilogue, and the gradient partial sums are those of the masked dL/dz."""
    M, d = dy.shape
    H = w_down.shape[0]
    _req(dy, torch.bfloat16, "dy"); _req(zT, torch.bfloat16, "zT"); _req(dzT, torch.bfloat16, "dzT")
    if keep_bits is not None:
        _req(keep_bits, torch.int32, "keep_bits")
        assert tuple(keep_bits.shape) == ((H + 31) // 32, M)
        if sign_bits is not None:
            _req(sign_bits, torch.int32, "sign_bits")
        with _Timed(f"sl_dz_bwd[{M}x{d}x{H}]", 2.0 * M * d * H):
            rc = _lib.load().forde_sl_dz_bwd_topk(_stream(), dy.data_ptr(), w_down.data_ptr(), zT.data_ptr(), _ptr(hT),
                                                  _ptr(sign_bits), keep_bits.data_ptr(), assign_i8.data_ptr(),
                                                  dzT.data_ptr(), _ptr(grad_partials), M, d, H, zT.stride(0))
        _lib.check(rc, "forde_sl_dz_bwd_topk")
        return
    if sign_bits is not None:
        _req(sign_bits, torch.int32, "sign_bits")
        assert tuple(sign_bits.shape) == sign_shape(M, H) and sign_bits.is_contiguous()
        with _Timed(f"sl_dz_bwd[{M}x{d}x{H}]", 2.0 * M * d * H):
            rc = _lib.load().forde_sl_dz_bwd_signs(_stream(), dy.data_ptr(), w_down.data_ptr(), zT.data_ptr(),
                                                   sign_bits.data_ptr(), assign_i8.data_ptr(), dzT.data_ptr(),
                                                   _ptr(grad_partials), M, d, H, zT.stride(0))
        _lib.check(rc, "forde_sl_dz_bwd_signs")
        return
    with _Timed(f"sl_dz_bwd[{M}x{d}x{H}]", 2.0 * M * d * H):
        rc = _lib.load().forde_sl_dz_bwd(_stream(), dy.data_ptr(), w_down.data_ptr(), zT.data_ptr(), hT.data_ptr(),
                                         assign_i8.data_ptr(),
                                         dzT.data_ptr(), _ptr(grad_partials), M, d, H, zT.stride(0))
    _lib.check(rc, "forde_sl_dz_bwd")


def sl_dx_bwd(dzT, w_up, dx, M, residual=None):
    """dx = dz . w_up^T (+ residual, which may be dx itself)."""
    d, H = w_up.shape
    _req(dzT, torch.bfloat16, "dzT"); _req(dx, torch.bfloat16, "dx")
    if residual is not None:
        _req(residual, torch.bfloat16, "residual")
    with _Timed(f"sl_dx_bwd[{M}x{d}x{H}]", 2.0 * M * d * H):
        if residual is None:
            rc = _lib.load().forde_sl_dx_bwd(_stream(), dzT.data_ptr(), w_up.data_ptr(), dx.data_ptr(), M, d, H,
                                             dzT.stride(0))
        else:
            rc = _lib.load().forde_sl_dx_bwd_add(_stream(), dzT.data_ptr(), w_up.data_ptr(), residual.data_ptr(),
                                                 dx.data_ptr(), M, d, H, dzT.stride(0))
    _lib.check(rc, "forde_sl_dx_bwd")


def sl_dw_up(x, dzT, dw_up):
    M, d = x.shape
    H = dzT.shape[0]
    _req(x, torch.bfloat16, "x"); _req(dzT, torch.bfloat16, "dzT"); _req(dw_up, torch.float32, "dw_up")
    with _Timed(f"sl_dw_up[{M}x{d}x{H}]", 2.0 * M * d * H):
        rc = _lib.load().forde_sl_dw_up(_stream(), x.data_ptr(), dzT.data_ptr(), dw_up.data_ptr(), M, d, H, dzT.stride(0))
    _lib.check(rc, "forde_sl_dw_up")


def sl_dw_down(hT, dy, dw_down):
    M, d = dy.shape
    H = hT.shape[0]
    _req(hT, torch.bfloat16, "hT"); _req(dy, torch.bfloat16, "dy"); _req(dw_down, torch.float32, "dw_down")
    with _Timed(f"sl_dw_down[{M}x{d}x{H}]", 2.0 * M * d * H):
        rc = _lib.load().forde_sl_dw_down(_stream(), hT.data_ptr(), dy.data_ptr(), dw_down.data_ptr(), M, d, H,
                                          hT.stride(0))
    _lib.check(rc, "forde_sl_dw_down")


def sl_topk_sparsify(hT, M: int, k: int, want_indices: bool = True):
    """In-place per-token top-k of |h| over the neuron axis of hT [H][ld].  Returns (idx (M,k) i32 | None,
    hoyer (M,) f32, keep_bits (ceil(H/32), M) i32)."""
    _req(hT, torch.bfloat16, "hT")
    H = hT.shape[0]
    dev = hT.device
    idx = torch.empty((M, k), dtype=torch.int32, device=dev) if want_indices else None
    hoyer = torch.empty(M, dtype=torch.float32, device=dev)
    keep = torch.empty(((H + 31) // 32, M), dtype=torch.int32, device=dev)
    rc = _lib.load().forde_sl_topk_sparsify(_stream(), hT.data_ptr(), hT.stride(0), M, H, k, _ptr(idx), hoyer.data_ptr(),
                                            keep.data_ptr())
    _lib.check(rc, "forde_sl_topk_sparsify")
    return idx, hoyer, keep


def colsum_bf16(a, out):
    _req(a, torch.bfloat16, "a"); _req(out, torch.float32, "out")
    rc = _lib.load().forde_colsum_bf16(_stream(), a.data_ptr(), a.stride(0), out.data_ptr(), a.shape[0], a.shape[1])
    _lib.check(rc, "forde_colsum_bf16")


def stats_finalize(act_partials, grad_partials, P, H, M, step_stats=None, neuron_stats=None, moments=None,
                   sample_count=None, step_count=None, db_up=None, moments_lo=None):
    """``moments`` (+ optional ``moments_lo``, the double-single residue) accumulate the raw sums."""
    rc = _lib.load().forde_stats_finalize(_stream(), _ptr(act_partials), _ptr(grad_partials), P, H, M,
                                          _ptr(step_stats), _ptr(neuron_stats), _ptr(moments), _ptr(moments_lo),
                                          _ptr(sample_count), _ptr(step_count), _ptr(db_up))
    _lib.check(rc, "forde_stats_finalize")


def pooled_moments(stats_leaf) -> torch.Tensor:
    """fp64 value of a stats_buffer leaf's raw-moment accumulator (high + low fp32 parts)."""
    m = stats_leaf["moments"].double()
    lo = stats_leaf.get("moments_lo")
    return m + lo.double() if lo is not None else m


def stats_from_moments(moments, sample_count, out=None):
    """(H,5) statistics from pooled raw moments (fp64, see ``pooled_moments``; fp32 input is widened)."""
    if moments.dtype == torch.float32:
        moments = moments.double()
    _req(moments, torch.float64, "moments"); _req(sample_count, torch.int32, "sample_count")
    H = moments.shape[0]
    if out is None:
        out = torch.empty((H, 5), dtype=torch.float32, device=moments.device)
    rc = _lib.load().forde_stats_from_moments(_stream(), moments.data_ptr(), sample_count.data_ptr(), H, out.data_ptr())
    _lib.check(rc, "forde_stats_from_moments")
    return out


# ------------------------------------------------------------------------------------------ optimizer step
def optim_num_partials(n: int) -> int:
    return _lib.load().forde_optim_num_partials(n)


def grad_sqnorm(grads, partials, step_count=None):
    """partials[blocks] = per-block sums of g^2 over the flat gradient buffer (fp32 or bf16); step_count += 1."""
    if grads.dtype == torch.float32:
        _req(grads, torch.float32, "grads"); g32, g16 = grads.data_ptr(), 0
    else:
        _req(grads, torch.bfloat16, "grads"); g32, g16 = 0, grads.data_ptr()
    _req(partials, torch.float32, "partials")
    rc = _lib.load().forde_grad_sqnorm(_stream(), g32, g16, grads.numel(), partials.data_ptr(), _ptr(step_count))
    _lib.check(rc, "forde_grad_sqnorm")


def adamw_clip_step(p32, grads, mu, nu, p16, partials, step_count, grad_norm_out, max_norm, lr, b1, b2, eps, wd):
    """optax.chain(clip_by_global_norm(max_norm), adamw(lr, b1, b2, eps, weight_decay=wd)) over flat buffers, plus the
    bf16 mirror refresh; grads fp32 or bf16.  Call grad_sqnorm first (it also advances step_count)."""
    _req(p32, torch.float32, "p32"); _req(mu, torch.float32, "mu"); _req(nu, torch.float32, "nu")
    if grads.dtype == torch.float32:
        _req(grads, torch.float32, "grads"); g32, g16 = grads.data_ptr(), 0
    else:
        _req(grads, torch.bfloat16, "grads"); g32, g16 = 0, grads.data_ptr()
    if p16 is not None:
        _req(p16, torch.bfloat16, "p16")
    _req(step_count, torch.int32, "step_count")
    n = p32.numel()
    assert grads.numel() == n and mu.numel() == n and nu.numel() == n
    rc = _lib.load().forde_adamw_clip_step(_stream(), p32.data_ptr(), g32, g16, mu.data_ptr(), nu.data_ptr(), _ptr(p16),
                                           n, partials.data_ptr(), partials.numel(), float(max_norm), float(lr),
                                           float(b1), float(b2), float(eps), float(wd), step_count.data_ptr(),
                                           _ptr(grad_norm_out))
    _lib.check(rc, "forde_adamw_clip_step")


def cast_f32_bf16(src, dst):
    _req(src, torch.float32, "src"); _req(dst, torch.bfloat16, "dst")
    assert src.numel() == dst.numel()
    rc = _lib.load().forde_cast_f32_bf16(_stream(), src.data_ptr(), dst.data_ptr(), src.numel())
    _lib.check(rc, "forde_cast_f32_bf16")


# ------------------------------------------------------------------------------------------ CLIP loss
def clip_fwd(img, txt_all, log_temp, row_offset: int):
    """Launch the forward kernels; returns a dict of the forward intermediates (all device tensors)."""
    _req(img, torch.bfloat16, "img"); _req(txt_all, torch.bfloat16, "txt"); _req(log_temp, torch.float32, "log_temp")
    Bl, E = img.shape
    Bg = txt_all.shape[0]
    dev = img.device
    f32 = dict(dtype=torch.float32, device=dev)
    r = {
        "u": torch.empty_like(img), "v": torch.empty_like(txt_all),
        "rn_i": torch.empty(Bl, **f32), "rn_t": torch.empty(Bg, **f32), "scale": torch.empty(1, **f32),
        "S": torch.empty((Bl, Bg), **f32), "row_lse": torch.empty(Bl, **f32),
        "col_max": torch.empty(Bg, **f32), "col_sum": torch.empty(Bg, **f32), "diag": torch.empty(Bl, **f32),
    }
    rc = _lib.load().forde_clip_fwd(_stream(), img.data_ptr(), txt_all.data_ptr(), log_temp.data_ptr(), Bl, Bg, E,
                                    int(row_offset), r["u"].data_ptr(), r["v"].data_ptr(), r["rn_i"].data_ptr(),
                                    r["rn_t"].data_ptr(), r["scale"].data_ptr(), r["S"].data_ptr(),
                                    r["row_lse"].data_ptr(), r["col_max"].data_ptr(), r["col_sum"].data_ptr(),
                                    r["diag"].data_ptr())
    _lib.check(rc, "forde_clip_fwd")
    return r


def clip_col_rescale(col_max_local, col_max_global, col_sum):
    rc = _lib.load().forde_clip_col_rescale(_stream(), col_max_local.data_ptr(), col_max_global.data_ptr(),
                                            col_sum.data_ptr(), col_sum.numel())
    _lib.check(rc, "forde_clip_col_rescale")


def clip_loss(fwd, col_max, col_sum, row_offset: int):
    """Returns (loss_share[1], col_lse[Bg]); loss_share sums to the loss over ranks."""
    Bl = fwd["row_lse"].numel()
    Bg = col_max.numel()
    col_lse = torch.empty_like(col_max)
    loss = torch.zeros(1, dtype=torch.float32, device=col_max.device)
    rc = _lib.load().forde_clip_loss(_stream(), fwd["row_lse"].data_ptr(), fwd["diag"].data_ptr(), col_max.data_ptr(),
                                     col_sum.data_ptr(), Bl, Bg, int(row_offset), col_lse.data_ptr(), loss.data_ptr())
    _lib.check(rc, "forde_clip_loss")
    return loss, col_lse


def clip_bwd(fwd, col_lse, row_offset: int):
    u, v, S = fwd["u"], fwd["v"], fwd["S"]
    Bl, E = u.shape
    Bg = v.shape[0]
    dev = u.device
    dS = torch.zeros((Bl, round_up(Bg, 8)), dtype=torch.bfloat16, device=dev)
    dU = torch.empty((Bl, E), dtype=torch.float32, device=dev)
    dV = torch.empty((Bg, E), dtype=torch.float32, device=dev)
    dlog_temp = torch.zeros(1, dtype=torch.float32, device=dev)
    rc = _lib.load().forde_clip_bwd(_stream(), u.data_ptr(), v.data_ptr(), S.data_ptr(), fwd["row_lse"].data_ptr(),
                                    col_lse.data_ptr(), fwd["scale"].data_ptr(), Bl, Bg, E, int(row_offset),
                                    dS.data_ptr(), dU.data_ptr(), dV.data_ptr(), dlog_temp.data_ptr())
    _lib.check(rc, "forde_clip_bwd")
    return dU, dV, dlog_temp


def l2norm_bwd(u, inv_norm, dU):
    B, E = u.shape
    out = torch.empty((B, E), dtype=torch.float32, device=u.device)
    rc = _lib.load().forde_l2norm_bwd(_stream(), u.data_ptr(), inv_norm.data_ptr(), dU.data_ptr(), out.data_ptr(), B, E)
    _lib.check(rc, "forde_l2norm_bwd")
    return out


# ------------------------------------------------------------------------------------------ add + LayerNorm
def add_layernorm_fwd(x, r, gamma, beta, eps: float = 1e-6):
    """out = LayerNorm(x + r) * gamma + beta over the last axis; returns (out, mean, rstd)."""
    _req(x, torch.bfloat16, "x"); _req(gamma, torch.float32, "gamma"); _req(beta, torch.float32, "beta")
    if r is not None:
        _req(r, torch.bfloat16, "r")
    cols = x.shape[-1]
    rows = x.numel() // cols
    out = torch.empty_like(x)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    rc = _lib.load().forde_add_layernorm_fwd(_stream(), x.data_ptr(), _ptr(r), gamma.data_ptr(), beta.data_ptr(),
                                             out.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, cols, float(eps))
    _lib.check(rc, "forde_add_layernorm_fwd")
    return out, mean, rstd


def add_layernorm_bwd(dy, x, r, gamma, mean, rstd, d_gamma=None, d_beta=None):
    """returns (ds, d_gamma, d_beta); ds is the gradient of both x and r.  d_gamma / d_beta (fp32, accumulated
    into) may be supplied."""
    _req(dy, torch.bfloat16, "dy")
    cols = x.shape[-1]
    rows = x.numel() // cols
    ds = torch.empty_like(x)
    nb = _lib.load().forde_ln_num_partials(rows)
    partials = torch.empty((nb, 2, cols), dtype=torch.float32, device=x.device)
    if d_gamma is None:
        d_gamma = torch.zeros(cols, dtype=torch.float32, device=x.device)
    if d_beta is None:
        d_beta = torch.zeros(cols, dtype=torch.float32, device=x.device)
    rc = _lib.load().forde_add_layernorm_bwd(_stream(), dy.data_ptr(), x.data_ptr(), _ptr(r), gamma.data_ptr(),
                                             mean.data_ptr(), rstd.data_ptr(), ds.data_ptr(), partials.data_ptr(),
                                             d_gamma.data_ptr(), d_beta.data_ptr(), rows, cols)
    _lib.check(rc, "forde_add_layernorm_bwd")
    return ds, d_gamma, d_beta


# ------------------------------------------------------------------------------------------ self-attention core
def attn_short_max_tokens() -> int:
    """Longest sequence the one-tile kernels (csrc/forde_attn.cu) take."""
    return _lib.load().forde_attn_max_tokens()


def attn_max_tokens() -> int:
    """Longest sequence any of this library's attention kernels takes (csrc/forde_attn2.cu)."""
    return _lib.load().forde_attn2_max_tokens()


def _attn_pitch(*tensors) -> int:
    """The attention entry points take one row pitch for every operand and result: all of them must be row-major with
    unit column stride and the SAME row stride (raises instead of reading or writing with the wrong pitch)."""
    ld = tensors[0].stride(0)
    for t in tensors:
        if t.dim() != 2 or t.stride(1) != 1 or t.stride(0) != ld:
            raise _lib.FordeError(f"attention operands must share one row pitch and have unit column stride "
                                  f"(got strides {[tuple(x.stride()) for x in tensors]})")
    return ld


def attn_fwd(q, k, v, B: int, T: int, heads: int, blocked=None):
    """q, k, v: bf16 [B*T, heads*64].  Returns (o [B*T, heads*64] bf16, lse [B, heads, T] fp32).
    T <= 128 runs the one-tile kernel, longer sequences (<= 256) the blocked one; ``blocked`` forces the choice."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, torch.bfloat16, n)
    assert q.shape == k.shape == v.shape == (B * T, heads * 64)
    ld = _attn_pitch(q, k, v)
    o = torch.empty_strided(q.shape, (ld, 1), dtype=q.dtype, device=q.device)   # the C ABI has ONE row pitch for all
    lse = torch.empty((B, heads, T), dtype=torch.float32, device=q.device)
    if blocked is None:
        blocked = T > attn_short_max_tokens()
    fn = _lib.load().forde_attn2_fwd if blocked else _lib.load().forde_attn_fwd
    rc = fn(_stream(), q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(), B, T, heads, ld)
    _lib.check(rc, "forde_attn2_fwd" if blocked else "forde_attn_fwd")
    return o, lse


def attn_bwd(q, k, v, d_o, lse, B: int, T: int, heads: int, o=None, blocked=None):
    """``o`` (the forward output) is read by the blocked kernel only (D = rowsum(dO * o))."""
    _req(d_o, torch.bfloat16, "d_o"); _req(lse, torch.float32, "lse")
    ld = _attn_pitch(q, k, v, d_o, *(() if o is None else (o,)))
    dq, dk, dv = (torch.empty_strided(q.shape, (ld, 1), dtype=q.dtype, device=q.device) for _ in range(3))
    if blocked is None:
        blocked = T > attn_short_max_tokens()
    if blocked:
        if o is None:
            raise _lib.FordeError("attn_bwd: the blocked kernel (T > 128) needs the forward output `o`")
        _req(o, torch.bfloat16, "o")
        rc = _lib.load().forde_attn2_bwd(_stream(), q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(),
                                         d_o.data_ptr(), lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                         B, T, heads, ld)
        _lib.check(rc, "forde_attn2_bwd")
        return dq, dk, dv
    rc = _lib.load().forde_attn_bwd(_stream(), q.data_ptr(), k.data_ptr(), v.data_ptr(), d_o.data_ptr(), lse.data_ptr(),
                                    dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, T, heads, ld)
    _lib.check(rc, "forde_attn_bwd")
    return dq, dk, dv


# ------------------------------------------------------------------------------------------ tower front ends
def patchify(images, patch: int):
    """images (B, S, S, 3) bf16 NHWC -> patch rows (B * g * g, round_up(p*p*3, 8)) bf16, first p*p*3 columns valid."""
    _req(images, torch.bfloat16, "images")
    B, S = images.shape[0], images.shape[1]
    g = S // patch
    k = patch * patch * 3
    ld = round_up(k, 8)
    out = torch.empty((B * g * g, ld), dtype=torch.bfloat16, device=images.device)
    rc = _lib.load().forde_patchify(_stream(), images.data_ptr(), out.data_ptr(), ld, B, S, patch)
    _lib.check(rc, "forde_patchify")
    return out[:, :k]


def vision_assemble_fwd(tok, cls16, pos16, B: int):
    _req(tok, torch.bfloat16, "patch tokens"); _req(cls16, torch.bfloat16, "cls"); _req(pos16, torch.bfloat16, "pos_embed")
    G, w = tok.shape[0] // B, tok.shape[1]
    x = torch.empty((B, G + 1, w), dtype=torch.bfloat16, device=tok.device)
    rc = _lib.load().forde_vision_assemble_fwd(_stream(), tok.data_ptr(), cls16.data_ptr(), pos16.data_ptr(), x.data_ptr(),
                                               B, G, w)
    _lib.check(rc, "forde_vision_assemble_fwd")
    return x


def vision_assemble_bwd(dx, d_cls, d_pos):
    """dx (B, G+1, w) bf16 -> d_patch_tok (B*G, w) bf16; d_cls (w,), d_pos (G+1, w) fp32 accumulated into."""
    _req(dx, torch.bfloat16, "dx"); _req(d_cls, torch.float32, "d_cls"); _req(d_pos, torch.float32, "d_pos")
    B, T, w = dx.shape
    d_tok = torch.empty((B * (T - 1), w), dtype=torch.bfloat16, device=dx.device)
    rc = _lib.load().forde_vision_assemble_bwd(_stream(), dx.data_ptr(), d_tok.data_ptr(), d_cls.data_ptr(),
                                               d_pos.data_ptr(), B, T - 1, w)
    _lib.check(rc, "forde_vision_assemble_bwd")
    return d_tok


def embed_tokens_fwd(ids, table16, pos16):
    _req(ids, torch.int32, "input_ids"); _req(table16, torch.bfloat16, "token_embed"); _req(pos16, torch.bfloat16, "pos_embed")
    B, T = ids.shape
    V, w = table16.shape
    x = torch.empty((B, T, w), dtype=torch.bfloat16, device=ids.device)
    rc = _lib.load().forde_embed_tokens_fwd(_stream(), ids.data_ptr(), table16.data_ptr(), pos16.data_ptr(), x.data_ptr(),
                                            B, T, w, V)
    _lib.check(rc, "forde_embed_tokens_fwd")
    return x


def embed_tokens_bwd(ids, dx, d_table, d_pos):
    _req(ids, torch.int32, "input_ids"); _req(dx, torch.bfloat16, "dx")
    _req(d_table, torch.float32, "d_table"); _req(d_pos, torch.float32, "d_pos")
    B, T, w = dx.shape
    rc = _lib.load().forde_embed_tokens_bwd(_stream(), ids.data_ptr(), dx.data_ptr(), d_table.data_ptr(), d_pos.data_ptr(),
                                            B, T, w, d_table.shape[0])
    _lib.check(rc, "forde_embed_tokens_bwd")


# ------------------------------------------------------------------------------------------ MoE variant
def moe_capacity_rows(M: int, k: int, E: int) -> int:
    return _lib.load().forde_moe_capacity_rows(M, k, E)


def moe_route(x, w_r16, b_r, k: int, aux_weight: float, expert_usage=None, step_count=None):
    """Router + top-k + histogram + aux loss + routing plan.  Returns a dict of device tensors (the routing state)."""
    _req(x, torch.bfloat16, "x"); _req(w_r16, torch.bfloat16, "w_r"); _req(b_r, torch.float32, "b_r")
    M, d = x.shape
    E = w_r16.shape[1]
    lib = _lib.load()
    R = lib.forde_moe_capacity_rows(M, k, E)
    dev = x.device
    i32 = dict(dtype=torch.int32, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    r = {
        "M": M, "d": d, "E": E, "k": k, "R": R,
        "logits": torch.empty((M, E), **f32), "probs": torch.empty((M, E), **f32),
        "topk_idx": torch.empty((M, k), **i32), "topk_w": torch.empty((M, k), **f32),
        "counts": torch.empty(E, **i32),
        "prob_partials": torch.empty((lib.forde_moe_num_prob_partials(M), E), **f32),
        "plan_scratch": torch.empty(lib.forde_moe_plan_scratch_ints(M, k, E), **i32),
        "seg_offsets": torch.empty(E + 1, **i32), "n_blocks": torch.empty(1, **i32),
        "block_expert": torch.empty(R // 128, **i32), "slot_of": torch.empty(M * k, **i32),
        "src_of_slot": torch.empty(R, **i32), "aux_loss": torch.empty(1, **f32),
    }
    rc = lib.forde_moe_route(_stream(), x.data_ptr(), w_r16.data_ptr(), b_r.data_ptr(), M, d, E, k, float(aux_weight),
                             r["logits"].data_ptr(), r["probs"].data_ptr(), r["topk_idx"].data_ptr(),
                             r["topk_w"].data_ptr(), r["counts"].data_ptr(), r["prob_partials"].data_ptr(),
                             r["plan_scratch"].data_ptr(), r["seg_offsets"].data_ptr(), r["n_blocks"].data_ptr(), r["block_expert"].data_ptr(),
                             r["slot_of"].data_ptr(), r["src_of_slot"].data_ptr(), r["aux_loss"].data_ptr(),
                             _ptr(expert_usage), _ptr(step_count))
    _lib.check(rc, "forde_moe_route")
    return r


def moe_gather(x, route):
    xg = torch.empty((route["R"], x.shape[1]), dtype=torch.bfloat16, device=x.device)
    rc = _lib.load().forde_moe_gather(_stream(), x.data_ptr(), route["src_of_slot"].data_ptr(), route["k"],
                                      xg.data_ptr(), route["R"], x.shape[1])
    _lib.check(rc, "forde_moe_gather")
    return xg


def moe_experts_fwd(xg, w_up16, b_up, w_down16, b_down, route):
    E, d, H = w_up16.shape
    R = route["R"]
    _req(w_up16, torch.bfloat16, "w_up_all"); _req(w_down16, torch.bfloat16, "w_down_all")
    _req(b_up, torch.float32, "b_up_all"); _req(b_down, torch.float32, "b_down_all")
    zg = torch.empty((R, H), dtype=torch.bfloat16, device=xg.device)
    hg = torch.empty((R, H), dtype=torch.bfloat16, device=xg.device)
    yg = torch.empty((R, d), dtype=torch.bfloat16, device=xg.device)
    rc = _lib.load().forde_moe_experts_fwd(_stream(), xg.data_ptr(), w_up16.data_ptr(), b_up.data_ptr(),
                                           w_down16.data_ptr(), b_down.data_ptr(), route["n_blocks"].data_ptr(),
                                           route["block_expert"].data_ptr(), zg.data_ptr(), hg.data_ptr(),
                                           yg.data_ptr(), R, d, H, E)
    _lib.check(rc, "forde_moe_experts_fwd")
    return zg, hg, yg


def moe_combine(yg, route, weighted: bool = True):
    M, k, d = route["M"], route["k"], yg.shape[1]
    out = torch.empty((M, d), dtype=torch.bfloat16, device=yg.device)
    rc = _lib.load().forde_moe_combine(_stream(), yg.data_ptr(), route["slot_of"].data_ptr(),
                                       route["topk_w"].data_ptr() if weighted else 0, out.data_ptr(), M, k, d)
    _lib.check(rc, "forde_moe_combine")
    return out


def moe_scatter(dout, route):
    _req(dout, torch.bfloat16, "dout")
    R, d = route["R"], dout.shape[1]
    dyg = torch.empty((R, d), dtype=torch.bfloat16, device=dout.device)
    rc = _lib.load().forde_moe_scatter(_stream(), dout.data_ptr(), route["src_of_slot"].data_ptr(),
                                       route["topk_w"].data_ptr(), route["k"], dyg.data_ptr(), R, d)
    _lib.check(rc, "forde_moe_scatter")
    return dyg


def moe_gate_grad(dout, yg, route):
    M, k = route["M"], route["k"]
    d_w = torch.empty((M, k), dtype=torch.float32, device=dout.device)
    rc = _lib.load().forde_moe_gate_grad(_stream(), dout.data_ptr(), yg.data_ptr(), route["slot_of"].data_ptr(),
                                         d_w.data_ptr(), M, k, dout.shape[1])
    _lib.check(rc, "forde_moe_gate_grad")
    return d_w


def moe_experts_bwd(xg, dyg, w_up16, w_down16, zg, hg, route):
    E, d, H = w_up16.shape
    R = route["R"]
    dev = xg.device
    dzg = torch.empty((R, H), dtype=torch.bfloat16, device=dev)
    dxg = torch.empty((R, d), dtype=torch.bfloat16, device=dev)
    dw_up = torch.empty((E, d, H), dtype=torch.float32, device=dev)
    dw_down = torch.empty((E, H, d), dtype=torch.float32, device=dev)
    db_up = torch.zeros((E, H), dtype=torch.float32, device=dev)
    db_down = torch.zeros((E, d), dtype=torch.float32, device=dev)
    rc = _lib.load().forde_moe_experts_bwd(_stream(), xg.data_ptr(), dyg.data_ptr(), w_up16.data_ptr(),
                                           w_down16.data_ptr(), zg.data_ptr(), hg.data_ptr(),
                                           route["n_blocks"].data_ptr(), route["block_expert"].data_ptr(),
                                           route["seg_offsets"].data_ptr(), dzg.data_ptr(), dxg.data_ptr(),
                                           dw_up.data_ptr(), db_up.data_ptr(), dw_down.data_ptr(), db_down.data_ptr(),
                                           R, d, H, E)
    _lib.check(rc, "forde_moe_experts_bwd")
    return dxg, dw_up, db_up, dw_down, db_down


def moe_router_bwd(x, w_r16, route, d_topk_w, d_aux, d_probs, aux_weight: float, dx):
    """dx (bf16, in place) += dlogits . w_r^T ; returns (dw_r [d][E], db_r [E]) fp32."""
    M, d, E, k = route["M"], route["d"], route["E"], route["k"]
    dw_r = torch.zeros((d, E), dtype=torch.float32, device=x.device)
    db_r = torch.zeros(E, dtype=torch.float32, device=x.device)
    rc = _lib.load().forde_moe_router_bwd(_stream(), x.data_ptr(), w_r16.data_ptr(), route["probs"].data_ptr(),
                                          route["topk_idx"].data_ptr(), route["topk_w"].data_ptr(),
                                          d_topk_w.data_ptr(), route["counts"].data_ptr(), _ptr(d_aux), _ptr(d_probs),
                                          float(aux_weight), M, d, E, k, dx.data_ptr(), dw_r.data_ptr(),
                                          db_r.data_ptr())
    _lib.check(rc, "forde_moe_router_bwd")
    return dw_r, db_r
