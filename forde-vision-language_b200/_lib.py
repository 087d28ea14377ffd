"""ctypes binding of libforde_b200.so (C ABI in include/forde_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no CPU or PyTorch
fallback: if the shared object is missing, or an entry point returns a non-zero code, we raise.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FORDE_LIB", os.path.join(_HERE, "libforde_b200.so"))  # FORDE_LIB: A/B builds (tools/)

_vp = ctypes.c_void_p
_i = ctypes.c_int
_ll = ctypes.c_longlong
_f = ctypes.c_float

# name -> (restype, argtypes).  Mirrors include/forde_b200.h one to one; tests/test_abi.py checks that every
# symbol declared in the header is exported and listed here.
SIGNATURES = {
    "forde_version": (_i, []),
    "forde_error_string": (ctypes.c_char_p, [_i]),
    "forde_last_cuda_error": (ctypes.c_char_p, []),
    "forde_sm_count": (_i, []),
    "forde_launch_count": (ctypes.c_ulonglong, []),
    "forde_sl_num_partials": (_i, [_i]),
    "forde_sl_sign_words": (_i, [_i]),
    "forde_sl_sign_layout": (_i, []),
    "forde_sl_up_fwd_signs": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _ll]),
    "forde_sl_dz_bwd_signs": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _ll]),
    "forde_gemm_bf16": (_i, [_vp, _vp, _i, _ll, _vp, _i, _ll, _vp, _ll, _vp, _i, _i, _i]),
    "forde_gemm_bf16_add": (_i, [_vp, _vp, _i, _ll, _vp, _i, _ll, _vp, _ll, _vp, _vp, _i, _i, _i]),
    "forde_gemm_f32": (_i, [_vp, _vp, _i, _ll, _vp, _i, _ll, _vp, _ll, _vp, _vp, _f, _i, _i, _i, _i]),
    "forde_sl_up_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _ll, _i]),
    "forde_sl_down_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _ll]),
    "forde_sl_down_fwd_add": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _ll]),
    "forde_sl_fwd": (_i, [_vp] * 11 + [_i, _i, _i, _ll, _i]),
    "forde_sl_dz_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _ll]),
    "forde_sl_dx_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _ll]),
    "forde_sl_dx_bwd_add": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _ll]),
    "forde_sl_dw_up": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _ll]),
    "forde_sl_dw_down": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _ll]),
    "forde_colsum_bf16": (_i, [_vp, _vp, _ll, _vp, _i, _i]),
    "forde_sl_bwd": (_i, [_vp] * 15 + [_i, _i, _i, _ll]),
    "forde_sl_topk_sparsify": (_i, [_vp, _vp, _ll, _i, _i, _i, _vp, _vp, _vp]),
    "forde_sl_dz_bwd_topk": (_i, [_vp] * 10 + [_i, _i, _i, _ll]),
    "forde_stats_finalize": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "forde_stats_from_moments": (_i, [_vp, _vp, _vp, _i, _vp]),
    "forde_clip_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i] + [_vp] * 10),
    "forde_clip_col_rescale": (_i, [_vp, _vp, _vp, _vp, _i]),
    "forde_clip_loss": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "forde_clip_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "forde_l2norm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i]),
    "forde_moe_capacity_rows": (_i, [_i, _i, _i]),
    "forde_moe_num_prob_partials": (_i, [_i]),
    "forde_moe_route": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f] + [_vp] * 15),
    "forde_moe_plan_scratch_ints": (_i, [_i, _i, _i]),
    "forde_moe_gather": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i]),
    "forde_moe_scatter": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _i]),
    "forde_moe_combine": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "forde_moe_gate_grad": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "forde_moe_router_bwd": (_i, [_vp] * 10 + [_f, _i, _i, _i, _i, _vp, _vp, _vp]),
    "forde_moe_experts_fwd": (_i, [_vp] * 11 + [_i, _i, _i, _i]),
    "forde_moe_experts_bwd": (_i, [_vp] * 16 + [_i, _i, _i, _i]),
    "forde_ln_num_partials": (_i, [_i]),
    "forde_add_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f]),
    "forde_add_layernorm_bwd": (_i, [_vp] * 11 + [_i, _i]),
    "forde_optim_num_partials": (_i, [_ll]),
    "forde_grad_sqnorm": (_i, [_vp, _vp, _vp, _ll, _vp, _vp]),
    "forde_adamw_clip_step": (_i, [_vp] * 7 + [_ll, _vp, _i] + [ctypes.c_double] * 6 + [_vp, _vp]),
    "forde_cast_f32_bf16": (_i, [_vp, _vp, _vp, _ll]),
    "forde_patchify": (_i, [_vp, _vp, _vp, _ll, _i, _i, _i]),
    "forde_vision_assemble_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "forde_vision_assemble_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i]),
    "forde_embed_tokens_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "forde_embed_tokens_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "forde_attn_max_tokens": (_i, []),
    "forde_attn_fwd": (_i, [_vp] * 6 + [_i, _i, _i, _ll]),
    "forde_attn_bwd": (_i, [_vp] * 9 + [_i, _i, _i, _ll]),
    "forde_attn2_max_tokens": (_i, []),
    "forde_attn2_fwd": (_i, [_vp] * 6 + [_i, _i, _i, _ll]),
    "forde_attn2_bwd": (_i, [_vp] * 10 + [_i, _i, _i, _ll]),
}


class FordeError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises FordeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FordeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback for the FORDE hot path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        lib = load()
        msg = lib.forde_error_string(rc).decode()
        detail = lib.forde_last_cuda_error().decode()
        raise FordeError(f"{what} failed with code {rc}: {msg}" + (f" [{detail}]" if detail else ""))
