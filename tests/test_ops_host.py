"""Host-side argument checks of forde_b200.ops that need no GPU."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_attention_wrappers_refuse_operands_with_different_row_pitches():
    """include/forde_b200.h: forde_attn*_fwd/_bwd take ONE `ld` for q, k, v, o, d_o and the gradients.  A strided view
    (e.g. a column slice of a packed q|k|v tensor) next to a dense tensor must raise, not be read with the wrong pitch."""
    import forde_b200  # noqa: F401
    from forde_b200 import _lib, ops

    dense = torch.zeros((8, 128), dtype=torch.bfloat16)
    sliced = torch.zeros((8, 384), dtype=torch.bfloat16)[:, 128:256]
    assert ops._attn_pitch(dense, dense, dense) == 128
    assert ops._attn_pitch(sliced, sliced, sliced) == 384          # one shared pitch is fine
    with pytest.raises(_lib.FordeError):
        ops._attn_pitch(dense, sliced, dense)
    with pytest.raises(_lib.FordeError):
        ops._attn_pitch(dense.t(), dense.t(), dense.t())           # column stride must be 1
