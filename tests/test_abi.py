"""The C-ABI library loads and exports every symbol include/forde_b200.h declares (no compute calls: CPU-only)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "forde_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(forde_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from forde_b200 import _lib
    return _lib.load()


def test_header_declares_the_path():
    syms = _declared_symbols()
    for must in ["forde_sl_up_fwd", "forde_sl_down_fwd", "forde_sl_dz_bwd", "forde_sl_dx_bwd", "forde_sl_dw_up",
                 "forde_sl_dw_down", "forde_stats_finalize", "forde_clip_fwd", "forde_clip_bwd"]:
        assert must in syms


def test_every_declared_symbol_is_exported_and_bound(lib):
    from forde_b200 import _lib
    syms = _declared_symbols()
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    unbound = [s for s in syms if s not in _lib.SIGNATURES]
    assert not unbound, f"declared in the header but not bound in _lib.SIGNATURES: {unbound}"
    extra = [s for s in _lib.SIGNATURES if s not in syms]
    assert not extra, f"bound in _lib.SIGNATURES but not declared in the header: {extra}"


def test_binding_arity_and_argument_kinds_match_the_header():
    """Every ctypes signature in _lib.SIGNATURES has as many arguments as the header declares, and pointers / integers
    / floating-point parameters line up position by position (a wrong count only shows up as a TypeError on the GPU)."""
    import ctypes
    from forde_b200 import _lib
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    decls = dict(re.findall(r"\b(forde_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src))
    for name, (_, args) in _lib.SIGNATURES.items():
        params = [p.strip() for p in decls[name].split(",")] if decls[name].strip() not in ("", "void") else []
        assert len(params) == len(args), f"{name}: header has {len(params)} parameters, binding {len(args)}"
        for i, (p, a) in enumerate(zip(params, args)):
            kind = "ptr" if "*" in p else ("fp" if re.match(r"(const\s+)?(float|double)\b", p) else "int")
            want = {"ptr": (ctypes.c_void_p, ctypes.c_char_p), "fp": (ctypes.c_float, ctypes.c_double),
                    "int": (ctypes.c_int, ctypes.c_longlong)}[kind]
            assert a in want, f"{name}: parameter {i} '{p}' is bound as {a}"
            if kind == "fp":
                assert (a is ctypes.c_double) == ("double" in p), f"{name}: parameter {i} '{p}' float/double mismatch"
            if kind == "int":
                assert (a is ctypes.c_longlong) == ("long long" in p), f"{name}: parameter {i} '{p}' int width mismatch"


def test_pure_host_entry_points(lib):
    assert lib.forde_version() >= 100
    assert lib.forde_error_string(0) == b"ok"
    assert b"alignment" in lib.forde_error_string(-2)
    assert lib.forde_sl_num_partials(0) == 0
    assert lib.forde_sl_num_partials(1) == 2
    assert lib.forde_sl_num_partials(256) == 2
    assert lib.forde_sl_num_partials(257) == 4
    assert lib.forde_sl_num_partials(50432) == 394
    assert lib.forde_sl_sign_words(0) == 0
    assert lib.forde_sl_sign_words(1) == 8
    assert lib.forde_sl_sign_words(257) == 16
    assert lib.forde_sl_sign_words(50432) == 1576          # = M / 32: one bit per sample


def test_argument_validation_needs_no_gpu(lib):
    # null pointers / bad shapes are rejected on the host before anything touches the device
    rc = lib.forde_sl_up_fwd(None, None, None, None, None, None, None, None, 16, 16, 16, 16, 1)
    assert rc == -1
    rc = lib.forde_sl_up_fwd_signs(None, None, None, None, None, None, None, None, None, 16, 16, 16, 16)
    assert rc == -1
    rc = lib.forde_sl_dz_bwd_signs(None, None, None, None, None, None, None, None, 16, 16, 16, 16)
    assert rc == -1
    rc = lib.forde_stats_finalize(None, None, None, 1, 16, 16, None, None, None, None, None, None, None)
    assert rc == -1
    rc = lib.forde_clip_fwd(None, None, None, None, 4, 4, 8, 0, *([None] * 10))
    assert rc == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from forde_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.FordeError, match="no CPU fallback"):
        _lib.load()


def test_ops_refuse_cpu_tensors(lib):
    import torch
    from forde_b200 import ops, _lib
    x = torch.zeros((8, 8), dtype=torch.bfloat16)
    with pytest.raises(_lib.FordeError, match="CUDA tensor"):
        ops.gemm_bf16(x, x, False, False)


def test_xla_ffi_shim_compiles_against_a_stub_of_the_ffi_types():
    """csrc/forde_xla_ffi.cc cannot meet jaxlib here; a syntax-only compile against tests/stubs/xla/ffi/api/ffi.h checks
    its includes, its calls into the C ABI (current signatures) and that every handler matches its binding."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not found")
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    cmd = [gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-DFORDE_WITH_XLA_FFI", "-I", os.path.join(ROOT, "tests", "stubs"),
           "-I", os.path.join(ROOT, "include"), "-I", cuda_inc,
           os.path.join(ROOT, "forde-vision-language_b200", "csrc", "forde_xla_ffi.cc")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
