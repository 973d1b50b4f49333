"""The C-ABI shared library builds, loads and exports every entry point include/w2v2.h declares (CPU only:
no kernel is launched; only argument-validation paths that return before touching the device are called)."""
import ctypes as C
import os
import re

import pytest

from wav2vec2 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "w2v2.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(w2v2_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_table_agree():
    assert _declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/w2v2.h but not exported"
    assert lib.w2v2_version() >= 100


def test_struct_layouts_match_header():
    # 64-bit pointers / int64 first, then int32 fields: sizes are what the C struct has on LP64
    assert C.sizeof(_lib.GemmArgs) == 8 * 8 + 11 * 4 + 4 + 8 * 8 + 3 * 8 + 8
    assert C.sizeof(_lib.PackJob) == 2 * 8 + 8 * 4
    assert C.sizeof(_lib.PosconvArgs) == 7 * 8 + 6 * 4 + 8 + 2 * 4


def test_argument_errors_are_reported_not_crashed(lib):
    assert lib.w2v2_gemm_bf16(None, None) < 0
    assert b"args is null" in lib.w2v2_last_error_string()
    args = _lib.GemmArgs()
    assert lib.w2v2_gemm_bf16(C.byref(args), None) < 0            # null operands
    assert lib.w2v2_posconv(None, None) < 0
    assert lib.w2v2_attn_fwd(None, None, 1, 1, 1, 64, None, None, None, 1, None) < 0
    assert lib.w2v2_ln_rows(None, None, None, 1e-5, 1, 768, 0, None, None, None, None) < 0
    assert lib.w2v2_ctc_workspace_bytes(2, 768, 256) == 2 * 2 * 768 * 513 * 4      # alpha and beta tables


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
