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


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every argument struct, as gcc lays out include/w2v2.h, equal the ctypes mirrors field by field."""
    import subprocess
    structs = {"w2v2_gemm_args": _lib.GemmArgs, "w2v2_posconv_args": _lib.PosconvArgs, "w2v2_pack_job": _lib.PackJob}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {']
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


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
