"""ctypes binding of ``libw2v2_sm100.so`` (declared in ``include/w2v2.h``).

There is deliberately no fallback: if the library is missing, or a call returns a non-zero
status, this module raises.  A CPU / eager path standing in for the CUDA kernels would void
every parity and performance claim made for this package.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)                      # .../gsoc-wav2vec2_b200
LIB_PATH = os.path.join(_ROOT, "lib", "libw2v2_sm100.so")
CSRC_DIR = os.path.join(_ROOT, "csrc")

GEMM_GELU = 1
GEMM_MN_MAJOR = 2
GEMM_GELU_TANH = 4
MODE_BF16, MODE_BF16X3, MODE_FP16, MODE_FP16X3, MODE_FP16F8 = 1, 3, 17, 19, 25      # W2V2_MODE_*
OUT_BF16, OUT_FP16, OUT_FP16F8 = 0, 1, 2                                              # W2V2_OUT_*


class GemmArgs(C.Structure):
    """Mirror of ``w2v2_gemm_args`` (include/w2v2.h)."""
    _fields_ = [
        ("a_hi", C.c_void_p), ("a_lo", C.c_void_p),
        ("a_row_len", C.c_int64), ("a_rows", C.c_int64), ("a_row_stride", C.c_int64), ("a_batch_stride", C.c_int64),
        ("w_hi", C.c_void_p), ("w_lo", C.c_void_p),
        ("w_rows", C.c_int32), ("K", C.c_int32), ("N", C.c_int32), ("rows_per_batch", C.c_int32),
        ("batch", C.c_int32), ("passes", C.c_int32), ("kb_split", C.c_int32), ("block_n", C.c_int32),
        ("max_ctas", C.c_int32), ("cluster", C.c_int32), ("flags", C.c_uint32),
        ("bias", C.c_void_p), ("scale", C.c_void_p), ("bias_batch_stride", C.c_int64),
        ("residual", C.c_void_p), ("row_valid", C.c_void_p),
        ("out_f32", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
        ("res_ln_stats", C.c_void_p), ("res_ln_gamma", C.c_void_p), ("res_ln_beta", C.c_void_p),
        ("w_row_stride", C.c_int64),
        ("row_replace_mask", C.c_void_p), ("row_replace_value", C.c_void_p),
        ("drop_p", C.c_float), ("drop_site", C.c_uint32), ("drop_seed", C.c_uint64),
        ("out_format", C.c_int32), ("ln_fold_parts", C.c_int32), ("ln_fold_stats", C.c_void_p),
        ("ln_eps", C.c_float), ("res_ln_parts", C.c_int32), ("row_stats_out", C.c_void_p),
        ("row_stats_final", C.c_void_p), ("row_stats_counter", C.c_void_p),
    ]


class PosconvArgs(C.Structure):
    """Mirror of ``w2v2_posconv_args`` (include/w2v2.h)."""
    _fields_ = [
        ("x_hi", C.c_void_p), ("x_lo", C.c_void_p), ("w_hi", C.c_void_p), ("w_lo", C.c_void_p),
        ("bias", C.c_void_p), ("resid", C.c_void_p), ("out_f32", C.c_void_p),
        ("batch", C.c_int32), ("frames", C.c_int32), ("hidden", C.c_int32), ("groups", C.c_int32),
        ("ktaps", C.c_int32), ("passes", C.c_int32),
        ("pre_out", C.c_void_p), ("shift", C.c_int32), ("linear", C.c_int32), ("gelu_approx", C.c_int32),
    ]


class PackJob(C.Structure):
    """Mirror of ``w2v2_pack_job`` (include/w2v2.h)."""
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32), ("src_ld", C.c_int32),
                ("dst_ld", C.c_int32), ("transpose", C.c_int32), ("dst_f32", C.c_int32), ("scale", C.c_float),
                ("reserved", C.c_int32)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
_U64, _U32 = C.c_uint64, C.c_uint32

# name -> argtypes (restype is int unless listed in _RESTYPES).  Kept in one table so the
# "library exports every declared symbol" test can walk it.
SIGNATURES = {
    "w2v2_version": [],
    "w2v2_last_error_string": [],
    "w2v2_gemm_bf16": [C.POINTER(GemmArgs), _P],
    "w2v2_wave_stats": [_P, _I, _I, _P, _P],
    "w2v2_conv0_fold": [_P, _P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P],
    "w2v2_conv0_im2col": [_P, _I, _I, _P, _P, _P],
    "w2v2_conv0_gn_gelu": [_P, _I, _I, _I, _P, _P, _P, _P, _P, _I, _I, _P],
    "w2v2_conv0_ln_gelu": [_P, _I, _I, _I, _P, _P, _P, _P, _F, _P, _P, _I, _I, _P],
    "w2v2_conv0": [_P, _I, _I, _I, _P, _I, _P, _I, _I, _P, _P, _P, _P],
    "w2v2_ln_rows": [_P, _P, _P, _F, _L, _I, _I, _P, _P, _P, _P],
    "w2v2_ln_rows_stats": [_P, _P, _P, _F, _L, _I, _I, _P, _P, _P, _P, _P],
    "w2v2_ln_rows_ex": [_P, _P, _P, _F, _L, _I, _I, _P, _P, _P, _P, _I, _P],
    "w2v2_row_stats_finalize": [_P, _I, _L, _I, _F, _P, _P],
    "w2v2_normalize_utterances": [_P, _P, _I, _I, _F, _P, _P],
    "w2v2_split_bf16": [_P, _L, _P, _P, _P],
    "w2v2_attn_fwd": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P],
    "w2v2_attn_fwd_ex": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _P],
    "w2v2_posconv": [C.POINTER(PosconvArgs), _P],
    "w2v2_ctc_loss": [_P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P],
    "w2v2_ctc_workspace_bytes": [_I, _I, _I],
    "w2v2_frame_argmax": [_P, _L, _I, _P, _P],
    "w2v2_lm_head_wgrad": [_P, _P, _L, _I, _I, _P, _P, _P],
    "w2v2_adam": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _P],
    "w2v2_pack_weights": [_P, _P, _I, _P],
    "w2v2_ln_bwd": [_P, _P, _P, _F, _L, _I, _P, _P, _P, _P, _P, _P],
    "w2v2_gelu_rows": [_P, _L, _I, _P, _P, _F, _U64, _U32, _P],
    "w2v2_dact_colsum": [_P, _P, _L, _I, _P, _P, _F, _U64, _U32, _P],
    "w2v2_transpose_bf16": [_P, _L, _I, _P, _L, _P],
    "w2v2_lm_head_dgrad": [_P, _P, _L, _I, _I, _P, _P],
    "w2v2_attn_bwd_workspace_bytes": [_I, _I, _I],
    "w2v2_attn_bwd": [_P, _P, _P, _I, _I, _I, _I, _P, _F, _P, _P, _F, _U64, _U32, _P],
    "w2v2_dropout_rows": [_P, _P, _L, _F, _U64, _U32, _P, _P, _P],
    "w2v2_dropout_mask": [_L, _F, _U64, _U32, _P, _P],
    "w2v2_attn_dropout_mask": [_I, _I, _F, _U64, _U32, _P, _P],
    "w2v2_attn_fwd_train": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _F, _U64, _U32, _P],
    "w2v2_posconv_wgrad": [_P, _P, _I, _I, _I, _I, _I, _P, _P],
}
_RESTYPES = {"w2v2_last_error_string": C.c_char_p, "w2v2_ctc_workspace_bytes": C.c_int64,
             "w2v2_attn_bwd_workspace_bytes": C.c_int64}

_lib = None


def build(verbose=False):
    """Compile the CUDA sources in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libw2v2_sm100.so failed (see output above)")
    return LIB_PATH


def load():
    """dlopen the kernel library; raise loudly when it is absent (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` or `make -C {CSRC_DIR}`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().w2v2_last_error_string()
        raise RuntimeError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")
