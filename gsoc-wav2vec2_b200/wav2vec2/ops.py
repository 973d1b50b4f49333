"""Thin torch-tensor wrappers over the C ABI (``include/w2v2.h``).

Every function takes CUDA tensors, passes raw device pointers + sizes to ``libw2v2_sm100.so`` on the
current torch stream and returns the output tensors it was given (or allocates them with
``torch.empty``).  No arithmetic happens here; there is no eager / CPU fallback.
"""
import ctypes as C
from typing import Optional

import torch

from . import _lib

BF16 = torch.bfloat16
TWO_PLANE_MODES = (_lib.MODE_BF16X3, _lib.MODE_FP16X3, _lib.MODE_FP16F8)   # precision modes whose operands have a second plane
LAUNCHES = 0   # kernels launched through this module (bench.py reports it as gpu_launches)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("w2v2 kernels need CUDA tensors (there is no CPU fallback)")


class Pair:
    """An activation / weight as operand planes: hi (+ optional lo).  bf16 modes: value ~= hi + lo (bf16 each); fp16 modes: hi (+ lo)
    = fp16 planes of value * 2^4 (weights: * 2^11); fp16f8: lo is the uint8 e4m3 pair plane [rows][2 K] (include/w2v2.h)."""
    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo=None):
        self.hi, self.lo = hi, lo

    def float(self):
        return self.hi.float() if self.lo is None else self.hi.float() + self.lo.float()


def split_bf16(x: torch.Tensor, with_lo: bool) -> Pair:
    x = x.contiguous().float()
    _need_cuda(x)
    hi = torch.empty(x.shape, dtype=BF16, device=x.device)
    lo = torch.empty(x.shape, dtype=BF16, device=x.device) if with_lo else None
    _count(); _lib.check(_lib.load().w2v2_split_bf16(_ptr(x), x.numel(), _ptr(hi), _ptr(lo), _stream()), "w2v2_split_bf16")
    return Pair(hi, lo)


def gemm(a: Pair, w: Pair, *, K: int, N: int, rows_per_batch: int, batch: int = 1,
         a_row_len: Optional[int] = None, a_rows: Optional[int] = None, a_row_stride: Optional[int] = None,
         a_batch_stride: Optional[int] = None, bias=None, scale=None, bias_batch_stride=0, residual=None,
         row_valid=None, gelu=False,
         out_f32=None, out_hi=None, out_lo=None, passes=1, kb_split=0, block_n=0, max_ctas=0, cluster=0, debug=0,
         res_ln=None, mn_major=False, w_row_stride=0, gelu_approx=False, row_replace=None, drop=None, out_format=0,
         ln_fold=None, row_stats_out=None, ln_eps=1e-5, row_stats_final=None):
    """D = epilogue(A . W^T).  ``w.hi`` is [w_rows, K] bf16 row-major.
    ``res_ln = (stats, gamma, beta)``: the residual term is LayerNorm(residual) recomputed from ``ln_rows(..., stats=)``.
    ``mn_major``: D[m][n] = sum_r X[r][m] Y[r][n] with a = X [a_rows, ld a_row_stride], w = Y [a_rows, ld w_row_stride].
    ``res_ln[0]`` may also be the partial-sum tensor [parts, rows, 2] that ``row_stats_out`` of an earlier GEMM wrote.
    ``ln_fold = (stats [parts, rows, 2], colsum [N])``: LayerNorm folded into this GEMM (``a`` = un-normalised input, ``w`` packed
    as gamma o W, ``bias`` = beta W + b).  ``row_stats_out`` [N/64, rows, 2]: (sum, sum of squares) of the fp32 result.
    ``gelu_approx``: the GELU is tf.nn.gelu(approximate=True) (config.is_gelu_approx).  ``row_replace = (mask uint8 [rows],
    value fp32 [N])``: SpecAugment row replacement; ``drop = (rate, seed, site)``: dropout before the residual add."""
    _need_cuda(a.hi, w.hi, bias, scale, residual, row_valid, out_f32, out_hi, out_lo)
    args = _lib.GemmArgs()
    two = passes in TWO_PLANE_MODES
    args.a_hi, args.a_lo = _ptr(a.hi), _ptr(a.lo) if two else None
    args.a_row_len = K if a_row_len is None else a_row_len
    args.a_rows = rows_per_batch if a_rows is None else a_rows
    args.a_row_stride = K if a_row_stride is None else a_row_stride
    args.a_batch_stride = (args.a_rows * args.a_row_stride) if a_batch_stride is None else a_batch_stride
    args.w_hi, args.w_lo = _ptr(w.hi), _ptr(w.lo) if two else None
    args.w_rows = w.hi.shape[0]
    args.K, args.N, args.rows_per_batch, args.batch = K, N, rows_per_batch, batch
    args.passes, args.kb_split, args.block_n, args.max_ctas, args.cluster = passes, kb_split, block_n, max_ctas, cluster
    args.flags = (_lib.GEMM_GELU if gelu else 0) | (_lib.GEMM_MN_MAJOR if mn_major else 0) | (debug << 8) | \
        (_lib.GEMM_GELU_TANH if (gelu and gelu_approx) else 0)
    if row_replace is not None:
        _need_cuda(*row_replace)
        args.row_replace_mask, args.row_replace_value = _ptr(row_replace[0]), _ptr(row_replace[1])
    if drop is not None and drop[0] > 0.0:
        args.drop_p, args.drop_seed, args.drop_site = float(drop[0]), int(drop[1]), int(drop[2])
    args.w_row_stride = w_row_stride
    args.bias, args.residual, args.row_valid = _ptr(bias), _ptr(residual), _ptr(row_valid)
    args.scale, args.bias_batch_stride = _ptr(scale), bias_batch_stride
    args.out_f32, args.out_hi, args.out_lo = _ptr(out_f32), _ptr(out_hi), _ptr(out_lo)
    args.out_format = out_format
    args.ln_eps = float(ln_eps)
    if res_ln is not None:
        _need_cuda(*res_ln)
        args.res_ln_stats, args.res_ln_gamma, args.res_ln_beta = (_ptr(t) for t in res_ln)
        args.res_ln_parts = res_ln[0].shape[0] if res_ln[0].dim() == 3 else 0
    if ln_fold is not None:
        _need_cuda(*ln_fold)
        args.ln_fold_stats, args.ln_fold_parts = _ptr(ln_fold[0]), (ln_fold[0].shape[0] if ln_fold[0].dim() == 3 else 0)
        args.scale = _ptr(ln_fold[1])
    if row_stats_out is not None:
        _need_cuda(row_stats_out)
        args.row_stats_out = _ptr(row_stats_out)
    if row_stats_final is not None:       # (stats [rows, 2], counters [ceil(rows / 32)] uint32 zero-initialised): finalised by this call
        _need_cuda(*row_stats_final)
        args.row_stats_final, args.row_stats_counter = _ptr(row_stats_final[0]), _ptr(row_stats_final[1])
    _count(); _lib.check(_lib.load().w2v2_gemm_bf16(C.byref(args), _stream()), "w2v2_gemm_bf16")


def wave_stats(wave: torch.Tensor, stats: torch.Tensor):
    _need_cuda(wave, stats)
    B, L = wave.shape
    _count(); _lib.check(_lib.load().w2v2_wave_stats(_ptr(wave), B, L, _ptr(stats), _stream()), "w2v2_wave_stats")


def conv0_fold(kernel, gamma, beta, stats, B, L, folded_w, folded_b, eps=1e-5, scale=None):
    C_ = kernel.shape[-1]
    _count(); _lib.check(_lib.load().w2v2_conv0_fold(_ptr(kernel), _ptr(gamma), _ptr(beta), _ptr(stats), B, L, C_, eps,
                                           _ptr(folded_w), _ptr(folded_b), _ptr(scale), _stream()), "w2v2_conv0_fold")


def conv0_im2col(wave, a: Pair):
    _need_cuda(wave, a.hi, a.lo)
    B, L = wave.shape
    _count(); _lib.check(_lib.load().w2v2_conv0_im2col(_ptr(wave), B, L, _ptr(a.hi), _ptr(a.lo), _stream()), "w2v2_conv0_im2col")


def conv0_gn_gelu(wave, kernel, scale, shift, out: Pair, passes=1, gelu_approx=False):
    """Fused extractor layer 0: conv (tensor cores, no im2col) + folded GroupNorm + GELU, written once."""
    _need_cuda(wave, kernel, scale, shift, out.hi, out.lo)
    B, L = wave.shape
    lo = out.lo if passes in TWO_PLANE_MODES else None
    _count(); _lib.check(_lib.load().w2v2_conv0_gn_gelu(_ptr(wave), B, L, kernel.shape[-1], _ptr(kernel), _ptr(scale),
                                              _ptr(shift), _ptr(out.hi), _ptr(lo), passes, 1 if gelu_approx else 0,
                                              _stream()),
                         "w2v2_conv0_gn_gelu")


def conv0_ln_gelu(wave, kernel, conv_bias, gamma, beta, eps, out: Pair, passes=1, gelu_approx=False):
    """Layer 0 of the layer-norm extractor in one kernel: conv (+ bias) + LayerNorm over the channels + GELU -> operand planes."""
    _need_cuda(wave, kernel, conv_bias, gamma, beta, out.hi)
    B, L = wave.shape
    two = passes in (3, 25)
    _count(); _lib.check(_lib.load().w2v2_conv0_ln_gelu(_ptr(wave), B, L, kernel.shape[-1], _ptr(kernel), _ptr(conv_bias),
                                              _ptr(gamma), _ptr(beta), float(eps), _ptr(out.hi),
                                              _ptr(out.lo) if two else None, passes, 1 if gelu_approx else 0, _stream()),
                         "w2v2_conv0_ln_gelu")


def conv0(wave, weights, w_batch_stride, bias, b_batch_stride, gelu, out_f32=None, out_hi=None, out_lo=None,
          channels=512):
    _need_cuda(wave, weights, bias, out_f32, out_hi, out_lo)
    B, L = wave.shape
    _count(); _lib.check(_lib.load().w2v2_conv0(_ptr(wave), B, L, channels, _ptr(weights), w_batch_stride, _ptr(bias),
                                      b_batch_stride, 1 if gelu else 0, _ptr(out_f32), _ptr(out_hi), _ptr(out_lo),
                                      _stream()), "w2v2_conv0")


def normalize_utterances(wave, lengths=None, eps=1e-5, out=None):
    """Device version of Wav2Vec2Processor._normalize: wave [B, L] fp32 (CUDA), lengths [B] int32 or None."""
    _need_cuda(wave, lengths, out)
    wave = wave.contiguous().float()
    B, L = wave.shape
    out = torch.empty_like(wave) if out is None else out
    _count(); _lib.check(_lib.load().w2v2_normalize_utterances(_ptr(wave), _ptr(lengths), B, L, float(eps), _ptr(out), _stream()),
                         "w2v2_normalize_utterances")
    return out


def ln_rows(x, gamma, beta, eps, rows, d, gelu=False, out_f32=None, out_hi=None, out_lo=None, stats=None, out_format=0):
    """``stats`` [rows, 2] fp32 (optional) receives (mean, rstd) per row for ``gemm(..., res_ln=(stats, gamma, beta))``."""
    _need_cuda(x, gamma, beta, out_f32, out_hi, out_lo, stats)
    _count(); _lib.check(_lib.load().w2v2_ln_rows_ex(_ptr(x), _ptr(gamma), _ptr(beta), float(eps), rows, d, int(gelu),
                                           _ptr(out_f32), _ptr(out_hi), _ptr(out_lo), _ptr(stats), out_format, _stream()),
                         "w2v2_ln_rows")


def row_stats_finalize(parts, dim, eps, stats):
    """parts [P, rows, 2] partial (sum, sum of squares) -> stats [rows, 2] = (mean, rstd) of a LayerNorm over ``dim`` columns."""
    _need_cuda(parts, stats)
    P, rows, _ = parts.shape
    _count(); _lib.check(_lib.load().w2v2_row_stats_finalize(_ptr(parts), P, rows, int(dim), float(eps), _ptr(stats), _stream()),
                         "w2v2_row_stats_finalize")
    return stats


def attn_fwd(qkv: Pair, B, T, H, dh, kv_len, out: Pair, passes=1, out_format=None):
    """``passes``: 1 / 3 (bf16 planes) or 17 / 19 (fp16 planes of q, k, v * 2^4).  ``out_format`` (default: bf16 planes for the
    bf16 modes, fp16 for the fp16 modes; 2 = fp16 + e4m3 pair plane for a mode-25 output projection); ``out.lo`` is written when
    the mode has three passes (second 16-bit plane) or the format requires it (2)."""
    _need_cuda(qkv.hi, out.hi, kv_len)
    if out_format is None:
        out_format = 1 if passes & 16 else 0
    three = (passes & 3) == 3
    out_lo = out.lo if (three or out_format == 2) else None
    _count(); _lib.check(_lib.load().w2v2_attn_fwd_ex(_ptr(qkv.hi), _ptr(qkv.lo) if three else None, B, T, H, dh,
                                            _ptr(kv_len), _ptr(out.hi), _ptr(out_lo), passes, out_format,
                                            _stream()), "w2v2_attn_fwd")


def posconv(x: Pair, w: Pair, bias, resid, out_f32, B, T, d, groups, ktaps, passes=1, gelu_approx=False):
    _need_cuda(x.hi, w.hi, bias, resid, out_f32)
    args = _lib.PosconvArgs()
    args.x_hi, args.x_lo = _ptr(x.hi), _ptr(x.lo) if (passes & 3) == 3 else None
    args.w_hi, args.w_lo = _ptr(w.hi), _ptr(w.lo) if (passes & 3) == 3 else None
    args.bias, args.resid, args.out_f32 = _ptr(bias), _ptr(resid), _ptr(out_f32)
    args.batch, args.frames, args.hidden, args.groups, args.ktaps, args.passes = B, T, d, groups, ktaps, passes
    args.gelu_approx = 1 if gelu_approx else 0
    _count(); _lib.check(_lib.load().w2v2_posconv(C.byref(args), _stream()), "w2v2_posconv")


def ctc_loss(logits, labels, blank, scale, want_grad=True):
    """Returns (loss_per_sample [B] fp32, grad [B,T,V] fp32 or None)."""
    _need_cuda(logits, labels)
    logits = logits.contiguous().float()
    labels = labels.contiguous().to(torch.int32)
    B, T, V = logits.shape
    Lmax = labels.shape[1]
    lib = _lib.load()
    ws = torch.empty(lib.w2v2_ctc_workspace_bytes(B, T, Lmax) // 4, dtype=torch.float32, device=logits.device)
    loss = torch.empty(B, dtype=torch.float32, device=logits.device)
    grad = torch.empty_like(logits) if want_grad else None
    _count(); _lib.check(lib.w2v2_ctc_loss(_ptr(logits), _ptr(labels), B, T, V, Lmax, int(blank), float(scale), _ptr(ws), None,
                                 _ptr(loss), _ptr(grad), _stream()), "w2v2_ctc_loss")
    return loss, grad


def frame_argmax(logits):
    _need_cuda(logits)
    logits = logits.contiguous().float()
    V = logits.shape[-1]
    rows = logits.numel() // V
    ids = torch.empty(logits.shape[:-1], dtype=torch.int32, device=logits.device)
    _count(); _lib.check(_lib.load().w2v2_frame_argmax(_ptr(logits), rows, V, _ptr(ids), _stream()), "w2v2_frame_argmax")
    return ids


def lm_head_wgrad(hidden, grad_logits, grad_kernel, grad_bias):
    _need_cuda(hidden, grad_logits, grad_kernel, grad_bias)
    rows, d = hidden.shape
    V = grad_logits.shape[-1]
    _count(); _lib.check(_lib.load().w2v2_lm_head_wgrad(_ptr(hidden), _ptr(grad_logits), rows, d, V, _ptr(grad_kernel),
                                              _ptr(grad_bias), _stream()), "w2v2_lm_head_wgrad")


def adam(w, g, m, v, lr_t, beta1, beta2, eps):
    _need_cuda(w, g, m, v)
    _count(); _lib.check(_lib.load().w2v2_adam(_ptr(w), _ptr(g), _ptr(m), _ptr(v), w.numel(), float(lr_t), float(beta1),
                                     float(beta2), float(eps), _stream()), "w2v2_adam")


# ------------------------------------------------------------------------------------------ stage-2 backward kernels
def ln_bwd(x, gamma, dy, eps, rows, d, dx_f32=None, dx_hi=None, dgamma=None, dbeta=None, colsum=None):
    """LayerNorm backward; dgamma / dbeta / colsum are accumulated into (caller zeroes them once per step)."""
    _need_cuda(x, gamma, dy, dx_f32, dx_hi, dgamma, dbeta, colsum)
    _count(); _lib.check(_lib.load().w2v2_ln_bwd(_ptr(x), _ptr(gamma), _ptr(dy), float(eps), rows, d, _ptr(dx_f32), _ptr(dx_hi),
                                       _ptr(dgamma), _ptr(dbeta), _ptr(colsum), _stream()), "w2v2_ln_bwd")


NO_DROP = (0.0, 0, 0)   # (rate, seed, site): dropout off


def gelu_rows(pre, out_hi, fast, out_lo=None, drop=NO_DROP):
    _need_cuda(pre, out_hi, out_lo)
    _count(); _lib.check(_lib.load().w2v2_gelu_rows(_ptr(pre), pre.numel(), 1 if fast else 0, _ptr(out_hi), _ptr(out_lo),
                                          float(drop[0]), int(drop[1]), int(drop[2]), _stream()), "w2v2_gelu_rows")


def dact_colsum(dy_hi, pre, rows, cols, out_hi=None, colsum=None, drop=NO_DROP):
    _need_cuda(dy_hi, pre, out_hi, colsum)
    _count(); _lib.check(_lib.load().w2v2_dact_colsum(_ptr(dy_hi), _ptr(pre), rows, cols, _ptr(out_hi), _ptr(colsum),
                                            float(drop[0]), int(drop[1]), int(drop[2]), _stream()), "w2v2_dact_colsum")


def dropout_rows(x, drop, resid=None, out_f32=None, out_hi=None):
    """out = (resid or 0) + dropout(x); fp32 (in place allowed) and / or a bf16 copy."""
    _need_cuda(x, resid, out_f32, out_hi)
    _count(); _lib.check(_lib.load().w2v2_dropout_rows(_ptr(x), _ptr(resid), x.numel(), float(drop[0]), int(drop[1]), int(drop[2]),
                                             _ptr(out_f32), _ptr(out_hi), _stream()), "w2v2_dropout_rows")


def dropout_mask(n, drop, device):
    out = torch.empty(n, dtype=torch.uint8, device=device)
    _lib.check(_lib.load().w2v2_dropout_mask(n, float(drop[0]), int(drop[1]), int(drop[2]), _ptr(out), _stream()), "w2v2_dropout_mask")
    return out


def attn_dropout_mask(BH, T, drop, device):
    out = torch.empty((BH, T, T), dtype=torch.uint8, device=device)
    _lib.check(_lib.load().w2v2_attn_dropout_mask(BH, T, float(drop[0]), int(drop[1]), int(drop[2]), _ptr(out), _stream()),
               "w2v2_attn_dropout_mask")
    return out


def attn_fwd_train(qkv: Pair, B, T, H, dh, kv_len, out: Pair, passes, drop):
    _need_cuda(qkv.hi, out.hi, kv_len)
    _count(); _lib.check(_lib.load().w2v2_attn_fwd_train(_ptr(qkv.hi), _ptr(qkv.lo) if passes == 3 else None, B, T, H, dh,
                                               _ptr(kv_len), _ptr(out.hi), _ptr(out.lo) if passes == 3 else None, passes,
                                               float(drop[0]), int(drop[1]), int(drop[2]), _stream()), "w2v2_attn_fwd_train")


def transpose_bf16(x, rows, cols, out, out_ld):
    _need_cuda(x, out)
    _count(); _lib.check(_lib.load().w2v2_transpose_bf16(_ptr(x), rows, cols, _ptr(out), out_ld, _stream()), "w2v2_transpose_bf16")


def lm_head_dgrad(grad_logits, kernel, out):
    _need_cuda(grad_logits, kernel, out)
    d, V = kernel.shape
    rows = grad_logits.numel() // V
    _count(); _lib.check(_lib.load().w2v2_lm_head_dgrad(_ptr(grad_logits), _ptr(kernel), rows, d, V, _ptr(out), _stream()),
                         "w2v2_lm_head_dgrad")


def attn_bwd(qkv_hi, ctx_hi, dctx_hi, B, T, H, dh, kv_len, q_scale, dqkv_hi, workspace=None, drop=NO_DROP):
    _need_cuda(qkv_hi, ctx_hi, dctx_hi, dqkv_hi, kv_len)
    lib = _lib.load()
    need = lib.w2v2_attn_bwd_workspace_bytes(B, T, H)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(need // 4, dtype=torch.float32, device=qkv_hi.device)
    _count(2); _lib.check(lib.w2v2_attn_bwd(_ptr(qkv_hi), _ptr(ctx_hi), _ptr(dctx_hi), B, T, H, dh, _ptr(kv_len), float(q_scale),
                                  _ptr(workspace), _ptr(dqkv_hi), float(drop[0]), int(drop[1]), int(drop[2]), _stream()),
                          "w2v2_attn_bwd")
    return workspace


def posconv_train(x: Pair, w: Pair, bias, resid, out_f32, B, T, d, groups, ktaps, passes=1, pre_out=None, shift=0, linear=False):
    """w2v2_posconv with the training options: pre-activation output / transposed-conv (input gradient) mode."""
    _need_cuda(x.hi, w.hi, bias, resid, out_f32, pre_out)
    args = _lib.PosconvArgs()
    args.x_hi, args.x_lo = _ptr(x.hi), _ptr(x.lo) if (passes & 3) == 3 else None
    args.w_hi, args.w_lo = _ptr(w.hi), _ptr(w.lo) if (passes & 3) == 3 else None
    args.bias, args.resid, args.out_f32 = _ptr(bias), _ptr(resid), _ptr(out_f32)
    args.batch, args.frames, args.hidden, args.groups, args.ktaps, args.passes = B, T, d, groups, ktaps, passes
    args.pre_out, args.shift, args.linear = _ptr(pre_out), shift, 1 if linear else 0
    _count(); _lib.check(_lib.load().w2v2_posconv(C.byref(args), _stream()), "w2v2_posconv")


def posconv_wgrad(x_hi, dpre_hi, B, T, d, groups, ktaps, grad_kernel):
    _need_cuda(x_hi, dpre_hi, grad_kernel)
    _count(); _lib.check(_lib.load().w2v2_posconv_wgrad(_ptr(x_hi), _ptr(dpre_hi), B, T, d, groups, ktaps, _ptr(grad_kernel),
                                              _stream()), "w2v2_posconv_wgrad")


def pack_weights(jobs_dev, tiles_dev, num_tiles):
    """ONE launch re-packing every updated weight (job / tile tables built once by training.Stage2Trainer)."""
    _need_cuda(jobs_dev, tiles_dev)
    _count(); _lib.check(_lib.load().w2v2_pack_weights(_ptr(jobs_dev), _ptr(tiles_dev), int(num_tiles), _stream()), "w2v2_pack_weights")
