"""The fp16-family precision modes (include/w2v2.h W2V2_MODE_*): 17 = fp16, 19 = split-fp16, 25 = fp16 main product + both
cross terms as e4m3 MMAs ("fp16f8").  Kernel by kernel through the C ABI against fp64 products of the fp32 inputs, then the whole
model against the oracle at the reference's tolerances.  The planes are decoded here exactly as include/w2v2.h defines them."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import w2v2_oracle as O                                     # noqa: E402 (checker only)
from wav2vec2 import RobustWav2Vec2Config, Wav2Vec2Config, Wav2Vec2ForCTC, ops   # noqa: E402
from wav2vec2.modeling import _split                                    # noqa: E402
from wav2vec2.ops import Pair                                           # noqa: E402

DEV = "cuda"
ACT, WGT = 16.0, 2048.0


def act_planes(x, mode):
    """Activation planes of ``x`` for ``mode`` built on the host exactly like the kernels' epilogues build them."""
    s = torch.clamp(x.float() * ACT, -65504, 65504)
    hi = s.to(torch.float16)
    if mode == 17:
        return Pair(hi.contiguous(), None)
    res = s - hi.float()
    if mode == 19:
        return Pair(hi.contiguous(), res.to(torch.float16).contiguous())
    rows, K = x.reshape(-1, x.shape[-1]).shape
    l8 = torch.clamp(res * 64, -448, 448).to(torch.float8_e4m3fn).view(torch.uint8).reshape(rows, K // 64, 1, 64)
    h8 = torch.clamp(hi.float() / 64, -448, 448).to(torch.float8_e4m3fn).view(torch.uint8).reshape(rows, K // 64, 1, 64)
    return Pair(hi.contiguous(), torch.cat([l8, h8], 2).reshape(tuple(x.shape[:-1]) + (2 * K,)).contiguous())


def decode(p: Pair, kind, cols):
    """Planes -> fp32 value (the value the NEXT kernel effectively consumes)."""
    hi = p.hi.float()
    if kind == "fp16":
        return hi / ACT
    if kind == "fp16x2":
        return (hi + p.lo.float()) / ACT
    rows = hi.reshape(-1, cols).shape[0]
    c8 = p.lo.reshape(rows, cols // 64, 2, 64)
    l8 = c8[:, :, 0].contiguous().view(torch.float8_e4m3fn).float().reshape(rows, cols)
    h8 = c8[:, :, 1].contiguous().view(torch.float8_e4m3fn).float().reshape(rows, cols)
    # the hi bytes are a 4-bit copy of the fp16 plane: check that instead of using it
    assert (h8 * 64 - hi.reshape(rows, cols)).abs().max() <= 0.07 * hi.abs().max() + 1e-3
    return ((hi.reshape(rows, cols) + l8 / 64) / ACT).reshape(hi.shape)


@pytest.mark.parametrize("mode,tol", [(17, 2e-3), (19, 2e-6), (25, 6e-5)])
@pytest.mark.parametrize("M,K,N,block_n", [(300, 192, 256, 0), (128 * 9 + 7, 768, 768, 0), (257, 512, 128, 128), (200, 768, 32, 32)])
def test_gemm_fp16_family(mode, tol, M, K, N, block_n):
    """D = A W^T + bias with the operands in the mode's planes; error relative to |A||W| row norms (fp64 reference of the fp32 data)."""
    torch.manual_seed(1)
    a32 = torch.randn(M, K, device=DEV) * torch.exp(torch.randn(M, 1, device=DEV))     # rows of very different magnitude
    w32 = torch.randn(N, K, device=DEV) / math.sqrt(K)
    bias = torch.randn(N, device=DEV)
    wp = _split(torch.cat([w32, torch.zeros((-N) % max(block_n, 32), K, device=DEV)]) if N % 32 else w32, mode)
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm(act_planes(a32, mode), wp, K=K, N=N, rows_per_batch=M, bias=bias, out_f32=out, passes=mode, block_n=block_n)
    torch.cuda.synchronize()
    ref = (a32.double() @ w32.double().t() + bias.double())
    scale = a32.double().norm(dim=1, keepdim=True) * w32.double().norm(dim=1)[None, :]
    err = ((out.double() - ref).abs() / scale).max().item()
    print(f"gemm mode {mode} M={M} K={K} N={N}: max err / (|a||w|) = {err:.3e}")
    assert err < tol


@pytest.mark.parametrize("mode,kind", [(17, "fp16"), (19, "fp16x2"), (25, "fp16f8")])
def test_gemm_output_planes(mode, kind):
    """GELU epilogue + the three fp16-family output layouts (W2V2_OUT_FP16 with / without the residual plane, W2V2_OUT_FP16F8)."""
    torch.manual_seed(2)
    M, K, N = 128 * 5 + 33, 256, 512
    a32, w32, bias = torch.randn(M, K, device=DEV), torch.randn(N, K, device=DEV) / math.sqrt(K), torch.randn(N, device=DEV)
    hi = torch.zeros(M, N, dtype=torch.float16, device=DEV)
    lo = None if kind == "fp16" else (torch.zeros(M, N, dtype=torch.float16, device=DEV) if kind == "fp16x2"
                                      else torch.zeros(M, 2 * N, dtype=torch.uint8, device=DEV))
    f32 = torch.empty(M, N, device=DEV)
    ops.gemm(act_planes(a32, mode), _split(w32, mode), K=K, N=N, rows_per_batch=M, bias=bias, gelu=True, out_f32=f32,
             out_hi=hi, out_lo=lo, passes=mode, out_format={"fp16": 1, "fp16x2": 1, "fp16f8": 2}[kind])
    torch.cuda.synchronize()
    got = decode(Pair(hi, lo), kind, N)
    # relative precision of the planes, plus the absolute floor of the e4m3 residual byte (subnormal step 2^-9 / 2^6 / 2^4)
    tol = {"fp16": 2.0 ** -11, "fp16x2": 2.0 ** -21, "fp16f8": 2.0 ** -15}[kind]
    err = ((got - f32).abs() / (f32.abs() + 2.0 ** -5)).max().item()
    print(f"out planes {kind}: max relative error of the decoded planes vs the fp32 output {err:.3e}")
    assert err < 1.2 * tol
    ref = torch.nn.functional.gelu(a32.double() @ w32.double().t() + bias.double())
    assert (f32.double() - ref).abs().max().item() < {17: 2e-2, 19: 1e-4, 25: 5e-4}[mode]


@pytest.mark.parametrize("d", [512, 768, 1024])
def test_ln_rows_fp16f8_planes(d):
    torch.manual_seed(3)
    rows = 777
    x = torch.randn(rows, d, device=DEV) * 3 + 1
    g, b = torch.randn(d, device=DEV), torch.randn(d, device=DEV)
    f32 = torch.empty(rows, d, device=DEV)
    hi = torch.zeros(rows, d, dtype=torch.float16, device=DEV)
    c8 = torch.zeros(rows, 2 * d, dtype=torch.uint8, device=DEV)
    ops.ln_rows(x, g, b, 1e-5, rows, d, out_f32=f32, out_hi=hi, out_lo=c8, out_format=2)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.double(), (d,), g.double(), b.double(), 1e-5)
    assert (f32.double() - ref).abs().max().item() < 1e-4
    got = decode(Pair(hi, c8), "fp16f8", d)
    assert ((got - f32).abs() / (f32.abs() + 2.0 ** -5)).max().item() < 1.2 * 2.0 ** -15
    want = act_planes(f32, 25)
    assert torch.equal(hi, want.hi) and torch.equal(c8, want.lo)        # bit-identical to the host construction


@pytest.mark.parametrize("mode,ofmt,tol", [(17, 1, 3e-3), (19, 1, 2e-5), (17, 2, 3e-3)])
@pytest.mark.parametrize("T", [49, 300, 768])
def test_attention_fp16_modes(mode, ofmt, tol, T):
    torch.manual_seed(4)
    B, H, dh = 2, 3, 64
    d = H * dh
    qkv32 = torch.randn(B, T, 3 * d, device=DEV)
    qkv32[..., :d] *= dh ** -0.5
    kv_len = torch.tensor([T, max(1, T - 17)], dtype=torch.int32, device=DEV)
    out = Pair(torch.zeros(B, T, d, dtype=torch.float16, device=DEV),
               torch.zeros(B, T, d, dtype=torch.float16, device=DEV) if (mode == 19 and ofmt == 1)
               else (torch.zeros(B, T, 2 * d, dtype=torch.uint8, device=DEV) if ofmt == 2 else None))
    ops.attn_fwd(act_planes(qkv32, mode), B, T, H, dh, kv_len, out, mode, out_format=ofmt)
    torch.cuda.synchronize()
    q, k, v = (t.reshape(B, T, H, dh).permute(0, 2, 1, 3).double() for t in qkv32.split(d, dim=-1))
    s = q @ k.transpose(-1, -2)
    keep = torch.arange(T, device=DEV)[None, None, None, :] < kv_len[:, None, None, None]
    ref = (torch.softmax(s.masked_fill(~keep, -1e30), -1) @ v).permute(0, 2, 1, 3).reshape(B, T, d)
    kind = "fp16f8" if ofmt == 2 else ("fp16x2" if out.lo is not None else "fp16")
    got = decode(out, kind, d)
    err = (got.double() - ref).abs().max().item()
    print(f"attention mode {mode} out_format {ofmt} T={T}: max err {err:.3e}")
    assert err < tol


def _model(cls_cfg, precision, seed=1, **kw):
    cfg = cls_cfg(**kw)
    params = O.random_params(cfg, seed=seed)
    m = Wav2Vec2ForCTC(cfg, input_shape=(1, 2048), precision=precision)
    m.set_variables(params)
    return cfg, params, m


@pytest.mark.parametrize("arch", ["base", "robust"])
@pytest.mark.parametrize("precision,tol", [("fp16f8", 1e-3), ("fp16", 4e-3)])
def test_model_logits_in_the_fp16_modes(arch, precision, tol):
    """The whole forward in the two fp16-family modes against the oracle: "fp16f8" holds the north star's 1e-3, plain "fp16" the
    reference's own logits tolerance of 4e-3 (tests/test_wav2vec2.py:155-157)."""
    cls = Wav2Vec2Config if arch == "base" else RobustWav2Vec2Config
    cfg, params, m = _model(cls, precision, num_layers=3)
    x = torch.randn(2, 20000, generator=torch.Generator().manual_seed(0))
    am = None
    if arch == "robust":
        am = torch.ones(2, 20000, dtype=torch.int32)
        am[0, -1000:] = 0
        x = x * am
    got = m(x.cuda(), attention_mask=None if am is None else am.cuda()).cpu()
    ref = O.wav2vec2_for_ctc(x, params, cfg, attention_mask=am)
    err = (got - ref).abs().max().item()
    print(f"{arch}/{precision}: logits max-abs err {err:.3e} (max |logit| {ref.abs().max():.2f})")
    assert err < tol


def test_fp16f8_sample_wav_full_depth_and_graph():
    """12 layers on the reference's sample.wav in "fp16f8", eager and as a CUDA graph; greedy CTC path identical to the oracle."""
    import os
    cfg, params, m = _model(Wav2Vec2Config, "fp16f8", seed=3)
    wav = O.read_wav_s16(os.path.join(os.path.dirname(__file__), "golden", "sample.wav"))
    x = torch.from_numpy(O.normalize_utterance(wav[None, :]))[None, :]
    got = m(x.cuda()).cpu()
    ref = O.wav2vec2_for_ctc(x, params, cfg)
    err = (got - ref).abs().max().item()
    print(f"sample.wav 12 layers fp16f8: logits max-abs err {err:.3e}")
    assert err < 1e-3 and torch.equal(got.argmax(-1), ref.argmax(-1))
    m.enable_cuda_graph(True)
    assert torch.equal(m(x.cuda()).cpu(), got)
