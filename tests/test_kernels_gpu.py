"""GPU parity tests, kernel by kernel, through the C ABI (ctypes) against torch fp32 / the oracle."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import w2v2_oracle as O          # noqa: E402  (checker only)
from wav2vec2 import ops                      # noqa: E402
from wav2vec2.config import Wav2Vec2Config    # noqa: E402
from wav2vec2.ops import Pair                 # noqa: E402

DEV = "cuda"


def _pair(x, lo):
    hi = x.to(torch.bfloat16)
    return Pair(hi.contiguous(), (x - hi.float()).to(torch.bfloat16).contiguous() if lo else None)


def _eff(p: Pair, passes):
    return p.hi.float() if passes == 1 else p.hi.float() + p.lo.float()


def _ref_gemm(a: Pair, w: Pair, passes):
    """What the kernel computes: hi*hi (+ lo*hi + hi*lo) with fp32 accumulation (fp64 here)."""
    ah, wh = a.hi.double(), w.hi.double()
    r = ah @ wh.t()
    if passes == 3:
        r = r + a.lo.double() @ wh.t() + ah @ w.lo.double().t()
    return r.float()


@pytest.mark.parametrize("passes", [1, 3])
@pytest.mark.parametrize("N,block_n", [(256, 256), (384, 128), (192, 64), (32, 32), (512, 0)])
def test_gemm_plain(passes, N, block_n):
    torch.manual_seed(0)
    M, K = 300, 192
    a = _pair(torch.randn(M, K, device=DEV), passes == 3)
    w = _pair(torch.randn(N, K, device=DEV) / math.sqrt(K), passes == 3)
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm(a, w, K=K, N=N, rows_per_batch=M, out_f32=out, passes=passes, block_n=block_n)
    torch.cuda.synchronize()
    ref = _ref_gemm(a, w, passes)
    err = (out - ref).abs().max().item()
    print(f"gemm N={N} bn={block_n} passes={passes}: max err {err:.3e}")
    assert err < 2e-4


@pytest.mark.parametrize("cluster", [0, 1, 3])
@pytest.mark.parametrize("M,K,N", [(128 * 37 + 5, 768, 768), (128 * 300, 256, 512), (129, 3072, 256)])
def test_gemm_persistent_many_tiles(cluster, M, K, N):
    """Many tiles per CTA (ring/phase wrap-around, accumulator double buffering), CTA-pair multicast vs single."""
    torch.manual_seed(11)
    a = _pair(torch.randn(M, K, device=DEV), False)
    w = _pair(torch.randn(N, K, device=DEV) / math.sqrt(K), False)
    bias = torch.randn(N, device=DEV)
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm(a, w, K=K, N=N, rows_per_batch=M, bias=bias, out_f32=out, cluster=cluster)
    torch.cuda.synchronize()
    ref = a.hi.float() @ w.hi.float().t() + bias
    err = (out - ref).abs().max().item()
    print(f"gemm M={M} K={K} N={N} cluster={cluster}: max err {err:.3e}")
    assert err < 1e-3


@pytest.mark.parametrize("passes", [1, 3])
def test_gemm_epilogue(passes):
    torch.manual_seed(1)
    B, T, K, N = 3, 200, 128, 256
    lo = passes == 3
    a = _pair(torch.randn(B * T, K, device=DEV), lo)
    w = _pair(torch.randn(N, K, device=DEV) / math.sqrt(K), lo)
    bias = torch.randn(N, device=DEV)
    resid = torch.randn(B * T, N, device=DEV)
    valid = torch.tensor([200, 57, 0], dtype=torch.int32, device=DEV)
    o32 = torch.full((B * T, N), float("nan"), device=DEV)
    ohi = torch.zeros(B * T, N, dtype=torch.bfloat16, device=DEV)
    olo = torch.zeros(B * T, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(a, w, K=K, N=N, rows_per_batch=T, batch=B, bias=bias, residual=resid, row_valid=valid, gelu=True,
             out_f32=o32, out_hi=ohi, out_lo=olo, passes=passes)
    torch.cuda.synchronize()
    ref = O.gelu_erf((_ref_gemm(a, w, passes) + bias).cpu()).to(DEV) + resid
    keep = (torch.arange(T, device=DEV)[None, :] < valid[:, None]).reshape(-1, 1)
    ref = torch.where(keep, ref, torch.zeros((), device=DEV))
    assert (o32 - ref).abs().max().item() < 3e-4
    assert (ohi.float() + olo.float() - o32).abs().max().item() < 2e-4      # hi + lo carries ~16 bits
    assert (ohi.float() - o32).abs().max().item() < 0.05


@pytest.mark.parametrize("passes", [1, 3])
@pytest.mark.parametrize("k,s,Tin", [(3, 2, 301), (2, 2, 290), (3, 2, 64)])
def test_gemm_as_strided_conv(passes, k, s, Tin):
    """Extractor layers 1..6 (feature_extractor.py:55) as an implicit GEMM with overlapping rows."""
    torch.manual_seed(2)
    B, C = 2, 512
    lo = passes == 3
    Tout = 1 + (Tin - k) // s
    x = _pair(torch.randn(B, Tin, C, device=DEV), lo)
    kern = torch.randn(k, C, C, device=DEV) / math.sqrt(k * C)             # TF layout [k, cin, cout]
    w = _pair(kern.permute(2, 0, 1).reshape(C, k * C).contiguous(), lo)
    out = torch.full((B, Tout, C), float("nan"), device=DEV)
    geo = dict(K=k * C, N=C, rows_per_batch=Tout, batch=B, a_row_len=k * C, a_rows=Tout, a_row_stride=s * C,
               a_batch_stride=Tin * C, passes=passes)
    ops.gemm(x, w, gelu=True, out_f32=out, **geo)
    torch.cuda.synchronize()
    w_eff = _eff(w, passes).reshape(C, k, C).permute(1, 2, 0)              # back to [k, cin, cout]
    ref = O.gelu_erf(O.conv1d_valid(_eff(x, passes).cpu(), w_eff.cpu(), None, stride=s)).to(DEV)
    err = (out - ref).abs().max().item()
    print(f"conv-as-gemm k={k} s={s} Tin={Tin} passes={passes}: max err {err:.3e}")
    # 3-pass drops the lo*lo term (~2^-16 relative): tolerance covers it
    assert err < (2e-3 if passes == 1 else 2e-4)
    if k == 3 and Tin % 2 == 0:
        # pair-row addressing (kb_split): frames (2t, 2t+1) from row t, frame 2t+2 from row t+1 - same numbers
        out2 = torch.full((B, Tout, C), float("nan"), device=DEV)
        geo2 = dict(geo, a_row_len=2 * C, a_rows=Tin // 2, a_row_stride=2 * C, kb_split=2 * C // 64)
        ops.gemm(x, w, gelu=True, out_f32=out2, **geo2)
        torch.cuda.synchronize()
        assert (out2 - out).abs().max().item() < 1e-5


def test_ln_rows():
    torch.manual_seed(3)
    for d in (512, 768, 1024):
        x = torch.randn(1000, d, device=DEV) * 3 + 1
        g, b = torch.randn(d, device=DEV), torch.randn(d, device=DEV)
        o32 = torch.empty_like(x)
        ohi = torch.empty(1000, d, dtype=torch.bfloat16, device=DEV)
        olo = torch.empty_like(ohi)
        ops.ln_rows(x, g, b, 1e-5, 1000, d, out_f32=o32, out_hi=ohi, out_lo=olo)
        ref = O.layer_norm(x.cpu().double(), g.cpu().double(), b.cpu().double(), 1e-5).float().to(DEV)
        assert (o32 - ref).abs().max().item() < 2e-5
        assert (ohi.float() + olo.float() - o32).abs().max().item() < 1e-4
        ops.ln_rows(x, g, b, 1e-5, 1000, d, gelu=True, out_f32=o32)
        assert (o32 - O.gelu_erf(ref.cpu()).to(DEV)).abs().max().item() < 3e-5


@pytest.mark.parametrize("L", [46797, 16000, 1210])
def test_conv0_groupnorm_gelu(L):
    """Layer 0: conv + GroupNorm over time + GELU from waveform statistics (tensorflow_addons.py:207-231)."""
    torch.manual_seed(4)
    B, C = 3, 512
    x = torch.randn(B, L)
    x[1] = x[1] * 0.3 + 0.05
    x[2, L // 2:] = 0.0                                         # zero padding enters the statistics (no mask in base)
    kern = torch.randn(10, 1, C) * 0.3
    gamma, beta = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    ref = O.gelu_erf(O.group_norm_per_channel(O.conv1d_valid(x[:, :, None], kern, None, stride=5), gamma, beta, 1e-5))
    T0 = ref.shape[1]
    xd = x.to(DEV)
    stats = torch.empty(B, 65, dtype=torch.float64, device=DEV)
    fw, fb = torch.empty(B, 10, C, device=DEV), torch.empty(B, C, device=DEV)
    ops.wave_stats(xd, stats)
    ops.conv0_fold(kern.reshape(10, C).to(DEV), gamma.to(DEV), beta.to(DEV), stats, B, L, fw, fb)
    hi = torch.empty(B, T0, C, dtype=torch.bfloat16, device=DEV)
    lo = torch.empty_like(hi)
    ops.conv0(xd, fw, 10 * C, fb, C, True, out_hi=hi, out_lo=lo)
    torch.cuda.synchronize()
    got = (hi.float() + lo.float()).cpu()
    err = (got - ref).abs().max().item()
    print(f"conv0 L={L}: max err {err:.3e} (|ref| max {ref.abs().max():.2f})")
    assert err < 2e-4
    assert (hi.float().cpu() - ref).abs().max().item() < 0.04   # bf16 rounding of O(5) values


@pytest.mark.parametrize("passes", [1, 3])
def test_conv0_tensor_core_route(passes):
    """Layer 0 as im2col + GEMM with the GroupNorm folded into a per-(b, c) scale / shift epilogue."""
    torch.manual_seed(9)
    B, C, L = 3, 512, 20011
    lo = passes == 3
    x = torch.randn(B, L)
    x[1] = x[1] * 0.3 + 0.05
    kern = torch.randn(10, 1, C) * 0.3
    gamma, beta = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    ref = O.gelu_erf(O.group_norm_per_channel(O.conv1d_valid(x[:, :, None], kern, None, stride=5), gamma, beta, 1e-5))
    T0 = ref.shape[1]
    xd = x.to(DEV)
    stats = torch.empty(B, 65, dtype=torch.float64, device=DEV)
    fs, fb = torch.empty(B, C, device=DEV), torch.empty(B, C, device=DEV)
    ops.wave_stats(xd, stats)
    ops.conv0_fold(kern.reshape(10, C).to(DEV), gamma.to(DEV), beta.to(DEV), stats, B, L, None, fb, scale=fs)
    a0 = Pair(torch.empty(B, T0, 64, dtype=torch.bfloat16, device=DEV), torch.empty(B, T0, 64, dtype=torch.bfloat16, device=DEV) if lo else None)
    ops.conv0_im2col(xd, a0)
    wg = torch.zeros(C, 64, device=DEV)
    wg[:, :10] = kern.reshape(10, C).t().to(DEV)
    w = _pair(wg, lo)
    hi = torch.empty(B, T0, C, dtype=torch.bfloat16, device=DEV)
    lo_t = torch.empty_like(hi) if lo else None
    ops.gemm(a0, w, K=64, N=C, rows_per_batch=T0, batch=B, a_row_len=64, a_rows=T0, a_row_stride=64, a_batch_stride=T0 * 64,
             bias=fb, scale=fs, bias_batch_stride=C, gelu=True, out_hi=hi, out_lo=lo_t, passes=passes)
    torch.cuda.synchronize()
    got = (hi.float() + (lo_t.float() if lo else 0)).cpu()
    err = (got - ref).abs().max().item()
    print(f"conv0 tensor-core route passes={passes}: max err {err:.3e} (|ref| max {ref.abs().max():.2f})")
    assert err < (0.08 if passes == 1 else 3e-4)


@pytest.mark.parametrize("passes", [1, 3])
@pytest.mark.parametrize("L", [46797, 20011, 1285, 14])
def test_conv0_fused(passes, L):
    """Layer 0 in one kernel (w2v2_conv0_gn_gelu): warp-MMA conv from an smem copy of the waveform + folded GroupNorm
    + GELU.  L = 1285 -> 256 frames (exactly one CTA tile), 14 -> a single frame, the others end in a ragged tile."""
    torch.manual_seed(9)
    B, C = 3, 512
    lo = passes == 3
    x = torch.randn(B, L)
    x[1] = x[1] * 0.3 + 0.05
    kern = torch.randn(10, 1, C) * 0.3
    gamma, beta = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    ref = O.gelu_erf(O.group_norm_per_channel(O.conv1d_valid(x[:, :, None], kern, None, stride=5), gamma, beta, 1e-5))
    T0 = ref.shape[1]
    xd, kd = x.to(DEV), kern.reshape(10, C).to(DEV)
    stats = torch.empty(B, 65, dtype=torch.float64, device=DEV)
    fs, fb = torch.empty(B, C, device=DEV), torch.empty(B, C, device=DEV)
    ops.wave_stats(xd, stats)
    ops.conv0_fold(kd, gamma.to(DEV), beta.to(DEV), stats, B, L, None, fb, scale=fs)
    guard = 64                                                   # canary rows behind the last frame must stay untouched
    hi = torch.full((B * T0 + guard, C), 7.0, dtype=torch.bfloat16, device=DEV)
    lo_t = torch.full((B * T0 + guard, C), 7.0, dtype=torch.bfloat16, device=DEV) if lo else None
    ops.conv0_gn_gelu(xd, kd, fs, fb, Pair(hi, lo_t), passes)
    torch.cuda.synchronize()
    got = (hi.float() + (lo_t.float() if lo else 0))[: B * T0].view(B, T0, C).cpu()
    err = (got - ref).abs().max().item()
    print(f"conv0 fused L={L} passes={passes}: max err {err:.3e} (|ref| max {ref.abs().max():.2f})")
    if T0 > 1:                                                   # (a single frame has zero variance: rstd = eps^-1/2 amplifies rounding)
        assert err < (0.08 if passes == 1 else 3e-4)
    assert torch.all(hi[B * T0:].float() == 7.0)


@pytest.mark.parametrize("passes", [1, 3])
@pytest.mark.parametrize("L,bias", [(20011, True), (1285, False), (14, True)])
def test_conv0_layernorm_fused(passes, L, bias):
    """Layer 0 of the layer-norm extractor in one kernel (w2v2_conv0_ln_gelu): conv (+ bias), LayerNorm over the 512 channels of a
    frame inside the CTA, GELU (feature_extractor.py:48-50,54-59).  Checked against the oracle's conv -> layer_norm -> gelu."""
    torch.manual_seed(11)
    B, C = 3, 512
    lo = passes == 3
    x = torch.randn(B, L)
    x[1] = x[1] * 0.3 + 0.05
    kern = torch.randn(10, 1, C) * 0.3
    cb = 0.2 * torch.randn(C) if bias else None
    gamma, beta = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    ref = O.gelu_erf(O.layer_norm(O.conv1d_valid(x[:, :, None], kern, cb, stride=5), gamma, beta, 1e-5))
    T0 = ref.shape[1]
    guard = 64
    hi = torch.full((B * T0 + guard, C), 7.0, dtype=torch.bfloat16, device=DEV)
    lo_t = torch.full((B * T0 + guard, C), 7.0, dtype=torch.bfloat16, device=DEV) if lo else None
    ops.conv0_ln_gelu(x.to(DEV), kern.reshape(10, C).to(DEV), None if cb is None else cb.to(DEV), gamma.to(DEV), beta.to(DEV),
                      1e-5, Pair(hi, lo_t), passes)
    torch.cuda.synchronize()
    got = (hi.float() + (lo_t.float() if lo else 0))[: B * T0].view(B, T0, C).cpu()
    err = (got - ref).abs().max().item()
    print(f"conv0 + LayerNorm fused L={L} passes={passes} bias={bias}: max err {err:.3e} (|ref| max {ref.abs().max():.2f})")
    assert err < (0.08 if passes == 1 else 3e-4)
    assert torch.all(hi[B * T0:].float() == 7.0)


@pytest.mark.parametrize("passes", [1, 3])
@pytest.mark.parametrize("T", [145, 768, 49])
def test_attention(passes, T):
    torch.manual_seed(5)
    B, H, dh = 2, 4, 64
    d = H * dh
    lo = passes == 3
    raw = torch.randn(B, T, 3 * d, device=DEV) * 1.5
    raw[:, :, :d] *= dh ** -0.5                                  # q arrives pre-scaled (encoder.py:28 folded into Wq)
    qkv = _pair(raw, lo)
    kv_len = torch.tensor([T, max(1, T - 37)], dtype=torch.int32, device=DEV)
    out = Pair(torch.zeros(B, T, d, dtype=torch.bfloat16, device=DEV), torch.zeros(B, T, d, dtype=torch.bfloat16, device=DEV))
    ops.attn_fwd(qkv, B, T, H, dh, kv_len, out, passes)
    torch.cuda.synchronize()
    x = _eff(qkv, passes).double().cpu()
    q, k, v = (t.reshape(B, T, H, dh).permute(0, 2, 1, 3) for t in x.split(d, dim=-1))
    s = q @ k.transpose(-1, -2)
    mask = torch.arange(T)[None, :] >= kv_len.cpu()[:, None]
    s = s + mask[:, None, None, :] * -10000.0                   # encoder.py:256-263
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B, T, d).float()
    got = (out.hi.float() + (out.lo.float() if lo else 0)).cpu()
    err = (got - ref).abs().max().item()
    print(f"attention T={T} passes={passes}: max err {err:.3e}")
    assert err < (2e-2 if passes == 1 else 1e-4)


def test_attention_one_tile_reference_kernel():
    """W2V2_ATTN_KERNEL=1 keeps the one-tile kernel of attn.cu as the A/B reference of the single-pass modes; the selection is read
    once per process, so the same attention checks run in a child process with the switch set."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, W2V2_ATTN_KERNEL="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-m", "gpu", "-k",
                        "test_attention and not one_tile and not deterministic", "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "6 passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("drop", [None, (0.1, 1234, 7)])
def test_attention_is_deterministic(drop):
    """The two-tile kernel hands S / P / O / row sums between five warp roles through mbarriers; a missing edge in that protocol
    shows as a run-to-run difference long before it shows as a large error.  60 launches on an odd number of q tiles (tile B of
    the last pair is out of range) with ragged key lengths must be bit-identical."""
    torch.manual_seed(3)
    B, T, H, dh = 3, 300, 4, 64
    d = H * dh
    raw = torch.randn(B, T, 3 * d, device=DEV) * 1.5
    raw[:, :, :d] *= dh ** -0.5
    qkv = Pair(raw.to(torch.bfloat16).contiguous(), None)
    kv_len = torch.tensor([T, T - 37, 5], dtype=torch.int32, device=DEV)
    ref = None
    for _ in range(60):
        out = Pair(torch.zeros(B, T, d, dtype=torch.bfloat16, device=DEV), None)
        if drop is None:
            ops.attn_fwd(qkv, B, T, H, dh, kv_len, out, 1)
        else:
            ops.attn_fwd_train(qkv, B, T, H, dh, kv_len, out, 1, drop)
        torch.cuda.synchronize()
        assert not torch.isnan(out.hi.float()).any()
        if ref is None:
            ref = out.hi.clone()
        else:
            assert torch.equal(ref, out.hi)


@pytest.mark.parametrize("passes", [1, 3])
@pytest.mark.parametrize("T,d,groups", [(145, 768, 16), (768, 768, 16), (300, 1024, 16)])
def test_posconv(passes, T, d, groups):
    torch.manual_seed(6)
    B, k = 2, 128
    lo = passes == 3
    cpg = d // groups
    x32 = torch.randn(B, T, d, device=DEV)
    x = _pair(x32, lo)
    kern = torch.randn(k, cpg, d, device=DEV) / math.sqrt(k * cpg)       # already weight-normalised, TF layout
    wp = _pair(kern.reshape(k, cpg // 8, 8, groups, cpg).permute(3, 0, 1, 4, 2).contiguous(), lo)
    bias = torch.randn(d, device=DEV) * 0.1
    out = torch.full((B, T, d), float("nan"), device=DEV)
    ops.posconv(x, wp, bias, x32, out, B, T, d, groups, k, passes)
    torch.cuda.synchronize()
    w_eff = _eff(wp, passes).permute(1, 2, 4, 0, 3).reshape(k, cpg, d)    # undo the packing
    xe = _eff(x, passes).cpu()
    conv = O.conv1d_valid(F.pad(xe, (0, 0, k // 2, k // 2)), w_eff.cpu(), bias.cpu(), 1, groups)[:, :-1]
    ref = x32.cpu() + O.gelu_erf(conv)
    err = (out.cpu() - ref).abs().max().item()
    print(f"posconv T={T} d={d} passes={passes}: max err {err:.3e}")
    assert err < (3e-3 if passes == 1 else 3e-4)


def test_ctc_loss_and_grad():
    torch.manual_seed(7)
    cfg = Wav2Vec2Config()
    B, T, V = 3, 120, 32
    logits = torch.randn(B, T, V) * 2
    np.random.seed(0)
    labels = torch.from_numpy(np.random.randint(1, 30, size=(B, 24))).int()
    labels[1, 10:] = 0
    labels[2, 5:7] = labels[2, 4]                               # repeated labels
    loss, grad = ops.ctc_loss(logits.to(DEV), labels.to(DEV), cfg.pad_id, 1.0 / 4.0)
    torch.cuda.synchronize()
    ref_total, ref_per = O.ctc_loss(labels, logits, cfg, division_factor=4.0)
    ref_total2, ref_grad = O.ctc_loss_and_grad(labels.long(), logits, cfg, division_factor=4.0)
    assert abs(ref_total - ref_total2) < 1e-6 * abs(ref_total)
    assert abs(loss.sum().item() - ref_total) < 1e-3             # tests/test_wav2vec2.py:235-237 tolerance
    assert np.allclose(loss.cpu().numpy() * 4.0, np.array(ref_per), atol=1e-3)
    assert (grad.cpu() - ref_grad.float()).abs().max().item() < 1e-4


def test_frame_argmax():
    torch.manual_seed(8)
    x = torch.randn(4, 77, 32, device=DEV)
    assert torch.equal(ops.frame_argmax(x).long(), x.argmax(-1))


def test_processor_normalize_on_device_matches_golden():
    """Wav2Vec2Processor._normalize on the GPU: the reference's known answer (tests/test_dataloader.py:56-63 via
    tests/golden/processor.npz), padded batches normalised over their real samples only."""
    import os
    import numpy as np
    from wav2vec2 import Wav2Vec2Processor
    G = os.path.join(os.path.dirname(__file__), "golden")
    wav = torch.from_numpy(O.read_wav_s16(os.path.join(G, "sample.wav")).astype(np.float32))
    want = torch.from_numpy(O.normalize_utterance(wav.numpy()))
    got = Wav2Vec2Processor(is_tokenizer=False)(wav.to(DEV)).cpu()
    assert got.shape == want.shape and (got - want).abs().max().item() < 2e-6
    z = np.load(os.path.join(G, "processor.npz"))
    assert np.allclose(got[32:40].numpy(), z["reference_vector"], atol=1e-6)     # the reference's own golden vector
    L = 50000
    batch = torch.zeros(3, L)
    lens = torch.tensor([L, 46797, 1234], dtype=torch.int32)
    batch[0] = torch.randn(L) * 3 + 1
    batch[1, :46797] = wav
    batch[2, :1234] = torch.randn(1234) * 0.01
    out = ops.normalize_utterances(batch.to(DEV), lens.to(DEV)).cpu()
    for b in range(3):
        n = int(lens[b])
        ref = torch.from_numpy(O.normalize_utterance(batch[b, :n].numpy()))
        assert (out[b, :n] - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())
        assert torch.all(out[b, n:] == 0)


@pytest.mark.parametrize("M,K,N,cluster", [(128 * 9 + 17, 768, 768, 0), (300, 3072, 768, 0), (98, 768, 768, 1), (257, 256, 192, 1)])
def test_gemm_residual_is_layernorm_recomputed(M, K, N, cluster):
    """w2v2_gemm_args.res_ln_*: out = A W^T + bias + LayerNorm(residual) with the (mean, rstd) of w2v2_ln_rows_stats, also IN PLACE
    (out == residual), in the 2-SM kernel (TMA-fetched and per-lane residual paths) and the 1-SM kernel."""
    torch.manual_seed(11)
    a = _pair(torch.randn(M, K, device=DEV), False)
    w = _pair(torch.randn(N, K, device=DEV) / K ** 0.5, False)
    bias = torch.randn(N, device=DEV)
    y = torch.randn(M, N, device=DEV) * 2 + 0.5
    gamma, beta = 1 + 0.1 * torch.randn(N, device=DEV), 0.1 * torch.randn(N, device=DEV)
    ln_f32 = torch.empty(M, N, device=DEV)
    stats = torch.empty(M, 2, device=DEV)
    ops.ln_rows(y, gamma, beta, 1e-5, M, N, out_f32=ln_f32, stats=stats)
    want = torch.empty(M, N, device=DEV)
    ops.gemm(a, w, K=K, N=N, rows_per_batch=M, bias=bias, residual=ln_f32, out_f32=want, cluster=cluster)
    got = y.clone()
    ops.gemm(a, w, K=K, N=N, rows_per_batch=M, bias=bias, residual=got, res_ln=(stats, gamma, beta), out_f32=got, cluster=cluster)
    torch.cuda.synchronize()
    mean, var = y.mean(-1), y.var(-1, unbiased=False)
    assert (stats[:, 0] - mean).abs().max().item() < 1e-5 and (stats[:, 1] - torch.rsqrt(var + 1e-5)).abs().max().item() < 1e-4
    assert torch.equal(got, want)              # same arithmetic as the LayerNorm kernel: bit-identical


@pytest.mark.parametrize("passes", [1, 17, 25])
def test_gemm_layernorm_fold_and_row_statistics(passes):
    """The LayerNorm fold at kernel level (include/w2v2.h ln_fold_* / row_stats_out / w2v2_row_stats_finalize):
    producer  y = r + x W1 + b1 (residual GEMM) also writes the operand planes of y and per-row partial (sum, sum of squares);
    consumer  LN(y) W2 + b2 computed from those planes with gamma folded into W2 and mean / rstd applied in the epilogue -
    against the explicit fp64 LayerNorm + matmul of the same fp32 data."""
    from wav2vec2.modeling import _KINDS, _effective_weight, _split
    torch.manual_seed(7)
    M, K1, d, N2 = 128 * 5 + 17, 256, 768, 512
    kind = {1: "bf16", 17: "fp16", 25: "fp16f8"}[passes]
    hi_dt, lo_kind, ofmt = _KINDS[kind]

    def planes_of(x):          # operand planes of an fp32 activation, as a producing kernel would write them
        if passes == 1:
            return Pair(x.to(torch.bfloat16).contiguous(), None)
        hi = torch.empty(x.shape, dtype=torch.float16, device=DEV)
        lo = torch.empty(x.shape[0], 2 * x.shape[1], dtype=torch.uint8, device=DEV) if passes == 25 else None
        s = torch.clamp(x * 16.0, -65504, 65504)
        hi.copy_(s.to(torch.float16))
        if lo is not None:
            res = s - hi.float()
            rows, C = x.shape
            l8 = torch.clamp(res * 64, -448, 448).to(torch.float8_e4m3fn).view(torch.uint8).reshape(rows, C // 64, 1, 64)
            h8 = torch.clamp(hi.float() / 64, -448, 448).to(torch.float8_e4m3fn).view(torch.uint8).reshape(rows, C // 64, 1, 64)
            lo.copy_(torch.cat([l8, h8], 2).reshape(rows, 2 * C))
        return Pair(hi, lo)
    x = torch.randn(M, K1, device=DEV)
    r = torch.randn(M, d, device=DEV) * 2 + 0.5                      # residual with a non-zero mean per row
    w1, b1 = torch.randn(d, K1, device=DEV) / math.sqrt(K1), torch.randn(d, device=DEV)
    gamma, beta = 1 + 0.2 * torch.randn(d, device=DEV), 0.3 * torch.randn(d, device=DEV)
    w2, b2 = torch.randn(N2, d, device=DEV) / math.sqrt(d), torch.randn(N2, device=DEV)
    # ---- producer
    y = r.clone()
    ys = Pair(torch.empty(M, d, dtype=hi_dt, device=DEV), torch.empty(M, 2 * d, dtype=torch.uint8, device=DEV) if passes == 25 else None)
    parts = torch.full((d // 64, M, 2), float("nan"), device=DEV)
    ops.gemm(planes_of(x), _split(w1, passes), K=K1, N=d, rows_per_batch=M, bias=b1, residual=y, out_f32=y, out_hi=ys.hi, out_lo=ys.lo,
             out_format=ofmt, row_stats_out=parts, passes=passes)
    stats = ops.row_stats_finalize(parts, d, 1e-5, torch.empty(M, 2, device=DEV))
    # the same producer finalising its statistics ITSELF (row_stats_final: the last column group of a 32-row block reduces the partials;
    # cta_group::2 kernel at this size) must give the stand-alone launch's result bit for bit and leave its arrival counters at zero;
    # a single-m-tile problem takes the other tile shape, where the call runs the stand-alone kernel on the same stream
    for rows in (M, 100):
        y2 = r[:rows].clone()
        parts2 = torch.full((d // 64, rows, 2), float("nan"), device=DEV)
        fin = torch.full((rows, 2), float("nan"), device=DEV)
        counters = torch.zeros((rows + 31) // 32, dtype=torch.int32, device=DEV)
        xr = planes_of(x[:rows].contiguous())
        for _ in range(2):                                              # twice: the counters re-arm themselves
            y2.copy_(r[:rows])
            ops.gemm(xr, _split(w1, passes), K=K1, N=d, rows_per_batch=rows, bias=b1, residual=y2, out_f32=y2, out_hi=ys.hi[:rows],
                     out_lo=None if ys.lo is None else ys.lo[:rows], out_format=ofmt, row_stats_out=parts2, passes=passes,
                     row_stats_final=(fin, counters))
        want = ops.row_stats_finalize(parts2, d, 1e-5, torch.empty(rows, 2, device=DEV))
        torch.cuda.synchronize()
        assert torch.equal(fin, want) and int(counters.abs().sum()) == 0
        if rows == M:
            assert torch.equal(fin, stats) and torch.equal(y2, y)
    y_ref = r.double() + x.double() @ w1.double().t() + b1.double()
    tol_y = {1: 5e-2, 17: 5e-3, 25: 3e-4}[passes]
    assert (y.double() - y_ref).abs().max().item() < tol_y
    mean, var = y.double().mean(1), y.double().var(1, unbiased=False)
    assert (stats[:, 0].double() - mean).abs().max().item() < 1e-4
    assert (stats[:, 1].double() * torch.sqrt(var + 1e-5) - 1).abs().max().item() < 1e-4
    # ---- consumer with the folded LayerNorm
    wf = _split(w2 * gamma[None, :], passes)
    cs = _effective_weight(wf, passes).sum(1).contiguous()
    bf = (b2 + w2 @ beta).contiguous()
    out = torch.empty(M, N2, device=DEV)
    ops.gemm(ys, wf, K=d, N=N2, rows_per_batch=M, bias=bf, out_f32=out, passes=passes, ln_fold=(stats, cs), ln_eps=1e-5)
    torch.cuda.synchronize()
    ln = torch.nn.functional.layer_norm(y.double(), (d,), gamma.double(), beta.double(), 1e-5)
    ref = ln @ w2.double().t() + b2.double()
    err = (out.double() - ref).abs().max().item()
    print(f"LayerNorm fold, mode {passes}: max err {err:.3e}")
    assert err < {1: 1.5e-1, 17: 2e-2, 25: 1.5e-3}[passes]
