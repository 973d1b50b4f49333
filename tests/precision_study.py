"""CPU numerical study (not a collected test): which tensor-core operand format keeps the logits within the reference's
1e-3 of the fp32 oracle?  Every contraction of the oracle forward (convs, Dense layers, Q K^T, P V) is re-run with its two
operands quantised the way a candidate mode would feed the tensor cores; products / accumulation stay fp32 like TMEM.

  python tests/precision_study.py [--seq 246000] [--layers 12] [--modes bf16,fp16,...]

Modes:  bf16 / fp16: one MMA on rounded operands;  bf16x3 / fp16x3: hi*hi + lo*hi + hi*lo;  fp16x2: (hi + lo) * hi (only the
activation is split);  fp16f8: fp16 hi*hi + the two cross terms in e4m3 (x_lo 2^10 * w_hi 2^5 and x_hi * w_lo 2^15, / 2^15)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from oracle import w2v2_oracle as O  # noqa: E402
from wav2vec2 import Wav2Vec2Config  # noqa: E402

MODE = {"name": "fp32"}
SITE = {"cur": "dense"}     # which kind of contraction is running: conv / dense / qk / pv


def site_mode():
    """`MODE["name"]` is either one mode for every site or `site=mode,...,*=mode`."""
    m = MODE["name"]
    if "=" not in m:
        return m
    table = dict(kv.split("=") for kv in m.split("/"))
    return table.get(SITE["cur"], table.get("*", "fp32"))


# static power-of-two scales of the e4m3 cross terms per site kind: (a_lo, b_hi, a_hi, b_lo), a_lo + b_hi = a_hi + b_lo = 15
F8_SCALES = {"dense": (8, 7, -3, 18), "conv": (8, 7, -3, 18), "pos": (8, 7, -3, 18), "qk": (12, 3, 3, 12), "pv": (11, 4, 8, 7)}


def _r(x, dt):
    return x.to(dt).float()


def _e4m3(x):
    return x.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float()


def terms(a, b):
    """List of (a_i, b_i, scale) operand pairs whose products sum to the mode's approximation of a . b."""
    m = site_mode()
    if m == "fp32":
        return [(a, b, 1.0)]
    if m in ("bf16", "fp16"):
        dt = torch.bfloat16 if m == "bf16" else torch.float16
        return [(_r(a, dt), _r(b, dt), 1.0)]
    dt = torch.bfloat16 if m.startswith("bf16") else torch.float16
    ah, bh = _r(a, dt), _r(b, dt)
    al, bl = a - ah, b - bh
    if m in ("bf16x3", "fp16x3"):
        return [(ah, bh, 1.0), (_r(al, dt), bh, 1.0), (ah, _r(bl, dt), 1.0)]
    if m == "fp16x2":
        return [(ah, bh, 1.0), (_r(al, dt), bh, 1.0)]
    if m in ("fp16f8", "fp16f8w", "fp16f8a"):
        s = 2.0 ** -15
        sal, sbh, sah, sbl = (2.0 ** e for e in F8_SCALES[SITE["cur"]])
        out = [(ah, bh, 1.0)]
        if m != "fp16f8w":
            out.append((_e4m3(al * sal), _e4m3(bh * sbh), s))
        if m != "fp16f8a":
            out.append((_e4m3(ah * sah), _e4m3(bl * sbl), s))
        return out
    if m == "f8s":
        # the recipe as the kernels would implement it: operands pre-scaled (activation x 2^4, weight x 2^11; both x 2^4 when the
        # "weight" is an activation too), fp16 main product, e4m3 cross terms at 2^+-6, everything at scale 2^15 (2^8) in ONE
        # fp32 accumulator, un-scaled by the epilogue
        sa, sb = 2.0 ** 4, (2.0 ** 4 if SITE["cur"] in ("qk", "pv") else 2.0 ** 11)
        a2, b2 = (a * sa).clamp(-65504, 65504), (b * sb).clamp(-65504, 65504)
        ah, bh = _r(a2, torch.float16), _r(b2, torch.float16)
        al, bl = a2 - ah, b2 - bh
        s = 1.0 / (sa * sb)
        return [(ah, bh, s), (_e4m3(al * 64.0), _e4m3(bh / 64.0), s), (_e4m3(ah / 64.0), _e4m3(bl * 64.0), s)]
    if m == "f16s":
        sa, sb = 2.0 ** 4, (2.0 ** 4 if SITE["cur"] in ("qk", "pv") else 2.0 ** 11)
        a2, b2 = (a * sa).clamp(-65504, 65504), (b * sb).clamp(-65504, 65504)
        return [(_r(a2, torch.float16), _r(b2, torch.float16), 1.0 / (sa * sb))]
    raise ValueError(m)


def q_dense(x, kernel, bias):
    # finer site names for the Dense layers, from their shapes: proj 512->d, qkv/out d->d, ffn1 d->4d, ffn2 4d->d, lm d->vocab
    kin, kout = kernel.shape
    SITE["cur"] = SITE.get("force") or ("lm" if kout < 64 else "proj" if kin == 512 else "ffn1" if kout > kin else "ffn2" if kin > kout else "dd")
    return sum(s * (a @ b) for a, b, s in terms(x, kernel)) + bias


def q_conv(x, kernel, bias=None, stride=1, groups=1):
    SITE["cur"] = "conv" if groups == 1 else "pos"      # the grouped conv is the positional embedding
    y = 0.0
    for a, b, s in terms(x, kernel):
        y = y + s * F.conv1d(a.transpose(1, 2), b.permute(2, 1, 0).contiguous(), stride=stride, groups=groups)
    if bias is not None:
        y = y + bias[None, :, None]
    return y.transpose(1, 2)


def q_attention(x, p, cfg, base, additive_mask=None, drop=None, layer=0):
    B, T, D = x.shape
    H = cfg.num_heads
    dh = D // H

    def split(t):
        return t.reshape(B, T, H, dh).permute(0, 2, 1, 3)
    SITE["force"] = "qkv"
    q = split(q_dense(x, p[base + "q_proj/kernel"], p[base + "q_proj/bias"])) * dh ** (-0.5)
    k = split(q_dense(x, p[base + "k_proj/kernel"], p[base + "k_proj/bias"]))
    v = split(q_dense(x, p[base + "v_proj/kernel"], p[base + "v_proj/bias"]))
    SITE["force"] = None
    SITE["cur"] = "qk"
    scores = sum(s * (a @ b.transpose(-1, -2)) for a, b, s in terms(q, k))
    if additive_mask is not None:
        scores = scores + additive_mask
    pr = torch.softmax(scores, dim=-1)
    SITE["cur"] = "pv"
    ctx = sum(s * (a @ b) for a, b, s in terms(pr, v))
    ctx = ctx.permute(0, 2, 1, 3).reshape(B, T, D)
    SITE["force"] = "out"
    y = q_dense(ctx, p[base + "out_proj/kernel"], p[base + "out_proj/bias"])
    SITE["force"] = None
    return y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq", type=int, default=246000)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--modes", default="bf16,fp16,fp16x2,fp16f8,bf16x3")
    ap.add_argument("--robust", action="store_true", help="large / robust architecture (24 layers, d = 1024)")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    from wav2vec2 import RobustWav2Vec2Config
    cfg = RobustWav2Vec2Config(num_layers=args.layers) if args.robust else Wav2Vec2Config(num_layers=args.layers)
    params = O.random_params(cfg, seed=args.seed)
    x = torch.randn(args.batch, args.seq, generator=torch.Generator().manual_seed(args.seed))
    with torch.no_grad():
        ref = O.wav2vec2_for_ctc(x, params, cfg)
        O.dense, O.conv1d_valid, O.attention = q_dense, q_conv, q_attention
        for mode in ["fp32"] + args.modes.split(","):
            MODE["name"] = mode
            t0 = time.perf_counter()
            got = O.wav2vec2_for_ctc(x, params, cfg)
            print(json.dumps({"mode": mode, "logits_max_abs": round(ref.abs().max().item(), 3),
                              "max_abs_err": (got - ref).abs().max().item(),
                              "rms_err": (got - ref).pow(2).mean().sqrt().item(),
                              "argmax_agreement": (got.argmax(-1) == ref.argmax(-1)).float().mean().item(),
                              "seconds": round(time.perf_counter() - t0, 1)}), flush=True)


if __name__ == "__main__":
    main()
