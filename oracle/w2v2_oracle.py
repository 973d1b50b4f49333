"""CPU fp32 oracle for the Wav2Vec2 forward / CTC path.  TEST INFRASTRUCTURE ONLY.

This file restates, op by op, what thevasudevgupta/gsoc-wav2vec2 computes in its
TensorFlow-2 graph, using plain torch CPU tensor ops on the reference's own
weight layouts (Conv kernels ``[k, Cin/groups, Cout]``, Dense kernels
``[in, out]``, activations channels-last ``[B, T, C]``).  Every function cites
the reference file:line it follows.  Nothing in the shipped package imports it:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may, and there only as the checker or
as the CPU baseline being timed.

Pinning (see DESIGN.md "Oracle"): the reference cannot run in this image (its
arithmetic lives in the un-vendored ``tensorflow==2.5`` wheel), so the oracle
is pinned against (1) the known-answer vector the reference's own test holds for
the processor (tests/test_dataloader.py:56-63), (2) the implementation the
reference's tests equate the TF model to at atol 1e-3 / 4e-3 —
``transformers`` PyTorch Wav2Vec2 (tests/test_wav2vec2.py:77-79,155-157,231-237)
— on shared seeded weights (agreement ~2e-6), and (3) the reference's
self-contained weight-norm conv test (tests/test_wav2vec2.py:239-282).
Golden vectors made that way are committed under tests/golden/.

Parameters are a flat ``dict[str, torch.Tensor]`` keyed by the reference's Keras
variable names without the ``:0`` suffix, e.g.
``wav2vec2/encoder/layers/3/attention/q_proj/kernel`` (the naming produced by
src/convert_torch_to_tf.py:12-18,38-44).
"""
import math
from itertools import groupby
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- primitives
def conv1d_valid(x, kernel, bias=None, stride=1, groups=1):
    """Keras ``Conv1D(padding='valid')`` on channels-last input.

    x [B,T,Cin], kernel [k, Cin/groups, Cout] (TF layout) -> [B, 1+(T-k)//s, Cout]
    (feature_extractor.py:31-37, tensorflow_addons.py:53).
    """
    w = kernel.permute(2, 1, 0).contiguous()  # -> [Cout, Cin/g, k]
    y = F.conv1d(x.transpose(1, 2), w, bias=bias, stride=stride, groups=groups)
    return y.transpose(1, 2)


def gelu_erf(x):
    """``tf.nn.gelu(approximate=False)`` (config.py:14; feature_extractor.py:58)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def gelu(x, cfg):
    """``tf.nn.gelu(batch, approximate=self.is_gelu_approx)`` (config.py:14; feature_extractor.py:58; encoder.py:127):
    erf form by default, 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) when ``is_gelu_approx``."""
    if getattr(cfg, "is_gelu_approx", False):
        return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))
    return gelu_erf(x)


def layer_norm(x, gamma, beta, eps):
    """Keras LayerNormalization over the last axis, biased variance."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * gamma + beta


def group_norm_per_channel(x, gamma, beta, eps=1e-5):
    """GroupNormalization(groups == channels): statistics over TIME per (b, c).

    tensorflow_addons.py:207-231 with the instance-norm branch (:162,211-216):
    ``tf.nn.moments`` over axis 1 (biased variance) then batch_normalization.
    """
    mu = x.mean(1, keepdim=True)
    var = ((x - mu) ** 2).mean(1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * gamma + beta


def dense(x, kernel, bias):
    """Keras Dense with ``[in, out]`` kernel."""
    return x @ kernel + bias


def weight_norm_kernel(weight_v, weight_g):
    """``l2_normalize(v, axes != 0) * g`` (tensorflow_addons.py:16-21,26-28).

    v [k, Cin/g, Cout], g [k,1,1]; l2_normalize = v * rsqrt(max(sum v^2, 1e-12)).
    """
    ss = (weight_v ** 2).sum(dim=(1, 2), keepdim=True)
    return weight_v * torch.rsqrt(torch.clamp(ss, min=1e-12)) * weight_g


# --------------------------------------------------------------------------- blocks
def feature_extractor_layer(x, p: Params, cfg, i: int, prefix="wav2vec2/"):
    """conv -> [norm] -> gelu (feature_extractor.py:54-59; norm choice :39-52)."""
    base = f"{prefix}feature_extractor/conv_layers/{i}/"
    bias = p.get(base + "conv/bias") if cfg.conv_bias else None
    y = conv1d_valid(x, p[base + "conv/kernel"], bias, stride=cfg.strides[i])
    if cfg.feature_extractor_norm_type == "group":
        if i == 0:
            y = group_norm_per_channel(y, p[base + "layer_norm/gamma"], p[base + "layer_norm/beta"], 1e-5)
    else:
        y = layer_norm(y, p[base + "layer_norm/gamma"], p[base + "layer_norm/beta"], 1e-5)
    return gelu(y, cfg)


def _drop(x, drop, key):
    """tf.keras.layers.Dropout in training mode with an EXPLICIT mask: ``drop[key]`` holds keep / (1 - rate) per element
    (0 where dropped).  ``drop is None`` (or a missing key) = eval mode / rate 0: identity."""
    if drop is None or key not in drop:
        return x
    return x * drop[key].to(x.dtype)


def feature_projection(x, p: Params, cfg, prefix="wav2vec2/", drop=None):
    """LN(512) -> Dense -> dropout (feature_extractor.py:92-95)."""
    base = f"{prefix}feature_projection/"
    y = layer_norm(x, p[base + "layer_norm/gamma"], p[base + "layer_norm/beta"], cfg.layer_norm_eps)
    return _drop(dense(y, p[base + "projection/kernel"], p[base + "projection/bias"]), drop, "proj")


def positional_conv_embedding(x, p: Params, cfg, prefix="wav2vec2/"):
    """Weight-normalised grouped conv + GELU (encoder.py:177-181; addons :50-53)."""
    base = f"{prefix}encoder/pos_conv_embed/conv/"
    k = cfg.num_conv_pos_embeddings
    kernel = weight_norm_kernel(p[base + "weight_v"], p[base + "weight_g"])
    pad = k // 2
    xp = F.pad(x, (0, 0, pad, pad))  # explicit zero pad on time (tensorflow_addons.py:52)
    y = conv1d_valid(xp, kernel, p[base + "bias"], stride=1, groups=cfg.num_conv_pos_embedding_groups)
    if k % 2 == 0:
        y = y[:, :-1, :]  # encoder.py:175,179-180
    return gelu(y, cfg)


def attention(x, p: Params, cfg, base: str, additive_mask=None, drop=None, layer=0):
    """encoder.py:22-54: q/k/v Dense, q scaled AFTER bias, softmax(QK^T + mask) V, out_proj."""
    B, T, D = x.shape
    H = cfg.num_heads
    dh = D // H

    def split(t):  # [B,T,D] -> [B,H,T,dh] (encoder.py:49-54)
        return t.reshape(B, T, H, dh).permute(0, 2, 1, 3)

    q = split(dense(x, p[base + "q_proj/kernel"], p[base + "q_proj/bias"])) * dh ** (-0.5)
    k = split(dense(x, p[base + "k_proj/kernel"], p[base + "k_proj/bias"]))
    v = split(dense(x, p[base + "v_proj/kernel"], p[base + "v_proj/bias"]))
    scores = q @ k.transpose(-1, -2)
    if additive_mask is not None:
        scores = scores + additive_mask
    ctx = _drop(torch.softmax(scores, dim=-1), drop, f"attn_probs.{layer}") @ v          # encoder.py:41-43
    ctx = ctx.permute(0, 2, 1, 3).reshape(B, T, D)
    return dense(ctx, p[base + "out_proj/kernel"], p[base + "out_proj/bias"])


def transformer_layer(x, p: Params, cfg, i: int, additive_mask=None, prefix="wav2vec2/", drop=None):
    """encoder.py:111-134 (eval: dropout = id, StochasticDepth = add, addons :386-390)."""
    base = f"{prefix}encoder/layers/{i}/"
    eps = cfg.layer_norm_eps
    pre = cfg.attention_norm_type == "prenorm"
    res = x
    if pre:
        x = layer_norm(x, p[base + "layer_norm/gamma"], p[base + "layer_norm/beta"], eps)
    x = _drop(attention(x, p, cfg, base + "attention/", additive_mask, drop, i), drop, f"attn_out.{i}") + res   # :117-119
    if not pre:
        x = layer_norm(x, p[base + "layer_norm/gamma"], p[base + "layer_norm/beta"], eps)
    res = x
    if pre:
        x = layer_norm(x, p[base + "final_layer_norm/gamma"], p[base + "final_layer_norm/beta"], eps)
    h = _drop(gelu(dense(x, p[base + "feed_forward/intermediate_dense/kernel"],
                             p[base + "feed_forward/intermediate_dense/bias"]), cfg), drop, f"ffn_mid.{i}")             # :127-128
    branch = dense(h, p[base + "feed_forward/output_dense/kernel"], p[base + "feed_forward/output_dense/bias"])
    # StochasticDepth (tensorflow_addons.py:374-394): eval = plain add; training = shortcut + b * residual with ONE Bernoulli
    # draw b per layer call (explicit here: drop["stochastic_depth.<i>"] in {0, 1})
    x = res + _drop(branch, drop, f"stochastic_depth.{i}")
    if not pre:
        x = layer_norm(x, p[base + "final_layer_norm/gamma"], p[base + "final_layer_norm/beta"], eps)
    return x


def frame_lengths(cfg, sample_lengths: torch.Tensor) -> torch.Tensor:
    """modeling.py:201-204 / losses.py:47-56: L <- 1 + (L - k) // s for each conv."""
    n = sample_lengths.clone().to(torch.int64)
    for k, s in zip(cfg.kernal_sizes, cfg.strides):
        n = 1 + torch.div(n - k, s, rounding_mode="floor")
    return n


def encoder(x, p: Params, cfg, frame_mask: Optional[torch.Tensor] = None, prefix="wav2vec2/",
            num_layers: Optional[int] = None, drop=None):
    """encoder.py:251-276.  frame_mask: bool [B,T] (True = real frame) or None."""
    additive = None
    if frame_mask is not None:
        x = torch.where(frame_mask[:, :, None], x, torch.zeros((), dtype=x.dtype))
        additive = (1.0 - frame_mask.to(x.dtype)) * -10000.0  # encoder.py:256-257
        additive = additive[:, None, None, :]                  # over keys: [B,1,1,Tk]
    x = x + positional_conv_embedding(x, p, cfg, prefix)
    base = f"{prefix}encoder/layer_norm/"
    if cfg.attention_norm_type == "postnorm":
        x = layer_norm(x, p[base + "gamma"], p[base + "beta"], cfg.layer_norm_eps)
    x = _drop(x, drop, "enc")                                                              # encoder.py:270
    for i in range(cfg.num_layers if num_layers is None else num_layers):
        x = transformer_layer(x, p, cfg, i, additive, prefix, drop)
    if cfg.attention_norm_type == "prenorm":
        x = layer_norm(x, p[base + "gamma"], p[base + "beta"], cfg.layer_norm_eps)
    return x


def apply_time_mask(features, masked_spec_embed, mask_indices):
    """spec_augment.py:119-128: ``where(mask[:, :, None], masked_spec_embed, features)``."""
    return torch.where(mask_indices.bool()[:, :, None], masked_spec_embed[None, None, :], features)


def wav2vec2_model(speech, p: Params, cfg, attention_mask=None, spec_mask=None, prefix="wav2vec2/",
                   return_intermediates=False, drop=None):
    """Wav2Vec2Model.call (modeling.py:169-209).  speech [B,L] fp32 -> [B,T',hidden]."""
    inter = {}
    x = speech[:, :, None]                                   # :188
    for i in range(len(cfg.filter_sizes)):                   # :189-190
        x = feature_extractor_layer(x, p, cfg, i, prefix)
        inter[f"conv{i}"] = x
    x = feature_projection(x, p, cfg, prefix, drop)          # :191
    inter["proj"] = x
    if spec_mask is not None:                                # :193-199 (training only)
        x = apply_time_mask(x, p[f"{prefix}masked_spec_embed"], spec_mask)
    frame_mask = None
    if attention_mask is not None:                           # :201-206
        n = frame_lengths(cfg, attention_mask.to(torch.int64).sum(-1))
        frame_mask = torch.arange(x.shape[1])[None, :] < n[:, None]
    x = encoder(x, p, cfg, frame_mask, prefix, drop=drop)    # :208
    if return_intermediates:
        return x, inter
    return x


def wav2vec2_for_ctc(speech, p: Params, cfg, attention_mask=None, spec_mask=None, drop=None):
    """Wav2Vec2ForCTC.call (modeling.py:239-255); variables live under ``wav2vec2-ctc/``
    in the reference, here the inner model keeps the ``wav2vec2/`` prefix."""
    h = wav2vec2_model(speech, p, cfg, attention_mask, spec_mask, drop=drop)
    return dense(_drop(h, drop, "head"), p["lm_head/kernel"], p["lm_head/bias"])          # modeling.py:252-254


# --------------------------------------------------------------------------- CTC loss
def ctc_loss(labels, logits, cfg, division_factor=1.0):
    """CTCLoss.call (losses.py:14-45): blank = pad_id, label_length = #non-pad,
    logit_length = T' for every sample, SUM over the batch, / division_factor.

    Log-space alpha recursion written out (what ``tf.nn.ctc_loss`` computes); float64
    accumulation so it can referee fp32 kernels.
    """
    B, T, V = logits.shape
    logp = torch.log_softmax(logits.double(), dim=-1)
    blank = cfg.pad_id
    total = 0.0
    per_sample = []
    for b in range(B):
        lab = [int(t) for t in labels[b].tolist() if int(t) != blank]
        ext = [blank]
        for t in lab:
            ext += [t, blank]
        S = len(ext)
        neg = -float("inf")
        alpha = [neg] * S
        alpha[0] = float(logp[b, 0, blank])
        if S > 1:
            alpha[1] = float(logp[b, 0, ext[1]])
        for t in range(1, T):
            new = [neg] * S
            for s in range(S):
                cands = [alpha[s]]
                if s >= 1:
                    cands.append(alpha[s - 1])
                if s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]:
                    cands.append(alpha[s - 2])
                m = max(cands)
                if m == neg:
                    continue
                new[s] = m + math.log(sum(math.exp(c - m) for c in cands)) + float(logp[b, t, ext[s]])
            alpha = new
        tail = [alpha[S - 1]] + ([alpha[S - 2]] if S > 1 else [])
        m = max(tail)
        ll = m + math.log(sum(math.exp(c - m) for c in tail)) if m != neg else neg
        per_sample.append(-ll)
        total += -ll
    return total / division_factor, per_sample


def ctc_loss_and_grad(labels, logits, cfg, division_factor=1.0):
    """Loss and d(loss)/d(logits) through torch's own CTC (float64), used to referee
    the CUDA alpha-beta kernel's gradient.  Same conventions as ``ctc_loss`` above."""
    B, T, V = logits.shape
    x = logits.double().clone().requires_grad_(True)
    lp = torch.log_softmax(x, -1).transpose(0, 1)
    lens = (labels != cfg.pad_id).sum(-1)
    # torch wants the targets left-packed; the reference counts non-pad entries (losses.py:32-33)
    packed = torch.zeros_like(labels)
    for b in range(B):
        row = labels[b][labels[b] != cfg.pad_id]
        packed[b, : len(row)] = row
    loss = F.ctc_loss(lp, packed, torch.full((B,), T, dtype=torch.long), lens,
                      blank=cfg.pad_id, reduction="sum", zero_infinity=False) / division_factor
    (g,) = torch.autograd.grad(loss, x)
    return float(loss), g


# --------------------------------------------------------------------------- processor
def normalize_utterance(x: np.ndarray) -> np.ndarray:
    """Wav2Vec2Processor._normalize (processor.py:101-106): (x-mean)/sqrt(var+1e-5), biased var."""
    x = np.asarray(x, dtype=np.float32)
    mean = x.mean(axis=-1, keepdims=True)
    var = x.var(axis=-1, keepdims=True)
    return np.squeeze((x - mean) / np.sqrt(var + 1e-5))


def greedy_ctc_decode(ids: List[int], id_to_token: Dict[int, str], pad_id=0, delimiter="|",
                      unk="<unk>") -> str:
    """Wav2Vec2Processor.decode (processor.py:71-89): collapse repeats, drop <pad>, '|' -> ' '."""
    ids = [k for k, _ in groupby(ids)]
    ids = [k for k in ids if k != pad_id]
    toks = [id_to_token.get(k, unk) for k in ids]
    return "".join(" " if t == delimiter else t for t in toks).strip()


def read_wav_s16(path: str) -> np.ndarray:
    """``tf.audio.decode_wav``: int16 PCM -> float32 in [-1, 1) (scale 1/32768)."""
    import wave
    with wave.open(path, "rb") as w:
        assert w.getsampwidth() == 2 and w.getnchannels() == 1
        raw = w.readframes(w.getnframes())
    return (np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0)


# --------------------------------------------------------------------------- parameters
def param_shapes(cfg, with_head=True) -> Dict[str, tuple]:
    """Every variable of the reference model with its TF shape."""
    d, ff = cfg.hidden_size, cfg.intermediate_size
    shapes = {"wav2vec2/masked_spec_embed": (d,)}
    cin = 1
    for i, (c, k) in enumerate(zip(cfg.filter_sizes, cfg.kernal_sizes)):
        base = f"wav2vec2/feature_extractor/conv_layers/{i}/"
        shapes[base + "conv/kernel"] = (k, cin, c)
        if cfg.conv_bias:
            shapes[base + "conv/bias"] = (c,)
        if cfg.feature_extractor_norm_type == "layer" or i == 0:
            shapes[base + "layer_norm/gamma"] = (c,)
            shapes[base + "layer_norm/beta"] = (c,)
        cin = c
    shapes["wav2vec2/feature_projection/layer_norm/gamma"] = (cin,)
    shapes["wav2vec2/feature_projection/layer_norm/beta"] = (cin,)
    shapes["wav2vec2/feature_projection/projection/kernel"] = (cin, d)
    shapes["wav2vec2/feature_projection/projection/bias"] = (d,)
    k, g = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
    shapes["wav2vec2/encoder/pos_conv_embed/conv/weight_v"] = (k, d // g, d)
    shapes["wav2vec2/encoder/pos_conv_embed/conv/weight_g"] = (k, 1, 1)
    shapes["wav2vec2/encoder/pos_conv_embed/conv/bias"] = (d,)
    shapes["wav2vec2/encoder/layer_norm/gamma"] = (d,)
    shapes["wav2vec2/encoder/layer_norm/beta"] = (d,)
    for i in range(cfg.num_layers):
        base = f"wav2vec2/encoder/layers/{i}/"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            shapes[base + f"attention/{n}/kernel"] = (d, d)
            shapes[base + f"attention/{n}/bias"] = (d,)
        shapes[base + "layer_norm/gamma"] = (d,)
        shapes[base + "layer_norm/beta"] = (d,)
        shapes[base + "feed_forward/intermediate_dense/kernel"] = (d, ff)
        shapes[base + "feed_forward/intermediate_dense/bias"] = (ff,)
        shapes[base + "feed_forward/output_dense/kernel"] = (ff, d)
        shapes[base + "feed_forward/output_dense/bias"] = (d,)
        shapes[base + "final_layer_norm/gamma"] = (d,)
        shapes[base + "final_layer_norm/beta"] = (d,)
    if with_head:
        shapes["lm_head/kernel"] = (d, cfg.vocab_size)
        shapes["lm_head/bias"] = (cfg.vocab_size,)
    return shapes


def random_params(cfg, seed=0, with_head=True) -> Params:
    """Seeded random weights at a realistic scale: fan-in scaled kernels, norm gains
    1 + 0.1 N(0,1) and norm/bias offsets 0.1 N(0,1) so that layout or ordering bugs show."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in param_shapes(cfg, with_head).items():
        if name.endswith("gamma"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("beta") or name.endswith("bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("weight_g"):
            t = 1.0 + 0.5 * torch.rand(shape, generator=g)
        elif name.endswith("masked_spec_embed"):
            t = torch.rand(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[:-1]))
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
            if "conv_layers" in name:
                t = t * math.sqrt(2.0)  # keep activations O(1) through GELU stacks
        out[name] = t.float()
    return out
