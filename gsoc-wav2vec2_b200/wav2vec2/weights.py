"""Checkpoint interchange: HF PyTorch ``state_dict`` <-> reference variable names/layouts.

The reference defines its weight-layout contract in src/convert_torch_to_tf.py:
key renaming (:13-18, special pos-conv keys :26-35) and transposes (:109-117:
pos-conv ``weight_g/weight_v`` by (2,1,0); every ``kernel`` fully transposed).
This module applies the same rules to plain tensors so a HF checkpoint (the
format real weights ship in) loads into the B200 model, and inverts them for
export.  Newer ``transformers`` name the weight-norm pair
``parametrizations.weight.original0/1`` (= g / v); both spellings are accepted.

Variable names are the reference's Keras names without the ``:0`` suffix and
without the outer ``wav2vec2-ctc/`` scope: ``wav2vec2/...`` and ``lm_head/...``.
"""
from typing import Dict

import torch

_POS_G = ("weight_g", "parametrizations.weight.original0")
_POS_V = ("weight_v", "parametrizations.weight.original1")


def _rename(hf_key: str) -> str:
    k = hf_key
    if not (k.startswith("wav2vec2.") or k.startswith("lm_head.")):
        k = "wav2vec2." + k           # bare Wav2Vec2Model checkpoints (no head)
    for g in _POS_G:
        k = k.replace("pos_conv_embed.conv." + g, "pos_conv_embed.conv.weight_g")
    for v in _POS_V:
        k = k.replace("pos_conv_embed.conv." + v, "pos_conv_embed.conv.weight_v")
    if k.endswith("layer_norm.weight"):
        k = k[: -len("weight")] + "gamma"
    elif k.endswith("layer_norm.bias"):
        k = k[: -len("bias")] + "beta"
    elif k.endswith(".weight"):
        k = k[: -len("weight")] + "kernel"
    return k.replace(".", "/")


def hf_to_reference(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """HF tensors -> reference-layout tensors (fp32, contiguous)."""
    out = {}
    for key, t in state_dict.items():
        if "masked_spec_embed" not in key and t.ndim == 0:
            continue
        name = _rename(key)
        t = t.detach().to(torch.float32)
        if name.endswith("pos_conv_embed/conv/weight_g") or name.endswith("pos_conv_embed/conv/weight_v"):
            t = t.permute(2, 1, 0)
        elif name.endswith("/kernel"):
            t = t.permute(*reversed(range(t.ndim)))
        out[name] = t.contiguous()
    return out


def reference_to_hf(params: Dict[str, torch.Tensor], new_style_weight_norm=True) -> Dict[str, torch.Tensor]:
    """Inverse of :func:`hf_to_reference` (for exporting to ``transformers``)."""
    out = {}
    for name, t in params.items():
        key = name.replace("/", ".")
        t = t.detach().to(torch.float32)
        if key.endswith("pos_conv_embed.conv.weight_g") or key.endswith("pos_conv_embed.conv.weight_v"):
            t = t.permute(2, 1, 0)
            if new_style_weight_norm:
                tail = "original0" if key.endswith("weight_g") else "original1"
                key = key[: -len("weight_g")] + "parametrizations.weight." + tail
        elif key.endswith("layer_norm.gamma"):
            key = key[: -len("gamma")] + "weight"
        elif key.endswith("layer_norm.beta"):
            key = key[: -len("beta")] + "bias"
        elif key.endswith(".kernel"):
            key = key[: -len("kernel")] + "weight"
            t = t.permute(*reversed(range(t.ndim)))
        out[key] = t.contiguous()
    return out
