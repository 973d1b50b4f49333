"""Hyper-parameter containers for the B200 Wav2Vec2 path.

The reference keeps its hyper-parameters in a dataclass (src/wav2vec2/config.py:6-38 for the base model, :63-73 for
the "robust" 24-layer variant).  For the drop-in to work, names, order and defaults have to be the same here -
``Wav2Vec2Config(**json.load(f))`` written by either implementation must load in the other, and the historical
spelling ``kernal_sizes`` is part of that surface.  The fields are declared ONCE in the table below, each with what
it controls on the CUDA path, and the dataclass is generated from that table.

On top of the reference surface this module adds the shape arithmetic the kernels need (``conv_frames``,
``num_frames``, ``head_size``), which the reference spreads over modeling.py:203-204 and losses.py:47-56.
"""
import dataclasses
import json
import os
from typing import Any, List

# (name, default, what it controls).  Mutable defaults are given as tuples and become fresh lists per instance.
_FIELD_TABLE = (
    ("vocab_size", 32, "columns of lm_head / CTC alphabet (modeling.py:231)"),
    ("dropout", 0.1, "rate of all six Dropout sites, training only (annotated int upstream, the value is a float)"),
    ("hidden_size", 768, "encoder width d; every Dense / LayerNorm of the encoder"),
    ("num_heads", 12, "attention heads; the sm_100a attention kernels need hidden_size / num_heads == 64"),
    ("num_layers", 12, "transformer layers"),
    ("intermediate_size", 3072, "FFN width"),
    ("is_gelu_approx", False, "tf.nn.gelu(approximate=...) switch of the reference: False = erf form, True = tanh form"),
    ("layer_norm_eps", 1e-5, "epsilon of every LayerNormalization"),
    ("survival_prob", 1.0, "StochasticDepth on the FFN branch, training only (tensorflow_addons.py:374-394)"),
    ("pad_id", 0, "CTC blank / label padding (losses.py:32-41)"),
    ("num_conv_pos_embeddings", 128, "taps of the positional convolution (encoder.py:177-181)"),
    ("num_conv_pos_embedding_groups", 16, "groups of the positional convolution"),
    ("filter_sizes", (512,) * 7, "output channels of the 7 extractor convs (feature_extractor.py:27-37)"),
    ("kernal_sizes", (10, 3, 3, 3, 3, 2, 2), "their kernel widths (sic)"),
    ("strides", (5, 2, 2, 2, 2, 2, 2), "their strides: 246000 samples -> 768 frames"),
    ("conv_bias", False, "bias on the extractor convs (robust: True)"),
    ("apply_spec_augment", True, "SpecAugment time masking, training only (spec_augment.py:93-128)"),
    ("mask_time_prob", 0.05, "SpecAugment: fraction of frames that start a span"),
    ("mask_time_length", 10, "SpecAugment: span length in frames"),
    ("attention_norm_type", "postnorm", "'postnorm' (base) or 'prenorm' (robust) transformer layers (encoder.py:111-134)"),
    ("feature_extractor_norm_type", "group", "'group' = GroupNorm on conv 0 only (base), 'layer' = LayerNorm on all 7 (robust)"),
    ("is_robust", False, "robust checkpoints expect an attention mask (modeling.py:183-186)"),
)


def _dataclass_fields(table):
    out = []
    for name, default, _doc in table:
        if isinstance(default, tuple):
            out.append((name, list, dataclasses.field(default_factory=lambda d=default: list(d))))
        else:
            out.append((name, type(default), dataclasses.field(default=default)))
    return out


_ConfigFields = dataclasses.make_dataclass("_ConfigFields", _dataclass_fields(_FIELD_TABLE))


@dataclasses.dataclass
class Wav2Vec2Config(_ConfigFields):
    """``Wav2Vec2Config(**fields)``: the reference's constructor surface (config.py:7-38), its validation and persistence."""

    FIELD_DOCS = {name: doc for name, _d, doc in _FIELD_TABLE}

    def __post_init__(self):
        # failure modes of the reference (config.py:40-49): ValueError for structural mismatches, AssertionError for
        # unknown norm switches
        widths = {len(self.filter_sizes), len(self.kernal_sizes), len(self.strides)}
        if len(widths) != 1:
            raise ValueError("Length of filter_sizes, kernal_sizes, strides must match.")
        if self.hidden_size % self.num_heads != 0:
            raise ValueError("Hidden size must be perfect multiple of num_heads.")
        assert self.feature_extractor_norm_type in ("group", "layer"), "Only `group` / `layer` are supported"
        assert self.attention_norm_type in ("prenorm", "postnorm"), "Only `prenorm` / `postnorm` are supported"

    # ---- persistence (config.py:51-60): <dir>/config.json
    def to_dict(self) -> dict:
        return dataclasses.asdict(self)

    def save_pretrained(self, save_dir):
        os.makedirs(save_dir, exist_ok=True)
        with open(os.path.join(save_dir, "config.json"), "w") as fh:
            json.dump(self.to_dict(), fh)

    @classmethod
    def from_json(cls, path: str):
        with open(path, "r") as fh:
            return cls(**json.load(fh))

    # ---- shape arithmetic used by the kernels, the loss and the mask derivation
    def conv_frames(self, num_samples: int) -> List[int]:
        """Frames after each VALID conv: ``T <- 1 + (T - k) // s`` (modeling.py:203-204, losses.py:47-56)."""
        frames, t = [], int(num_samples)
        for k, s in zip(self.kernal_sizes, self.strides):
            t = 1 + (t - k) // s
            frames.append(t)
        return frames

    def num_frames(self, num_samples: int) -> int:
        return self.conv_frames(num_samples)[-1]

    @property
    def head_size(self) -> int:
        return self.hidden_size // self.num_heads


def _with_defaults(base, name, doc, **overrides: Any):
    """A subclass of ``base`` whose listed fields have other defaults (still a dataclass, still ``isinstance`` of base)."""
    fields = [(k, type(v), dataclasses.field(default=v)) for k, v in overrides.items()]
    cls = dataclasses.make_dataclass(name, fields, bases=(base,))
    cls.__doc__ = doc
    cls.__module__ = __name__
    return cls


# wav2vec2-large-robust / xlsr-53 (config.py:63-73): pre-norm encoder, LayerNorm + bias on every extractor conv, 24 x 1024
RobustWav2Vec2Config = _with_defaults(
    Wav2Vec2Config, "RobustWav2Vec2Config",
    "wav2vec2-large-robust / xlsr-53: pre-norm encoder, LayerNorm convs with bias, 24 layers of width 1024.",
    attention_norm_type="prenorm", feature_extractor_norm_type="layer", is_robust=True, conv_bias=True,
    hidden_size=1024, intermediate_size=4096, num_heads=16, num_layers=24)
