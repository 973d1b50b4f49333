"""Hyper-parameter containers for the B200 Wav2Vec2 path.

Field names, order and defaults follow the reference configuration dataclass
(reference: src/wav2vec2/config.py:6-38 for the base model, :63-73 for the
"robust" 24-layer variant) so that ``Wav2Vec2Config(**json.load(f))`` written by
either implementation loads in the other.  The historical spelling
``kernal_sizes`` is part of that surface and is kept on purpose.

On top of the reference surface this module adds the shape arithmetic the CUDA
path needs (``conv_frames``), which the reference spreads over
modeling.py:203-204 and losses.py:47-56.
"""
import dataclasses
import json
import os
from dataclasses import dataclass, field
from typing import List

_NORM_KINDS_EXTRACTOR = ("group", "layer")
_NORM_KINDS_ENCODER = ("prenorm", "postnorm")


def _seven(value):
    return field(default_factory=lambda: list(value))


@dataclass
class Wav2Vec2Config:
    # -- CTC head / transformer encoder --------------------------------------
    vocab_size: int = 32
    dropout: int = 0.1  # (sic) annotated int in the reference, value is a float
    hidden_size: int = 768
    num_heads: int = 12
    num_layers: int = 12
    intermediate_size: int = 3072
    is_gelu_approx: bool = False
    layer_norm_eps: float = 1e-5
    survival_prob: float = 1.0
    pad_id: int = 0

    # -- positional convolution ------------------------------------------------
    num_conv_pos_embeddings: int = 128
    num_conv_pos_embedding_groups: int = 16

    # -- strided Conv1D feature extractor ------------------------------------
    filter_sizes: list = _seven([512] * 7)
    kernal_sizes: list = _seven([10, 3, 3, 3, 3, 2, 2])
    strides: list = _seven([5, 2, 2, 2, 2, 2, 2])
    conv_bias: bool = False

    # -- SpecAugment (training only) -----------------------------------------
    apply_spec_augment: bool = True
    mask_time_prob: float = 0.05
    mask_time_length: int = 10

    # -- architecture switches -------------------------------------------------
    attention_norm_type: str = "postnorm"
    feature_extractor_norm_type: bool = "group"  # (sic) str value, bool annotation upstream
    is_robust: bool = False

    def __post_init__(self):
        # Same failure modes as the reference (config.py:40-49): ValueError for
        # structural mismatches, AssertionError for unknown norm switches.
        n = len(self.filter_sizes)
        if len(self.kernal_sizes) != n or len(self.strides) != n:
            raise ValueError("Length of filter_sizes, kernal_sizes, strides must match.")
        if self.hidden_size % self.num_heads:
            raise ValueError("Hidden size must be perfect multiple of num_heads.")
        assert self.feature_extractor_norm_type in _NORM_KINDS_EXTRACTOR, \
            "Only `group` / `layer` are supported"
        assert self.attention_norm_type in _NORM_KINDS_ENCODER, \
            "Only `prenorm` / `postnorm` are supported"

    # -- persistence (config.py:51-60) ---------------------------------------
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

    # -- shape arithmetic used by kernels, loss and mask derivation -------------
    def conv_frames(self, num_samples: int) -> List[int]:
        """Frames after each VALID conv: ``T <- 1 + (T - k) // s`` (modeling.py:203-204)."""
        out, t = [], int(num_samples)
        for k, s in zip(self.kernal_sizes, self.strides):
            t = 1 + (t - k) // s
            out.append(t)
        return out

    def num_frames(self, num_samples: int) -> int:
        return self.conv_frames(num_samples)[-1]

    @property
    def head_size(self) -> int:
        return self.hidden_size // self.num_heads


@dataclass
class RobustWav2Vec2Config(Wav2Vec2Config):
    """wav2vec2-large-robust / xlsr-53: pre-norm encoder, LayerNorm convs with bias."""
    attention_norm_type: str = "prenorm"
    feature_extractor_norm_type: str = "layer"
    is_robust: bool = True
    conv_bias: bool = True

    hidden_size: int = 1024
    intermediate_size: int = 4096
    num_heads: int = 16
    num_layers: int = 24
