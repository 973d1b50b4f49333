"""SpecAugment time masking (training only), mirroring src/wav2vec2/spec_augment.py:43-128.

Host-side logic: like the reference, span starts are drawn with the *numpy* global RNG (the
reference samples ``np.random.uniform`` at trace time, spec_augment.py:13-15, and the span count
with ``np.random.rand``, :53).  Only the final ``where(mask, masked_spec_embed, features)``
touches device memory.
"""
import numpy as np
import torch


def _compute_mask_indices(shape, mask_prob, mask_length, min_masks=2):
    """Returns a {0,1} int64 array [batch, seqlen]; same algorithm as spec_augment.py:43-90."""
    batch_size, seqlen = shape
    if mask_length > seqlen:
        raise ValueError(f"`mask_length` ({mask_length}) must be smaller than `seq_length` ({seqlen}).")
    num_spans = max(int(mask_prob * (seqlen / mask_length) + float(np.random.rand(1)[0])), min_masks)
    if num_spans * mask_length > seqlen:
        num_spans = seqlen // mask_length
    # gumbel top-k over a uniform distribution == sampling start indices without replacement
    noise = np.random.uniform(0, 1, (batch_size, seqlen - (mask_length - 1)))
    scores = 1.0 - np.log(noise)
    starts = np.argsort(-scores, axis=-1, kind="stable")[:, :num_spans]
    idx = (starts[:, :, None] + np.arange(mask_length)[None, None, :]).reshape(batch_size, -1)
    mask = np.zeros((batch_size, seqlen), dtype=np.int64)
    np.put_along_axis(mask, idx, 1, axis=-1)
    return mask


def apply_spec_augmentation(features, masked_spec_augment, mask_prob, mask_length):
    """features [B,T,D]; frames inside a sampled span are replaced by ``masked_spec_augment`` [D]."""
    mask = _compute_mask_indices(tuple(features.shape[:2]), mask_prob, mask_length, min_masks=2)
    mask = torch.from_numpy(mask).to(features.device).bool()[:, :, None]
    return torch.where(mask, masked_spec_augment.to(features.dtype)[None, None, :], features)
