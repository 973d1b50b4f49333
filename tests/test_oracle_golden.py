"""Pins the CPU oracle against the golden vectors (CPU only).  See tests/golden/make_golden.py for provenance."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import w2v2_oracle as O
from wav2vec2.config import RobustWav2Vec2Config, Wav2Vec2Config

G = os.path.join(os.path.dirname(__file__), "golden")
SMALL = dict(hidden_size=128, num_heads=2, num_layers=2, intermediate_size=256, num_conv_pos_embedding_groups=2)


def _case(name):
    z = np.load(os.path.join(G, f"{name}.npz"))
    cfg = (RobustWav2Vec2Config if name.startswith("robust") else Wav2Vec2Config)(**SMALL)
    am = torch.from_numpy(z["attention_mask"]) if z["attention_mask"].size else None
    return cfg, int(z["seed"]), torch.from_numpy(z["speech"]), am, torch.from_numpy(z["logits"]), torch.from_numpy(z["hidden"])


@pytest.mark.parametrize("name", ["base_small", "robust_small"])
def test_oracle_matches_hf_golden(name):
    """The reference's tests equate TF to HF at atol 1e-3 (tests/test_wav2vec2.py:77-79); the restatement is far tighter."""
    cfg, seed, x, am, logits, hidden = _case(name)
    params = O.random_params(cfg, seed=seed)
    got_h = O.wav2vec2_model(x, params, cfg, attention_mask=am)
    got_l = O.wav2vec2_for_ctc(x, params, cfg, attention_mask=am)
    assert got_l.shape == logits.shape and got_h.shape == hidden.shape
    assert (got_h - hidden).abs().max().item() < 2e-5
    assert (got_l - logits).abs().max().item() < 2e-5


def test_processor_known_answer():
    """Known-answer vector held by the reference itself: tests/test_dataloader.py:56-63."""
    z = np.load(os.path.join(G, "processor.npz"))
    wav = O.read_wav_s16(os.path.join(G, "sample.wav"))
    assert len(wav) == 46797 == int(z["num_samples"])
    norm = O.normalize_utterance(wav[None, :])
    assert np.allclose(norm[32:40], z["reference_vector"])          # same default tolerances as the reference test
    assert np.abs(norm[32:40] - z["reference_vector"]).max() < 1e-7
    # quiet recording: var(x) ~ 7e-5, so the reference's eps = 1e-5 leaves var(normalised) = 0.878, not 1
    assert abs(float(norm.mean())) < 1e-5 and abs(float(norm.var()) - 0.8779) < 1e-3


def test_ctc_known_answer():
    """CTC NLL for the reference's test labels (np.random.seed(0); randint(1,30,(2,24)), tests/test_wav2vec2.py:41-42)."""
    z = np.load(os.path.join(G, "ctc.npz"))
    cfg = Wav2Vec2Config()
    np.random.seed(0)
    assert np.array_equal(np.random.randint(1, 30, size=(2, 24)), z["labels"])
    total, per = O.ctc_loss(torch.from_numpy(z["labels"]), torch.from_numpy(z["logits"]), cfg, division_factor=1.0)
    assert np.allclose(per, z["loss_per_sample"], atol=1e-6 * 500)
    assert abs(total - z["loss_per_sample"].sum()) < 1e-3          # tests/test_wav2vec2.py:235-237 tolerance
    total2, grad = O.ctc_loss_and_grad(torch.from_numpy(z["labels"]), torch.from_numpy(z["logits"]), cfg)
    assert abs(total - total2) < 1e-6 * abs(total)
    assert grad.shape == z["logits"].shape


def test_weight_norm_conv_matches_torch():
    """The reference's only self-contained unit test (tests/test_wav2vec2.py:239-282): Conv1DWithWeightNorm(16, 3,
    padding=1, groups=2) == nn.utils.weight_norm(nn.Conv1d, dim=2), weights moved with a (2,1,0) transpose, atol 1e-4."""
    np.random.seed(0)
    array = np.random.uniform(size=(2, 128, 32)).astype(np.float32)
    conv = nn.Conv1d(32, 16, 3, padding=1, groups=2)
    conv = nn.utils.weight_norm(conv, dim=2)
    with torch.no_grad():
        want = conv(torch.from_numpy(array).transpose(2, 1)).transpose(2, 1)
    v = conv.weight_v.detach().permute(2, 1, 0)       # -> TF layout [k, cin/g, cout]
    g = conv.weight_g.detach().permute(2, 1, 0)       # -> [k, 1, 1]
    kernel = O.weight_norm_kernel(v, g)
    x = torch.nn.functional.pad(torch.from_numpy(array), (0, 0, 1, 1))
    got = O.conv1d_valid(x, kernel, conv.bias.detach(), stride=1, groups=2)
    assert np.allclose(want.numpy(), got.numpy(), atol=1e-4)


def test_group_norm_is_per_channel_over_time():
    """tensorflow_addons.py:207-231 with groups == channels: biased statistics over time per (b, c)."""
    torch.manual_seed(0)
    x = torch.randn(2, 50, 8) * 3 + 1
    y = O.group_norm_per_channel(x, torch.ones(8), torch.zeros(8), 1e-5)
    want = torch.nn.functional.group_norm(x.transpose(1, 2), 8, eps=1e-5).transpose(1, 2)
    assert torch.allclose(y, want, atol=1e-5)


def test_frame_lengths_formula():
    cfg = Wav2Vec2Config()
    assert cfg.conv_frames(246000) == [49199, 24599, 12299, 6149, 3074, 1537, 768]     # src/main.py:48-50
    assert cfg.conv_frames(46797)[-1] == 145 and cfg.conv_frames(16000)[-1] == 49
    n = O.frame_lengths(cfg, torch.tensor([246000, 245000, 16000]))
    assert n.tolist() == [768, 765, 49]


def test_oracle_training_masks_are_identity_when_all_kept():
    """The oracle's explicit Dropout / StochasticDepth masks (checker of the stage-2 train step): all-ones masks reproduce
    the eval forward bit for bit, a dropped FFN branch leaves its parameters without gradient, and a dropped unit of the
    `head` mask removes exactly that unit's contribution."""
    import torch
    from oracle import w2v2_oracle as O
    from wav2vec2.config import Wav2Vec2Config
    cfg = Wav2Vec2Config(num_layers=2)
    p = O.random_params(cfg, seed=3)
    x = torch.randn(2, 4000, generator=torch.Generator().manual_seed(0))
    ref = O.wav2vec2_for_ctc(x, p, cfg)
    B, T, d, ffn, H = 2, ref.shape[1], cfg.hidden_size, cfg.intermediate_size, cfg.num_heads
    ones = {"proj": torch.ones(B, T, d), "enc": torch.ones(B, T, d), "head": torch.ones(B, T, d)}
    for i in range(cfg.num_layers):
        ones.update({f"attn_probs.{i}": torch.ones(B, H, T, T), f"attn_out.{i}": torch.ones(B, T, d),
                     f"ffn_mid.{i}": torch.ones(B, T, ffn), f"stochastic_depth.{i}": torch.ones(())})
    assert torch.equal(O.wav2vec2_for_ctc(x, p, cfg, drop=ones), ref)
    q = {k: t.clone().requires_grad_(True) for k, t in p.items()}
    drop = dict(ones)
    drop["stochastic_depth.1"] = torch.zeros(())
    O.wav2vec2_for_ctc(x, q, cfg, drop=drop).sum().backward()
    assert q["wav2vec2/encoder/layers/1/feed_forward/output_dense/kernel"].grad.abs().max() == 0
    assert q["wav2vec2/encoder/layers/0/feed_forward/output_dense/kernel"].grad.abs().max() > 0
    head = dict(ones)
    head["head"] = torch.ones(B, T, d)
    head["head"][:, :, 5] = 0.0
    got = O.wav2vec2_for_ctc(x, p, cfg, drop=head)
    hidden = O.wav2vec2_model(x, p, cfg)
    assert torch.allclose(ref - got, hidden[:, :, 5:6] * p["lm_head/kernel"][5][None, None, :], atol=1e-5)
