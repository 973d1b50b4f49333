"""End-to-end GPU parity: the B200 model behind the reference surface vs the CPU fp32 oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import w2v2_oracle as O                                     # noqa: E402 (checker only)
from wav2vec2 import RobustWav2Vec2Config, Wav2Vec2Config, Wav2Vec2ForCTC, Wav2Vec2Model   # noqa: E402


def _build(cls, cfg, precision, seed=1):
    params = O.random_params(cfg, seed=seed, with_head=cls is Wav2Vec2ForCTC)
    m = cls(cfg, input_shape=(1, 2048), precision=precision)
    m.set_variables(params)
    return m, params


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-3), ("bf16", 1e-1)])
def test_base_ctc_logits_match_oracle(precision, tol):
    """tests/test_wav2vec2.py:_test_inference / test_end2end analogue: logits vs oracle (atol 1e-3 in parity mode)."""
    cfg = Wav2Vec2Config(num_layers=3)
    m, params = _build(Wav2Vec2ForCTC, cfg, precision)
    torch.manual_seed(0)
    x = torch.randn(2, 20000)
    got = m(x.cuda(), training=False).cpu()
    ref = O.wav2vec2_for_ctc(x, params, cfg)
    err = (got - ref).abs().max().item()
    print(f"base/{precision}: logits max-abs err {err:.3e} (max |logit| {ref.abs().max():.2f})")
    assert got.shape == ref.shape
    assert err < tol


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-3), ("bf16", 1e-1)])
def test_robust_hidden_states_with_mask(precision, tol):
    """test_wav2vec2_robust analogue (tests/test_wav2vec2.py:58-62,85-87): prenorm, layer-norm convs, mask."""
    cfg = RobustWav2Vec2Config(num_layers=2)
    m, params = _build(Wav2Vec2Model, cfg, precision)
    torch.manual_seed(0)
    x = torch.randn(2, 20000)
    am = torch.ones(2, 20000, dtype=torch.int32)
    am[0, -1000:] = 0
    am[1, -132:] = 0
    got = m(x.cuda(), attention_mask=am.cuda(), training=False).cpu()
    ref = O.wav2vec2_model(x, params, cfg, attention_mask=am)
    err = (got - ref).abs().max().item()
    print(f"robust/{precision}: hidden max-abs err {err:.3e} (max |h| {ref.abs().max():.2f})")
    assert err < tol


def test_sample_wav_full_depth():
    """BASELINE config 1: base model, data/sample.wav (normalised), batch 1, all 12 layers, parity mode."""
    import os
    cfg = Wav2Vec2Config()
    m, params = _build(Wav2Vec2ForCTC, cfg, "bf16x3", seed=3)
    wav = O.read_wav_s16(os.path.join(os.path.dirname(__file__), "golden", "sample.wav"))
    x = torch.from_numpy(O.normalize_utterance(wav[None, :]))[None, :]
    got = m(x.cuda()).cpu()
    ref = O.wav2vec2_for_ctc(x, params, cfg)
    assert got.shape == (1, 145, 32)
    err = (got - ref).abs().max().item()
    print(f"sample.wav 12 layers: logits max-abs err {err:.3e}")
    assert err < 1e-3
    assert torch.equal(got.argmax(-1), ref.argmax(-1))        # greedy CTC path identical (test_wav2vec2.py:165-170)
