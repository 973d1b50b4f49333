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


@pytest.mark.parametrize("name", ["base_small", "robust_small"])
def test_golden_fixture_logits(name):
    """CUDA path vs the committed golden vectors (HF outputs; tests/golden/make_golden.py), parity mode, atol 1e-3."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}.npz"))
    small = dict(hidden_size=128, num_heads=2, num_layers=2, intermediate_size=256, num_conv_pos_embedding_groups=2)
    cfg = (RobustWav2Vec2Config if name.startswith("robust") else Wav2Vec2Config)(**small)
    m = Wav2Vec2ForCTC(cfg, precision="bf16x3")
    m.set_variables(O.random_params(cfg, seed=int(z["seed"])))
    am = torch.from_numpy(z["attention_mask"]).cuda() if z["attention_mask"].size else None
    got = m(torch.from_numpy(z["speech"]).cuda(), attention_mask=am).cpu().numpy()
    err = np.abs(got - z["logits"]).max()
    print(f"{name}: max-abs err vs golden {err:.3e}")
    assert np.allclose(got, z["logits"], atol=1e-3)


def test_ctc_loss_matches_golden_and_reference_tolerance():
    """CTCLoss surface (losses.py) on the reference's test labels; loss atol 1e-3 (tests/test_wav2vec2.py:235-237)."""
    import os
    from wav2vec2 import CTCLoss
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ctc.npz"))
    cfg = Wav2Vec2Config()
    loss_fn = CTCLoss(cfg, (2, 46797), division_factor=1)          # 46797 samples -> 145 frames
    loss = loss_fn(torch.from_numpy(z["labels"]).int().cuda(), torch.from_numpy(z["logits"]).cuda())
    assert abs(loss.item() - float(z["loss_per_sample"].sum())) < 1e-3 * 10   # fp32 kernel on an NLL of ~1000


def test_batch_sharding_gives_identical_logits():
    """Multi-GPU contract: utterances are independent, so any batch split reproduces the same logits."""
    from wav2vec2 import parallel
    cfg = Wav2Vec2Config(num_layers=2)
    m, _ = _build(Wav2Vec2ForCTC, cfg, "bf16")
    x = torch.randn(4, 16000, generator=torch.Generator().manual_seed(3)).cuda()
    full = m(x)
    parts = torch.cat([m(parallel.shard_batch(x, r, 2)) for r in range(2)])
    assert torch.equal(full, parts)


def test_stage1_train_step_matches_torch_autograd():
    """src/main.py:210-223 (head-only fine-tuning): CTC loss, lm_head gradients and the Keras-Adam update of the CUDA step
    vs torch autograd + torch.optim.Adam(eps=1e-7) on the same hidden states."""
    from wav2vec2 import CTCLoss
    from wav2vec2.training import Stage1Trainer
    cfg = Wav2Vec2Config(num_layers=2, dropout=0.0, apply_spec_augment=False)
    m, params = _build(Wav2Vec2ForCTC, cfg, "bf16x3", seed=4)
    B, L = 3, 16000
    x = torch.randn(B, L, generator=torch.Generator().manual_seed(1))
    np.random.seed(0)
    labels = torch.from_numpy(np.random.randint(1, 30, size=(B, 12))).int()
    labels[1, 7:] = 0
    trainer = Stage1Trainer(m, CTCLoss(cfg, (B, L), division_factor=B), learning_rate=1e-3)
    _, hidden = m.forward_with_hidden(x.cuda(), training=True)
    h = hidden.clone().cpu().double()
    W0 = params["lm_head/kernel"].double().clone().requires_grad_(True)
    b0 = params["lm_head/bias"].double().clone().requires_grad_(True)
    loss = trainer.step(x.cuda(), labels.cuda())
    T = cfg.num_frames(L)
    logits = (h @ W0 + b0).view(B, T, -1)
    lp = torch.log_softmax(logits, -1).transpose(0, 1)
    lens = (labels != 0).sum(-1)
    ref_loss = torch.nn.functional.ctc_loss(lp, labels.long(), torch.full((B,), T), lens, blank=0, reduction="sum") / B
    ref_loss.backward()
    gW, gb = W0.grad.clone(), b0.grad.clone()

    def keras_adam(w, g, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7, t=1):   # Keras (non-amsgrad) update, first step
        m, v = (1 - b1) * g, (1 - b2) * g * g
        lr_t = lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
        return w.detach() - lr_t * m / (v.sqrt() + eps)
    W1, b1_ = keras_adam(W0, gW), keras_adam(b0, gb)
    assert abs(loss.item() - ref_loss.item()) < 1e-3 * max(1.0, abs(ref_loss.item()))
    got_gW = trainer.flat_g[: trainer.d * trainer.V].view(trainer.d, trainer.V).cpu().double()
    assert (got_gW - gW).abs().max().item() < 1e-4 * max(1.0, gW.abs().max().item())
    assert (trainer.flat_g[trainer.d * trainer.V:].cpu().double() - gb).abs().max().item() < 1e-4 * max(1.0, gb.abs().max().item())
    # the first Adam step is ~ lr * sign(g): only entries whose gradient is well above fp32 noise are comparable
    big = gW.abs() > 1e-3 * gW.abs().max()
    dW = (m.variables["lm_head/kernel"].cpu().double() - W1).abs()
    print(f"stage-1 step: loss {loss.item():.4f} vs {ref_loss.item():.4f}; max |dW - ref| on significant entries {dW[big].max():.2e}")
    assert dW[big].max().item() < 2e-5 and dW.max().item() < 2.1e-3
    assert (m.variables["lm_head/bias"].cpu().double() - b1_).abs().max().item() < 2e-5
    # the packed copy follows the update: a second forward uses the new head
    logits2 = m(x.cuda())
    Wn, bn = m.variables["lm_head/kernel"].cpu().double(), m.variables["lm_head/bias"].cpu().double()
    want = (h @ Wn + bn).view(B, T, -1).float()
    assert (logits2.cpu() - want).abs().max().item() < 1e-3


def test_cuda_graph_replay_is_bit_identical():
    """The whole forward replayed as one CUDA graph gives the same logits as the eager launch sequence, also with a mask."""
    cfg = RobustWav2Vec2Config(num_layers=2)
    m, _ = _build(Wav2Vec2ForCTC, cfg, "bf16")
    g = torch.Generator().manual_seed(7)
    x1, x2 = torch.randn(2, 16000, generator=g).cuda(), torch.randn(2, 16000, generator=g).cuda()
    am = torch.ones(2, 16000, dtype=torch.int32, device="cuda")
    am[1, -3000:] = 0
    eager1, eager2 = m(x1, attention_mask=am), m(x2, attention_mask=am)
    m.enable_cuda_graph(True)
    assert torch.equal(m(x1, attention_mask=am), eager1)
    assert torch.equal(m(x2, attention_mask=am), eager2)      # replay with new data in the static input buffer
    assert torch.equal(m(x1, attention_mask=am), eager1)
    assert len(m._graphs) == 1


@pytest.mark.parametrize("L,B", [(400, 1), (719, 3), (1039, 5), (5000, 2)])
def test_tiny_and_ragged_inputs_match_oracle(L, B):
    """Edge shapes: the shortest waveform the extractor accepts (400 samples -> ONE frame), 1 / 2 / 3 frames, odd batch sizes,
    every kernel running a single ragged tile."""
    cfg = Wav2Vec2Config(num_layers=2)
    m, params = _build(Wav2Vec2ForCTC, cfg, "bf16x3")
    x = torch.randn(B, L, generator=torch.Generator().manual_seed(L))
    got = m(x.cuda()).cpu()
    ref = O.wav2vec2_for_ctc(x, params, cfg)
    assert got.shape == ref.shape == (B, cfg.num_frames(L), cfg.vocab_size)
    err = (got - ref).abs().max().item()
    print(f"L={L} B={B}: {ref.shape[1]} frames, logits max-abs err {err:.3e}")
    assert err < 1e-3


def test_too_short_input_raises():
    cfg = Wav2Vec2Config(num_layers=1)
    m, _ = _build(Wav2Vec2ForCTC, cfg, "bf16")
    with pytest.raises(ValueError):
        m(torch.randn(1, 399).cuda())


def test_cuda_graphs_follow_weight_updates_and_shape_changes():
    """a captured graph must never replay against replaced weights or a reallocated arena.
    (1) set_variables after capture -> the next call equals a fresh eager model with the new weights;
    (2) shape A graph, shape B call, shape A again -> still equals eager A (every graph owns its arena)."""
    cfg = Wav2Vec2Config(num_layers=2)
    m, params = _build(Wav2Vec2ForCTC, cfg, "bf16")
    g = torch.Generator().manual_seed(11)
    xa, xb = torch.randn(2, 16000, generator=g).cuda(), torch.randn(3, 9000, generator=g).cuda()
    ea, eb = m(xa), m(xb)
    m.enable_cuda_graph(True)
    assert torch.equal(m(xa), ea)
    assert torch.equal(m(xb), eb)
    keep = m(xa)                                      # a user-held result while other shapes run
    assert torch.equal(m(xb), eb) and torch.equal(m(xa), ea) and torch.equal(keep, ea)
    params2 = O.random_params(cfg, seed=9)
    m.set_variables(params2)
    assert len(m._graphs) == 0                        # the captured graphs pointed at the old packed weights
    fresh = Wav2Vec2ForCTC(cfg, precision="bf16")
    fresh.set_variables(params2)
    assert torch.equal(m(xa), fresh(xa))


@pytest.mark.parametrize("arch", ["base", "robust"])
def test_tf_approximate_gelu_switch(arch):
    """config.is_gelu_approx=True (config.py:14): tf.nn.gelu(approximate=True) at the extractor, positional-conv and FFN
    sites (feature_extractor.py:58, encoder.py:127,181) - parity mode vs the oracle with the same switch."""
    kw = dict(num_layers=2, is_gelu_approx=True)
    cfg = Wav2Vec2Config(**kw) if arch == "base" else RobustWav2Vec2Config(**kw)
    m, params = _build(Wav2Vec2ForCTC, cfg, "bf16x3")
    x = torch.randn(2, 12000, generator=torch.Generator().manual_seed(5))
    got = m(x.cuda()).cpu()
    ref = O.wav2vec2_for_ctc(x, params, cfg)
    exact = O.wav2vec2_for_ctc(x, params, Wav2Vec2Config(num_layers=2) if arch == "base" else RobustWav2Vec2Config(num_layers=2))
    err = (got - ref).abs().max().item()
    print(f"{arch} is_gelu_approx: max-abs err {err:.3e}; approx-vs-erf oracle difference {(ref - exact).abs().max():.3e}")
    assert err < 1e-3
    fast = Wav2Vec2ForCTC(cfg, precision="bf16")
    fast.set_variables(params)
    assert (fast(x.cuda()).cpu() - ref).abs().max().item() < 1e-1


def test_utterance_without_valid_frames_stays_finite():
    """An attention mask shorter than the receptive field gives ZERO valid frames: the reference's additive -10000 on every key
    cancels in the softmax (encoder.py:256-263), so it attends over all keys and stays finite - no NaN from exp(-inf + inf)."""
    cfg = RobustWav2Vec2Config(num_layers=2)
    m, params = _build(Wav2Vec2ForCTC, cfg, "bf16x3")
    x = torch.randn(2, 8000, generator=torch.Generator().manual_seed(6))
    am = torch.ones(2, 8000, dtype=torch.int32)
    am[1, 300:] = 0                                   # 300 samples < 400: no frame survives
    x = x * am
    got = m(x.cuda(), attention_mask=am.cuda()).cpu()
    ref = O.wav2vec2_for_ctc(x, params, cfg, attention_mask=am)
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() < 1e-3


def test_training_call_applies_spec_augment_in_the_projection_epilogue():
    """model(x, training=True) with dropout 0: the SpecAugment replacement (modeling.py:193-199) happens inside the projection
    GEMM; with the mask the call sampled (same numpy seed) the oracle gives the same hidden states, also with an attention mask."""
    from wav2vec2.spec_augment import _compute_mask_indices
    cfg = RobustWav2Vec2Config(num_layers=2, dropout=0.0, apply_spec_augment=True)
    m, params = _build(Wav2Vec2Model, cfg, "bf16x3")
    x = torch.randn(2, 20000, generator=torch.Generator().manual_seed(8))
    am = torch.ones(2, 20000, dtype=torch.int32)
    am[0, -4000:] = 0
    T = cfg.num_frames(20000)
    np.random.seed(123)
    got = m(x.cuda(), attention_mask=am.cuda(), training=True).cpu()
    np.random.seed(123)
    mask = _compute_mask_indices((2, T), cfg.mask_time_prob, cfg.mask_time_length, min_masks=2)
    ref = O.wav2vec2_model(x, params, cfg, attention_mask=am, spec_mask=torch.from_numpy(mask).bool())
    assert mask.sum() > 0
    assert (got - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("arch", ["base", "robust"])
@pytest.mark.parametrize("precision", ["bf16", "fp16", "fp16f8"])
def test_layernorm_fold_equals_the_unfolded_path(arch, precision, monkeypatch):
    """The LayerNorm fold (gamma into the next Dense, mean / rstd in its epilogue, no stand-alone LayerNorm pass) against the
    same model with W2V2_LN_FOLD=0, and against the oracle: the fold must not cost accuracy in any precision mode."""
    cls = Wav2Vec2Config if arch == "base" else RobustWav2Vec2Config
    cfg = cls(num_layers=3)
    params = O.random_params(cfg, seed=2)
    # LayerNorm inputs with a mean well away from zero in some rows: the fold subtracts rstd * mean * colsum explicitly
    x = torch.randn(2, 20000, generator=torch.Generator().manual_seed(4)) + 0.3
    ref = O.wav2vec2_for_ctc(x, params, cfg)
    errs = {}
    for fold in ("1", "qkv", "0"):          # both LayerNorms of a layer / only the one in front of q, k, v / none
        monkeypatch.setenv("W2V2_LN_FOLD", fold)
        m = Wav2Vec2ForCTC(cfg, precision=precision)
        m.set_variables(params)
        errs[fold] = (m(x.cuda()).cpu() - ref).abs().max().item()
        assert m._fold == (fold != "0") and m._fold_ff1 == (fold == "1")
    print(f"{arch}/{precision}: logits max-abs err folded {errs['1']:.3e}, q/k/v only {errs['qkv']:.3e}, unfolded {errs['0']:.3e}")
    tol = {"bf16": 1e-1, "fp16": 6e-3, "fp16f8": 1e-3}[precision]
    for mode in ("1", "qkv"):
        assert errs[mode] < tol and errs[mode] < 2.0 * errs["0"] + 1e-4
