"""Host-side logic of the drop-in surface (CPU only): config, variable inventory, checkpoint mapping, processor,
spec-augment, mask derivation, persistence, error behaviour."""
import dataclasses
import logging
import os

import numpy as np
import pytest
import torch

from oracle import w2v2_oracle as O
from wav2vec2 import (CTCLoss, RobustWav2Vec2Config, Wav2Vec2Config, Wav2Vec2ForCTC, Wav2Vec2Model,
                      Wav2Vec2Processor)
from wav2vec2.modeling import variable_shapes
from wav2vec2.spec_augment import _compute_mask_indices, apply_spec_augmentation
from wav2vec2.weights import hf_to_reference, reference_to_hf

SMALL = dict(hidden_size=128, num_heads=2, num_layers=2, intermediate_size=256, num_conv_pos_embedding_groups=2)


def test_config_surface_matches_reference():
    cfg = Wav2Vec2Config()
    d = dataclasses.asdict(cfg)
    # field names (incl. the historical `kernal_sizes`) and defaults of src/wav2vec2/config.py:6-38
    assert list(d)[:10] == ["vocab_size", "dropout", "hidden_size", "num_heads", "num_layers", "intermediate_size",
                            "is_gelu_approx", "layer_norm_eps", "survival_prob", "pad_id"]
    assert d["kernal_sizes"] == [10, 3, 3, 3, 3, 2, 2] and d["strides"] == [5, 2, 2, 2, 2, 2, 2]
    assert (cfg.hidden_size, cfg.num_heads, cfg.num_layers, cfg.intermediate_size, cfg.vocab_size) == (768, 12, 12, 3072, 32)
    r = RobustWav2Vec2Config()
    assert (r.hidden_size, r.num_heads, r.num_layers, r.intermediate_size) == (1024, 16, 24, 4096)
    assert r.attention_norm_type == "prenorm" and r.feature_extractor_norm_type == "layer" and r.conv_bias and r.is_robust


def test_config_validation_errors():
    with pytest.raises(ValueError):
        Wav2Vec2Config(strides=[5, 2])
    with pytest.raises(ValueError):
        Wav2Vec2Config(hidden_size=770)
    with pytest.raises(AssertionError):
        Wav2Vec2Config(attention_norm_type="sandwich")
    with pytest.raises(AssertionError):
        Wav2Vec2Config(feature_extractor_norm_type="batch")


def test_config_json_roundtrip(tmp_path):
    cfg = Wav2Vec2Config(num_layers=3, dropout=0.0)
    cfg.save_pretrained(str(tmp_path))
    assert Wav2Vec2Config.from_json(os.path.join(tmp_path, "config.json")) == cfg


def test_variable_inventory_is_the_references():
    shapes = variable_shapes(Wav2Vec2Config(), with_head=True)
    assert len(shapes) == 213                                           # notebooks/wav2vec2_onnx.ipynb:125
    assert sum(int(np.prod(s)) for s in shapes.values()) == 94396320
    assert shapes == O.param_shapes(Wav2Vec2Config())
    assert shapes["wav2vec2/encoder/pos_conv_embed/conv/weight_v"] == (128, 48, 768)
    assert shapes["wav2vec2/feature_extractor/conv_layers/0/conv/kernel"] == (10, 1, 512)
    assert "wav2vec2/feature_extractor/conv_layers/1/layer_norm/gamma" not in shapes     # group norm: layer 0 only
    assert "wav2vec2/feature_extractor/conv_layers/6/conv/bias" in variable_shapes(RobustWav2Vec2Config(), True)


def test_checkpoint_mapping_roundtrip_both_weight_norm_spellings():
    cfg = Wav2Vec2Config(**SMALL)
    params = O.random_params(cfg, seed=0)
    for new_style in (True, False):
        hf = reference_to_hf(params, new_style_weight_norm=new_style)
        key = "wav2vec2.encoder.pos_conv_embed.conv." + ("parametrizations.weight.original1" if new_style else "weight_v")
        assert key in hf and hf[key].shape == (128, 64, 128)            # HF layout [cout, cin/g, k]
        assert hf["wav2vec2.encoder.layers.0.attention.q_proj.weight"].shape == (128, 128)
        back = hf_to_reference(hf)
        assert set(back) == set(params)
        assert all(torch.equal(back[k], params[k]) for k in params)


def test_model_constructor_and_errors():
    with pytest.raises(ValueError):
        Wav2Vec2Model({"hidden_size": 768})
    with pytest.raises(ValueError):
        Wav2Vec2ForCTC("not a config")
    Wav2Vec2Model(Wav2Vec2Config(is_gelu_approx=True, **SMALL), device="cpu")      # both GELU forms of config.py:14 are accepted
    with pytest.raises(ValueError):
        Wav2Vec2Model(Wav2Vec2Config(hidden_size=96, num_heads=2), device="cpu")   # head_size 48: not built
    m = Wav2Vec2ForCTC(Wav2Vec2Config(**SMALL), input_shape=(1, 2048), device="cpu")
    assert set(m.variables) == set(variable_shapes(m.config, True))
    g = m.variables["wav2vec2/encoder/pos_conv_embed/conv/weight_g"]
    v = m.variables["wav2vec2/encoder/pos_conv_embed/conv/weight_v"]
    assert torch.allclose(g, v.pow(2).sum((1, 2), keepdim=True).sqrt())   # tensorflow_addons.py:45-48
    m.freeze_feature_extractor()
    assert not m.trainable["wav2vec2/feature_extractor/conv_layers/3/conv/kernel"]
    assert m.trainable["wav2vec2/encoder/layers/0/attention/q_proj/kernel"]
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m(torch.zeros(1, 2048))


def test_save_and_load_pretrained(tmp_path):
    m = Wav2Vec2ForCTC(Wav2Vec2Config(**SMALL), device="cpu").init_random(3)
    m.save_pretrained(str(tmp_path))
    m2 = Wav2Vec2ForCTC.from_pretrained(str(tmp_path), device="cpu") if False else None
    from safetensors.torch import load_file
    sd = load_file(os.path.join(tmp_path, "model.safetensors"))
    assert set(sd) == set(m.variables) and all(torch.equal(sd[k], m.variables[k].cpu()) for k in sd)
    with pytest.raises(ValueError, match="Couldn't download"):
        Wav2Vec2ForCTC.from_pretrained(os.path.join(tmp_path, "missing"))


def test_mask_warnings_match_reference(caplog):
    m = Wav2Vec2Model(Wav2Vec2Config(**SMALL), device="cpu")
    with caplog.at_level(logging.WARNING):
        m._warn_mask(torch.ones(1, 10))
    assert "should not pass `attention_mask`" in caplog.text                # modeling.py:185-186
    r = Wav2Vec2Model(RobustWav2Vec2Config(**SMALL), device="cpu")
    caplog.clear()
    with caplog.at_level(logging.WARNING):
        r._warn_mask(None)
    assert "should pass `attention_mask`" in caplog.text                    # modeling.py:183-184


def test_frame_lengths_from_mask():
    m = Wav2Vec2Model(Wav2Vec2Config(**SMALL), device="cpu")
    am = torch.ones(3, 20000, dtype=torch.int32)
    am[0, -1000:] = 0
    am[1, -132:] = 0
    got = m._frame_lengths(am, 62)
    want = O.frame_lengths(m.config, am.sum(-1))
    assert got.tolist() == want.tolist() == [59, 61, 62]


def test_processor_tokenizer_and_decode():
    tok = Wav2Vec2Processor(is_tokenizer=True)
    ids = tok("She had your dark suit - in greasy wash water all year.")
    assert ids[:4] == [12, 11, 5, 4] and 4 in ids and 3 not in ids          # S H E | ; '-' -> space; no <unk>
    assert tok.decode(ids) == "SHE HAD YOUR DARK SUIT IN GREASY WASH WATER AL YEAR"    # repeats collapse (LL -> L, || -> |)
    assert tok.decode(ids, group_tokens=False) == "SHE HAD YOUR DARK SUIT   IN GREASY WASH WATER ALL YEAR"
    assert tok.decode([0, 0, 11, 11, 0, 5, 5, 15, 0, 15, 8, 4, 4, 0]) == "HELLO"
    vocab = tok.get_vocab()
    assert len(vocab) == 32 and vocab["<pad>"] == 0 and vocab["|"] == 4
    fe = Wav2Vec2Processor(is_tokenizer=False)
    x = torch.randn(1, 5000) * 3 + 2
    y = fe(x)
    assert y.shape == (5000,) and abs(float(y.mean())) < 1e-5 and abs(float(y.var(unbiased=False)) - 1) < 1e-3


def test_spec_augment_mask_properties():
    np.random.seed(0)
    m = _compute_mask_indices((4, 768), 0.05, 10, min_masks=2)
    assert m.shape == (4, 768) and set(np.unique(m)) <= {0, 1}
    spans = m.sum(-1)
    assert ((spans >= 10) & (spans <= 40)).all()                             # 3-4 spans of 10 frames, may overlap
    with pytest.raises(ValueError):
        _compute_mask_indices((1, 5), 0.05, 10)
    feats = torch.zeros(4, 768, 8)
    out = apply_spec_augmentation(feats, torch.ones(8), 0.05, 10)
    assert set(out.sum(-1).unique().tolist()) <= {0.0, 8.0}


def test_ctc_loss_surface():
    cfg = Wav2Vec2Config()
    loss = CTCLoss(cfg, (32, 246000), division_factor=32)
    assert loss._get_logit_length(246000) == 768                             # losses.py:47-56
    if torch.cuda.is_available():
        with pytest.raises(ValueError):
            loss(torch.ones(2, 8, dtype=torch.int32).cuda(), torch.zeros(2, 100, 32).cuda())


def test_posconv_backward_host_algebra():
    """Host-side pieces of the stage-2 backward (wav2vec2/training.py): the kernel of the positional conv's input gradient
    (flipped taps, in/out swapped per group, window shifted by one frame) and the weight-norm chain rule
    (tensorflow_addons.py:16-21), both against torch autograd on CPU."""
    import torch
    import torch.nn.functional as F
    from wav2vec2.training import transposed_conv_kernel, weight_norm_backward
    torch.manual_seed(0)
    B, T, G, cpg, k = 2, 37, 4, 8, 16
    d = G * cpg
    x = torch.randn(B, T, d, dtype=torch.float64, requires_grad=True)
    wv = torch.randn(k, cpg, d, dtype=torch.float64, requires_grad=True)
    wg = (1 + 0.3 * torch.rand(k, 1, 1, dtype=torch.float64)).requires_grad_()
    kern = wv * torch.rsqrt(torch.clamp(wv.pow(2).sum(dim=(1, 2), keepdim=True), min=1e-12)) * wg   # [k, cin/g, cout]

    def conv(inp, kr, left):           # out[t] = sum_j kr[j] . inp[t + j - left], zero outside [0, T)
        w = kr.permute(2, 1, 0)        # torch layout [cout, cin/g, k]
        xp = F.pad(inp.transpose(1, 2), (left, k - 1 - left))
        return F.conv1d(xp, w, groups=G).transpose(1, 2)
    pre = conv(x, kern, k // 2)        # encoder.py:177-181: pad k/2 both sides, drop the last frame
    dpre = torch.randn(B, T, d, dtype=torch.float64)
    pre.backward(dpre)
    # input gradient = forward conv of dpre with the transposed kernel, window shifted by one frame (left pad k/2 - 1)
    dx = conv(dpre, transposed_conv_kernel(kern.detach(), G), k // 2 - 1)
    assert torch.allclose(dx, x.grad, atol=1e-10)
    # weight-norm chain rule from the gradient w.r.t. the normalised kernel
    dkern = torch.autograd.grad(conv(x.detach(), kern, k // 2), kern, dpre)[0]
    dv, dg = weight_norm_backward(dkern, wv.detach(), wg.detach())
    assert torch.allclose(dv, wv.grad, atol=1e-10) and torch.allclose(dg, wg.grad, atol=1e-10)


def test_stage2_learning_rate_schedule():
    """training_utils.py:23-25."""
    from wav2vec2.finetune import FineTuneArgs, stage2_learning_rate
    a = FineTuneArgs(stage2_lr1=1e-4, stage2_lr2=5e-5, stage2_transition_epochs=2)
    assert [stage2_learning_rate(e, a) for e in range(5)] == [1e-4, 1e-4, 1e-4, 5e-5, 5e-5]


def test_weight_planes_of_every_precision_mode_decode_to_the_weight():
    """`_split(w, mode)` (the packing of every Dense / conv kernel) against `_effective_weight` (what the tensor cores multiply by):
    bf16 2^-9, bf16x3 2^-17, fp16 2^-12, fp16x3 2^-23, fp16f8 2^-15 relative (+ the e4m3 residual's subnormal floor);
    the fp16f8 byte plane holds, per 64-wide k-block, 64 hi bytes then 64 residual bytes (include/w2v2.h)."""
    import math
    from wav2vec2.modeling import _effective_weight, _split, _Modes
    torch.manual_seed(0)
    w = torch.randn(96, 128) / math.sqrt(128)
    for mode, tol in ((1, 2.0 ** -8), (3, 2.0 ** -16), (17, 2.0 ** -11), (19, 2.0 ** -21), (25, 2.0 ** -14)):
        p = _split(w, mode)
        eff = _effective_weight(p, mode)
        rel = ((eff - w).abs() / (w.abs() + 2.0 ** -8)).max().item()
        assert rel < tol, (mode, rel)
    p = _split(w, 25)
    assert p.hi.dtype == torch.float16 and p.lo.dtype == torch.uint8 and tuple(p.lo.shape) == (96, 256)
    blk = p.lo.reshape(96, 2, 2, 64)                                  # [row][k-block][hi8 | lo8][64]
    hi8 = blk[:, :, 0].contiguous().view(torch.float8_e4m3fn).float().reshape(96, 128)
    assert ((hi8 * 64 - p.hi.float()).abs() <= 0.07 * p.hi.float().abs() + 0.13).all()     # a 4-bit copy of the fp16 plane / 64
    assert torch.equal(_split(w, True).lo, _split(w, 3).lo) and _split(w, False).lo is None and _split(w, 1).lo is None
    for prec, (g, a, ps) in {"bf16": (1, 1, 1), "bf16x3": (3, 3, 3), "fp16": (17, 17, 17), "fp16f8": (25, 17, 17)}.items():
        m = _Modes(prec)
        assert (m.gemm, m.attn, m.pos) == (g, a, ps)
