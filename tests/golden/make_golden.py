"""Regenerates tests/golden/*.npz.  Run from the repo root in the BUILD container:  python tests/golden/make_golden.py

The reference (thevasudevgupta/gsoc-wav2vec2) cannot be imported here - its arithmetic lives in the un-vendored
tensorflow==2.5 wheel - so the golden outputs come from the implementation the reference's OWN tests equate the
TF model to (atol 1e-3 / 4e-3): `transformers` PyTorch Wav2Vec2 (tests/test_wav2vec2.py:47-79,109-170), loaded with
seeded weights through the reference's weight-mapping rules (src/convert_torch_to_tf.py:88-123).  Only inputs and
expected outputs are stored; the weights are regenerated from the seed (oracle.random_params) at test time.
Also stored: the processor known-answer vector held by the reference (tests/test_dataloader.py:56-63) and CTC
losses from torch's ctc_loss for the reference's test labels (tests/test_wav2vec2.py:41-42,214-237).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))

import transformers  # noqa: E402
from oracle import w2v2_oracle as O  # noqa: E402
from wav2vec2.config import RobustWav2Vec2Config, Wav2Vec2Config  # noqa: E402
from wav2vec2.weights import reference_to_hf  # noqa: E402

CASES = {
    # name: (config, seed, num_samples, masked tail per row)
    "base_small": (Wav2Vec2Config(hidden_size=128, num_heads=2, num_layers=2, intermediate_size=256,
                                  num_conv_pos_embedding_groups=2), 11, 8000, None),
    "robust_small": (RobustWav2Vec2Config(hidden_size=128, num_heads=2, num_layers=2, intermediate_size=256,
                                          num_conv_pos_embedding_groups=2), 12, 8000, (1000, 132)),
}


def hf_model(cfg):
    hc = transformers.Wav2Vec2Config(
        vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_layers,
        num_attention_heads=cfg.num_heads, intermediate_size=cfg.intermediate_size, conv_bias=cfg.conv_bias,
        feat_extract_norm=cfg.feature_extractor_norm_type, do_stable_layer_norm=cfg.attention_norm_type == "prenorm",
        num_conv_pos_embeddings=cfg.num_conv_pos_embeddings, num_conv_pos_embedding_groups=cfg.num_conv_pos_embedding_groups,
        hidden_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, final_dropout=0.0, layerdrop=0.0,
        activation_dropout=0.0)
    return transformers.Wav2Vec2ForCTC(hc).eval()


def main():
    for name, (cfg, seed, L, tails) in CASES.items():
        params = O.random_params(cfg, seed=seed)
        m = hf_model(cfg)
        missing = m.load_state_dict(reference_to_hf(params), strict=False)
        assert not missing.missing_keys and not missing.unexpected_keys, missing
        x = torch.randn(2, L, generator=torch.Generator().manual_seed(seed + 100))
        am = None
        if tails is not None:
            am = torch.ones(2, L, dtype=torch.long)
            am[0, -tails[0]:] = 0
            am[1, -tails[1]:] = 0
        with torch.no_grad():
            out = m(x, attention_mask=am, output_hidden_states=False)
            hid = m.wav2vec2(x, attention_mask=am).last_hidden_state
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), speech=x.numpy(), logits=out.logits.numpy(),
                            hidden=hid.numpy(), seed=seed,
                            attention_mask=(am.numpy() if am is not None else np.zeros(0)))
        print(name, "logits", tuple(out.logits.shape), "max|logit|", float(out.logits.abs().max()))
    # processor known answer (reference tests/test_dataloader.py:60-62) and our reading of the same wav
    wav = O.read_wav_s16(os.path.join(HERE, "sample.wav"))
    norm = O.normalize_utterance(wav[None, :])
    np.savez_compressed(os.path.join(HERE, "processor.npz"),
                        reference_vector=np.array([0.01438822, 0.01776027, 0.01438822, 0.02113231, 0.01438822,
                                                   0.00764414, 0.00764414, -0.00921606], np.float32),
                        normalized_32_40=norm[32:40], num_samples=len(wav))
    # CTC: the reference's test labels, logits N(0,1)*2, loss via torch.nn.functional.ctc_loss (what HF uses)
    np.random.seed(0)
    labels = np.random.randint(1, 30, size=(2, 24))
    logits = torch.randn(2, 145, 32, generator=torch.Generator().manual_seed(5)) * 2
    lp = torch.log_softmax(logits.double(), -1).transpose(0, 1)
    loss = torch.nn.functional.ctc_loss(lp, torch.from_numpy(labels), torch.full((2,), 145), torch.full((2,), 24),
                                        blank=0, reduction="none")
    np.savez_compressed(os.path.join(HERE, "ctc.npz"), labels=labels, logits=logits.numpy(), loss_per_sample=loss.numpy())
    print("ctc", loss.numpy())


if __name__ == "__main__":
    main()
