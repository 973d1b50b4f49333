"""Logits / hidden-state parity at the FULL BASELINE sizes (246000 samples -> 768 frames), both architectures, both
precisions, against the CPU oracle on the same seeded weights (the oracle is the checker only).  Prints one JSON line per case."""
import json
import logging
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tests/ -> repo root
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))
import torch  # noqa: E402
from oracle import w2v2_oracle as O  # noqa: E402
from wav2vec2 import RobustWav2Vec2Config, Wav2Vec2Config, Wav2Vec2ForCTC  # noqa: E402

logging.getLogger("wav2vec2.modeling").setLevel(logging.ERROR)
torch.set_num_threads(os.cpu_count() or 1)
L, B = 246000, 2
for name, cfg in (("base (12 layers, d=768, group-norm extractor, post-norm)", Wav2Vec2Config()),
                  ("large / robust (24 layers, d=1024, layer-norm extractor, pre-norm, attention mask)", RobustWav2Vec2Config())):
    params = O.random_params(cfg, seed=0)
    x = torch.randn(B, L, generator=torch.Generator().manual_seed(0))
    mask = None
    if cfg.is_robust:
        mask = torch.ones(B, L, dtype=torch.int32)
        mask[0, -1000:] = 0                                 # tests/test_wav2vec2.py:58-62
        mask[1, -132:] = 0
        x = x * mask
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = O.wav2vec2_for_ctc(x, params, cfg, attention_mask=mask)
    cpu_s = time.perf_counter() - t0
    for precision in ("bf16x3", "bf16"):
        m = Wav2Vec2ForCTC(cfg, input_shape=(B, L), precision=precision)
        m.set_variables(params)
        got = m(x.cuda(), attention_mask=None if mask is None else mask.cuda()).cpu()
        err = (got - ref).abs().max().item()
        agree = (got.argmax(-1) == ref.argmax(-1)).float().mean().item()
        print(json.dumps({"model": name, "precision": precision, "batch": B, "seq": L, "frames": int(ref.shape[1]),
                          "logits_max_abs": round(ref.abs().max().item(), 3), "max_abs_err": err, "argmax_agreement": agree,
                          "oracle_cpu_seconds": round(cpu_s, 1)}), flush=True)
        del m
        torch.cuda.empty_cache()
