"""Throughput sweep over BASELINE.json configs[3..4]: seq x batch for wav2vec2-base, plus wav2vec2-large (24 layers,
d=1024, layer-norm convs, attention mask) at batch 16 x 246000.  CUDA events, 3 warm-ups, 5 timed forwards each."""
import json
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))
import torch  # noqa: E402
from wav2vec2 import RobustWav2Vec2Config, Wav2Vec2Config, Wav2Vec2ForCTC  # noqa: E402


def time_model(model, x, mask=None, steps=5):
    for _ in range(3):
        model(x, attention_mask=mask)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        model(x, attention_mask=mask)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def main():
    prec = os.environ.get("PRECISION", "bf16")
    graph = os.environ.get("GRAPH", "0") == "1"
    base = Wav2Vec2ForCTC(Wav2Vec2Config(), precision=prec).init_random(0)
    if graph:
        base.enable_cuda_graph(True)
    for L in (16000, 64000, 246000):
        for B in (1, 8, 32, 128):
            if B == 128 and L == 246000:
                continue
            x = torch.randn(B, L, device="cuda")
            ms = time_model(base, x)
            print(json.dumps({"model": "base", "precision": prec, "graph": graph, "batch": B, "seq": L,
                              "ms": round(ms, 3), "audio_s_per_s": round(B * L / 16000 / (ms / 1e3), 1)}), flush=True)
    del base
    torch.cuda.empty_cache()
    if os.environ.get("LARGE", "1") == "1":
        logging.getLogger("wav2vec2.modeling").setLevel(logging.ERROR)
        large = Wav2Vec2ForCTC(RobustWav2Vec2Config(), precision=prec).init_random(0)
        B, L = 16, 246000
        x = torch.randn(B, L, device="cuda")
        mask = torch.ones(B, L, dtype=torch.int32, device="cuda")
        mask[0, -1000:] = 0
        ms = time_model(large, x, mask)
        print(json.dumps({"model": "large (robust, 24 layers, d=1024)", "precision": prec, "batch": B, "seq": L,
                          "ms": round(ms, 3), "audio_s_per_s": round(B * L / 16000 / (ms / 1e3), 1)}), flush=True)


if __name__ == "__main__":
    main()
