"""Section timing of the stage-2 train step (CUDA events) + per-kernel-class share: where the step's time goes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from wav2vec2 import CTCLoss, Wav2Vec2Config, Wav2Vec2ForCTC, ops  # noqa: E402
from wav2vec2.training import Stage2Trainer  # noqa: E402

B, L = int(os.environ.get("BATCH", "8")), int(os.environ.get("SEQ", "246000"))
cfg = Wav2Vec2Config(dropout=float(os.environ.get("DROPOUT", "0.1")))
model = Wav2Vec2ForCTC(cfg, input_shape=(B, L), precision="bf16").init_random(0)
tr = Stage2Trainer(model, CTCLoss(cfg, (B, L), division_factor=B))
x = torch.randn(B, L, generator=torch.Generator().manual_seed(0)).cuda()
lab = np.zeros((B, 256), dtype=np.int32)
lab[:, :24] = np.random.randint(1, 30, size=(B, 24))
labels = torch.from_numpy(lab).cuda()
for _ in range(3):
    tr.step(x, labels)
torch.cuda.synchronize()


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


import time  # noqa: E402
acc = {}
for _ in range(5):
    t0 = time.perf_counter()
    e = [ev()]
    logits = tr._forward(x); e.append(ev())
    loss, dl = tr.loss_fn(labels, logits, return_grad=True); e.append(ev())
    tr._backward(dl); e.append(ev())
    tr.t += 1
    ops.adam(tr.flat_w, tr.flat_g, tr.m, tr.v, 1e-5, 0.9, 0.999, 1e-7); e.append(ev())
    tr._repack(); e.append(ev())
    torch.cuda.synchronize()
    host = (time.perf_counter() - t0) * 1e3
    for name, a, b in zip(("forward", "ctc", "backward", "adam", "repack"), e[:-1], e[1:]):
        acc[name] = acc.get(name, 0.0) + a.elapsed_time(b) / 5
    acc["host wall"] = acc.get("host wall", 0.0) + host / 5
print({k: round(v, 3) for k, v in acc.items()})

# per-op device time of the backward (eager, events around each wrapper)
records = []
orig = {}
for name in ("gemm", "ln_bwd", "dact_colsum", "transpose_bf16", "attn_bwd", "posconv_train", "posconv_wgrad", "lm_head_wgrad",
             "lm_head_dgrad", "split_bf16", "gelu_rows", "ln_rows", "attn_fwd", "attn_fwd_train", "dropout_rows", "pack_weights"):
    fn = getattr(ops, name)
    orig[name] = fn

    def make(fn, name):
        def inner(*a, **kw):
            s = ev()
            r = fn(*a, **kw)
            records.append((name if name != "gemm" else f"gemm K={kw.get('K')} N={kw.get('N')} rows={kw.get('rows_per_batch')}", s, ev()))
            return r
        return inner
    setattr(ops, name, make(fn, name))
logits = tr._forward(x)
loss, dl = tr.loss_fn(labels, logits, return_grad=True)
n_fwd = len(records)
tr._backward(dl)
torch.cuda.synchronize()
agg = {}
for i, (k, s, e) in enumerate(records):
    key = ("fwd " if i < n_fwd else "bwd ") + k
    t, c = agg.get(key, (0.0, 0))
    agg[key] = (t + s.elapsed_time(e), c + 1)
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{k:52s} {c:4d} launches {t:8.3f} ms")
