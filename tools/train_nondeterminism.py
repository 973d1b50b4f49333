import sys
sys.path.insert(0, "gsoc-wav2vec2_b200"); sys.path.insert(0, ".")
import numpy as np, torch
from oracle import w2v2_oracle as O
from wav2vec2 import CTCLoss, Wav2Vec2Config, Wav2Vec2ForCTC
from wav2vec2.training import Stage2Trainer
cfg = Wav2Vec2Config(num_layers=2, dropout=0.1, apply_spec_augment=False)
B, L = 2, 16000
g = torch.Generator().manual_seed(1)
xs = [torch.randn(B, L, generator=g).cuda() for _ in range(2)]
np.random.seed(0)
labels = torch.from_numpy(np.random.randint(1, 30, size=(B, 10))).int().cuda()
runs = []
for k in range(8):
    prefetch = k % 2 == 1
    m = Wav2Vec2ForCTC(cfg, input_shape=(1, 2048), precision="bf16")
    m.set_variables(O.random_params(cfg, seed=4))
    tr = Stage2Trainer(m, CTCLoss(cfg, (B, L), division_factor=B), learning_rate=1e-4, seed=3)
    losses = []
    for i in range(3):
        nxt = xs[(i + 1) % 2] if prefetch else None
        losses.append(tr.step(xs[i % 2], labels, next_speech=nxt).item())
    runs.append((losses, tr.flat_w.clone()))
    print(k, "prefetch" if prefetch else "plain   ", losses)
for k in range(1, 8):
    print(k, "loss diff vs run 0: %.2e" % max(abs(a - b) for a, b in zip(runs[0][0], runs[k][0])), "weight diff %.2e" % (runs[0][1] - runs[k][1]).abs().max().item())
