"""One warm-up + one measured forward of a shallow model at the benchmark shape, for ncu captures:
   ncu --set full --clock-control none --import-source on -s <launches of forward 1> -c <launches of forward 2> ...
Layers are identical, so a 1-layer model exposes every kernel/shape of the 12-layer one."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))
import torch  # noqa: E402
from wav2vec2 import Wav2Vec2Config, Wav2Vec2ForCTC, ops  # noqa: E402

layers = int(os.environ.get("LAYERS", "1"))
B = int(os.environ.get("BATCH", "32"))
L = int(os.environ.get("SEQ", "246000"))
prec = os.environ.get("PRECISION", "bf16")
cfg = Wav2Vec2Config(num_layers=layers)
m = Wav2Vec2ForCTC(cfg, input_shape=(B, L), precision=prec).init_random(0)
x = torch.randn(B, L, generator=torch.Generator().manual_seed(0)).cuda()
for i in range(2):
    n0 = ops.LAUNCHES
    m(x)
    torch.cuda.synchronize()
    print(f"forward {i}: {ops.LAUNCHES - n0} launches")
