"""conv0 (+ GroupNorm scale / shift + GELU) microbenchmark at the benchmark shape (32 x 246000): graph of back-to-back launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))
import torch
from wav2vec2 import ops
from wav2vec2.ops import Pair
B, L, C = 32, 246000, 512
T0 = 1 + (L - 10) // 5
x = torch.randn(B, L, device="cuda")
k = torch.randn(10, C, device="cuda") * 0.3
fs, fb = torch.rand(B, C, device="cuda") + 0.5, torch.randn(B, C, device="cuda") * 0.1
outs = [Pair(torch.empty(B * T0, C, dtype=torch.bfloat16, device="cuda"), None) for _ in range(2)]
for i in range(3):
    ops.conv0_gn_gelu(x, k, fs, fb, outs[i % 2], 1)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(10):
        ops.conv0_gn_gelu(x, k, fs, fb, outs[i % 2], 1)
g.replay(); torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); g.replay(); e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
by = B * (4.0 * L + 2.0 * C * T0)
print(f"conv0 32 x 246000: {ms * 1e3:.1f} us  {by / ms / 1e9:.2f} TB/s algorithmic")
