"""ln_rows microbenchmark at the benchmark shape (24576 x 768 fp32 -> bf16 operand + row statistics): one CUDA graph of ITERS
launches over rotating buffers, CUDA events.  W2V2_LN_CTAS_PER_SM selects the persistent grid (999: one row per warp)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))
import torch  # noqa: E402
from wav2vec2 import ops  # noqa: E402

M, d = int(os.environ.get("M", 24576)), int(os.environ.get("D", 768))
iters, reps = 24, 4
xs = [torch.randn(M, d, device="cuda") for _ in range(reps)]
hi = [torch.empty(M, d, dtype=torch.bfloat16, device="cuda") for _ in range(reps)]
st = torch.empty(M, 2, device="cuda")
g_, b_ = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
for i in range(4):
    ops.ln_rows(xs[i % reps], g_, b_, 1e-5, M, d, out_hi=hi[i % reps], stats=st)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(iters):
        ops.ln_rows(xs[i % reps], g_, b_, 1e-5, M, d, out_hi=hi[i % reps], stats=st)
g.replay()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
g.replay()
e.record()
torch.cuda.synchronize()
us = s.elapsed_time(e) * 1e3 / iters
byt = M * d * 6 + M * 8
print(f"ln_rows M={M} d={d} per_sm={os.environ.get('W2V2_LN_CTAS_PER_SM', 'default')}: {us:6.1f} us  {byt / us / 1e6:6.2f} TB/s")
