"""Row kernels of the train step's backward at its shapes (6144 rows = 8 x 768 frames): ln_bwd (d = 768) and dact_colsum
(column sum over 768 / 2304 columns, GELU' over 3072).  One CUDA graph of ITERS launches over rotating buffers, CUDA events.
W2V2_LNBWD_CTAS: CTAs of ln_bwd (each ends in up to 3 d global atomics)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))
import torch  # noqa: E402
from wav2vec2 import ops  # noqa: E402

M, d = 6144, 768
iters, reps = 24, 4


def timed(fn, label, nbytes):
    for i in range(4):
        fn(i % reps)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i % reps)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / iters
    print(f"{label:44s} {us:6.1f} us  {nbytes / us / 1e6:5.2f} TB/s")


x = [torch.randn(M, d, device="cuda") for _ in range(reps)]
dy = [torch.randn(M, d, device="cuda") for _ in range(reps)]
dx = [torch.empty(M, d, device="cuda") for _ in range(reps)]
dxh = [torch.empty(M, d, dtype=torch.bfloat16, device="cuda") for _ in range(reps)]
gam = torch.randn(d, device="cuda")
dg, db, cs = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
timed(lambda i: ops.ln_bwd(x[i], gam, dy[i], 1e-5, M, d, dx_f32=dx[i], dx_hi=dxh[i], dgamma=dg, dbeta=db, colsum=cs),
      f"ln_bwd {M} x {d} (ctas={os.environ.get('W2V2_LNBWD_CTAS', 'default')})", M * d * 14)
for cols, with_pre in ((768, False), (2304, False), (3072, True)):
    g_ = [torch.randn(M, cols, device="cuda").to(torch.bfloat16) for _ in range(reps)]
    pre = [torch.randn(M, cols, device="cuda") for _ in range(reps)] if with_pre else None
    out = [torch.empty(M, cols, dtype=torch.bfloat16, device="cuda") for _ in range(reps)] if with_pre else None
    csum = torch.zeros(cols, device="cuda")
    timed(lambda i: ops.dact_colsum(g_[i], pre[i] if with_pre else None, M, cols, out_hi=out[i] if with_pre else None, colsum=csum),
          f"dact_colsum {M} x {cols}{' + gelu grad' if with_pre else ''}", M * cols * (8 if with_pre else 2))
