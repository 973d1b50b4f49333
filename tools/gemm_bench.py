"""GEMM microbenchmark through the C ABI: TFLOP/s per variant (cluster=0: cta_group::2 pairs, 1: single CTAs,
3: 1-SM MMAs + pair multicast) on the encoder / extractor shapes.  CUDA events, L2 flushed between launches."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))
import torch  # noqa: E402
from wav2vec2 import _lib  # noqa: E402
if os.environ.get("W2V2_LIB_DIR"):       # A/B of compile-time variants: a library built into another directory
    _lib.LIB_PATH = os.path.join(os.environ["W2V2_LIB_DIR"], "libw2v2_sm100.so")
from wav2vec2 import ops  # noqa: E402
from wav2vec2.ops import Pair  # noqa: E402

dev = "cuda"
shapes = [("ffn1", 24576, 768, 3072, True), ("qkv", 24576, 768, 2304, False), ("out", 24576, 768, 768, False),
          ("ffn2", 24576, 3072, 768, False), ("conv1-like", 24576 * 8, 1536, 512, True),
          ("out+res32", 24576, 768, 768, "res"), ("ffn2+res32", 24576, 3072, 768, "res"), ("out f32", 24576, 768, 768, "f32")]
only = os.environ.get("ONLY")
if only:
    shapes = [s for s in shapes if s[0] in only.split(",")]
variants = [int(v) for v in os.environ.get("VARIANTS", "0,1,3").split(",")]
iters = int(os.environ.get("ITERS", "10"))
debug = int(os.environ.get("DEBUG", "0"))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for name, M, K, N, gelu in shapes:
    a = Pair(torch.randn(M, K, device=dev).to(torch.bfloat16))
    w = Pair((torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16))
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    res = torch.randn(M, N, device=dev) if gelu == "res" else None
    o32 = torch.empty(M, N, device=dev) if gelu in ("res", "f32") else None
    for cl in variants:
        ts = []
        for i in range(iters + 2):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            if o32 is not None:
                ops.gemm(a, w, K=K, N=N, rows_per_batch=M, bias=bias, residual=res, out_f32=o32, cluster=cl, debug=debug)
            else:
                ops.gemm(a, w, K=K, N=N, rows_per_batch=M, bias=bias, gelu=gelu, out_hi=out, cluster=cl, debug=debug)
            e.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(s.elapsed_time(e))
        t = sorted(ts)[len(ts) // 2]
        print(f"{name:11s} M={M} K={K} N={N} cluster={cl}: {t * 1e3:8.1f} us  {2.0 * M * K * N / t / 1e9:7.1f} TFLOP/s")
