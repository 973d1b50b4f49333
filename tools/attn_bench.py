"""Attention forward microbenchmark through the C ABI at the benchmark shape (B=32, T=768, 12 heads of 64): CUDA events around
one graph replay of ITERS back-to-back launches (the 113 MB qkv tensor + 38 MB output of one launch fit the 126 MB L2 only partly; REPS distinct
buffers are cycled so every launch reads cold data).  W2V2_ATTN_SPLIT=0/1 selects the softmax layout (read once per process)."""
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
B, T, H, dh = int(os.environ.get("B", 32)), int(os.environ.get("T", 768)), int(os.environ.get("H", 12)), 64
d = H * dh
iters, reps = int(os.environ.get("ITERS", 24)), 4
torch.manual_seed(0)
for passes in [int(v) for v in os.environ.get("PASSES", "1,17").split(",")]:
    dt = torch.float16 if passes & 16 else torch.bfloat16
    bufs = []
    for _ in range(reps):
        raw = torch.randn(B, T, 3 * d, device=dev) * 1.5
        raw[:, :, :d] *= dh ** -0.5
        if passes & 16:
            raw *= 16.0
        qkv = Pair(raw.to(dt).view(torch.bfloat16), raw.to(dt).view(torch.bfloat16) if passes & 3 == 3 else None)
        out = Pair(torch.zeros(B, T, d, dtype=torch.bfloat16, device=dev), torch.zeros(B, T, d, dtype=torch.bfloat16, device=dev))
        bufs.append((qkv, out))
    kv = torch.full((B,), T, dtype=torch.int32, device=dev)
    for i in range(4):
        ops.attn_fwd(bufs[i % reps][0], B, T, H, dh, kv, bufs[i % reps][1], passes)
    torch.cuda.synchronize()
    # one CUDA graph of ITERS launches: the host cost of a ctypes call + tensor-map encode (~90 us) must not pace the kernel
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            ops.attn_fwd(bufs[i % reps][0], B, T, H, dh, kv, bufs[i % reps][1], passes)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / iters
    flops = 4.0 * B * H * T * T * dh
    print(f"attn_fwd passes={passes} split={os.environ.get('W2V2_ATTN_SPLIT', 'default')} B={B} T={T} H={H}: {us:7.1f} us/launch  "
          f"{flops / us / 1e6:6.1f} TFLOP/s  (x12 layers = {us * 12 / 1e3:.3f} ms)")
