"""GEMM cost of the precision modes (W2V2_MODE_*: 1 bf16, 17 fp16, 25 fp16 + e4m3 cross terms, 3 bf16x3) on the encoder /
extractor shapes, through the C ABI.  Two regimes per (shape, mode): ISOLATED launches (L2 flushed, a sync between launches:
the clock stays at its maximum) and a SUSTAINED back-to-back loop (the 1 kW power cap sets the clock); the SM clock of the
sustained loop is sampled through NVML.  Tells apart what a mode costs in tensor-pipe cycles from what it costs under the cap."""
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))
import torch  # noqa: E402
from wav2vec2 import ops  # noqa: E402
from wav2vec2.modeling import _KINDS, _split  # noqa: E402
from wav2vec2.ops import Pair  # noqa: E402

dev = "cuda"
import pynvml  # noqa: E402
pynvml.nvmlInit()
H = pynvml.nvmlDeviceGetHandleByIndex(0)


def planes(x, mode):
    if mode in (1, 3):
        hi = x.to(torch.bfloat16)
        return Pair(hi, (x - hi.float()).to(torch.bfloat16) if mode == 3 else None)
    s = torch.clamp(x * 16.0, -65504, 65504)
    hi = s.to(torch.float16)
    if mode == 17:
        return Pair(hi, None)
    res = s - hi.float()
    rows, K = x.shape
    l8 = torch.clamp(res * 64, -448, 448).to(torch.float8_e4m3fn).view(torch.uint8).reshape(rows, K // 64, 1, 64)
    h8 = torch.clamp(hi.float() / 64, -448, 448).to(torch.float8_e4m3fn).view(torch.uint8).reshape(rows, K // 64, 1, 64)
    return Pair(hi, torch.cat([l8, h8], 2).reshape(rows, 2 * K).contiguous())


shapes = [("ffn1+gelu", 24576, 768, 3072, True), ("qkv", 24576, 768, 2304, False), ("ffn2", 24576, 3072, 768, False),
          ("conv1-like+gelu", 24576 * 8, 1536, 512, True)]
only = os.environ.get("ONLY")
if only:
    shapes = [s for s in shapes if s[0] in only.split(",")]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for name, M, K, N, gelu in shapes:
    x32 = torch.randn(M, K, device=dev)
    w32 = torch.randn(N, K, device=dev) / math.sqrt(K)
    bias = torch.randn(N, device=dev)
    for mode in (1, 17, 25, 3):
        a, w = planes(x32, mode), _split(w32, mode)
        kind = {1: "bf16", 3: "bf16x2", 17: "fp16", 25: "fp16f8"}[mode]
        hi_dt, lo_kind, ofmt = _KINDS[kind]
        out_hi = torch.empty(M, N, dtype=hi_dt, device=dev)
        out_lo = None if lo_kind is None else (torch.empty(M, N, dtype=hi_dt, device=dev) if lo_kind == "same"
                                               else torch.empty(M, 2 * N, dtype=torch.uint8, device=dev))

        def run():
            ops.gemm(a, w, K=K, N=N, rows_per_batch=M, bias=bias, gelu=gelu, out_hi=out_hi, out_lo=out_lo, passes=mode, out_format=ofmt)
        ts = []
        for i in range(8):
            flush.zero_()
            torch.cuda.synchronize()
            time.sleep(0.02)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); run(); e.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(s.elapsed_time(e))
        iso = sorted(ts)[len(ts) // 2]
        clocks, stop = [], threading.Event()

        def sample():
            while not stop.is_set():
                clocks.append(pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM))
                time.sleep(0.005)
        n = max(20, int(300.0 / iso))          # ~0.3 s of back-to-back launches
        for _ in range(n // 4):
            run()
        th = threading.Thread(target=sample); th.start()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            run()
        e.record()
        torch.cuda.synchronize()
        stop.set(); th.join()
        sus = s.elapsed_time(e) / n
        mhz = sorted(clocks)[len(clocks) // 2] if clocks else 0
        fl = 2.0 * M * K * N
        print(f"{name:16s} mode {mode:2d}: isolated {iso * 1e3:8.1f} us ({fl / iso / 1e9:7.1f} TF/s algorithmic)   sustained {sus * 1e3:8.1f} us "
              f"({fl / sus / 1e9:7.1f} TF/s) at {mhz} MHz -> {sus * mhz:9.0f} kcycles", flush=True)
