import torch
n = 1_600_000_000
x = torch.empty(n // 2, dtype=torch.bfloat16, device="cuda")
y = torch.empty(n // 2, dtype=torch.bfloat16, device="cuda")
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it
ms = t(lambda: x.fill_(1.0)); print(f"fill 1.6 GB: {ms:.3f} ms  {n/ms/1e9:.2f} TB/s written")
ms = t(lambda: x.zero_()); print(f"memset 1.6 GB: {ms:.3f} ms  {n/ms/1e9:.2f} TB/s written")
ms = t(lambda: y.copy_(x)); print(f"copy 1.6 GB -> 1.6 GB: {ms:.3f} ms  {2*n/ms/1e9:.2f} TB/s read+write")
ms = t(lambda: x.sum()); print(f"read 1.6 GB: {ms:.3f} ms  {n/ms/1e9:.2f} TB/s read")
