import os, sys
sys.path.insert(0, "gsoc-wav2vec2_b200")
import torch
from wav2vec2 import ops
from wav2vec2.ops import Pair
torch.manual_seed(0)
for (B, T, H) in [(2, 49, 12), (2, 145, 4), (8, 768, 12), (3, 300, 16)]:
    d = H * 64
    raw = torch.randn(B, T, 3 * d, device="cuda") * 1.5
    raw[:, :, :d] *= 0.125
    qkv = Pair(raw.to(torch.bfloat16), None)
    kv = torch.tensor([T] + [max(1, T - 37)] * (B - 1), dtype=torch.int32, device="cuda")
    for drop in (None, (0.1, 1234, 7)):
        ref = None
        bad = 0
        for it in range(200):
            out = Pair(torch.zeros(B, T, d, dtype=torch.bfloat16, device="cuda"), None)
            if drop is None:
                ops.attn_fwd(qkv, B, T, H, 64, kv, out, 1)
            else:
                ops.attn_fwd_train(qkv, B, T, H, 64, kv, out, 1, drop)
            torch.cuda.synchronize()
            if ref is None:
                ref = out.hi.clone()
            elif not torch.equal(ref, out.hi):
                bad += 1
                if bad == 1:
                    diff = (ref.float() - out.hi.float()).abs()
                    idx = (diff > 0).nonzero()
                    print("   first mismatch: count", idx.shape[0], "max", diff.max().item(), "rows", sorted(set((int(i[0]), int(i[1])) for i in idx[:2000]))[:8], "cols", sorted(set(int(i[2]) // 64 for i in idx[:2000]))[:12])
        print(f"B={B} T={T} H={H} drop={drop is not None}: {bad} / 199 runs differ from the first, nan={bool(torch.isnan(ref.float()).any())}")
