"""Timeline of the persistent attention forward kernel from %globaltimer stamps (needs a library built with -DAT_STAMPS:
   mkdir -p gsoc-wav2vec2_b200/lib_stamps && nvcc <flags of csrc/Makefile> -DAT_STAMPS -c csrc/attn.cu -o lib_stamps/attn.o
   && nvcc -shared -o lib_stamps/libw2v2_sm100.so lib_stamps/attn.o <the other objects of lib/>).
Stamp slots per tile (thread 0 = softmax warp 0): 0 tile start, 3..8 S(j) seen, 9 chunk loop done, 10 last PV retired, 11 context written."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))
import torch, numpy as np
from wav2vec2 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "gsoc-wav2vec2_b200", "lib_stamps", "libw2v2_sm100.so")
from wav2vec2 import ops
from wav2vec2.ops import Pair
B, T, H, dh = 32, 768, 12, 64
d = H * dh
raw = torch.randn(B, T, 3 * d, device="cuda") * 1.5
raw[:, :, :d] *= dh ** -0.5
qkv = Pair(raw.to(torch.bfloat16), None)
out = Pair(torch.zeros(B, T, d, dtype=torch.bfloat16, device="cuda"), None)
kv = torch.full((B,), T, dtype=torch.int32, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    ops.attn_fwd(qkv, B, T, H, dh, kv, out, 1)
torch.cuda.synchronize()
nt = 6 * H * B
buf = (C.c_ulonglong * (nt * 32))()
lib = _lib.load()
lib.w2v2_attn_debug_stamps.argtypes = [C.c_void_p, C.c_int]
assert lib.w2v2_attn_debug_stamps(buf, nt * 32) == 0
a = np.frombuffer(buf, dtype=np.uint64).reshape(nt, 32).astype(np.int64)
t0 = a[:, 0].min()
print("kernel span us (first tile start -> last context written):", (a[:, 11].max() - t0) / 1e3)
names = {0: "tile start", 3: "S0 seen", 4: "S1 seen", 5: "S2 seen", 6: "S3 seen", 7: "S4 seen", 8: "S5 seen", 9: "loop done", 10: "PV retired", 11: "ctx written"}
rel = a - a[:, [0]]
print("median time since tile start (us), all tiles / first tile of a CTA / later tiles:")
first = np.arange(nt) < 296
for i, nm in names.items():
    print(f"  {nm:12s} {np.median(rel[:, i]) / 1e3:7.2f}   {np.median(rel[first, i]) / 1e3:7.2f}   {np.median(rel[~first, i]) / 1e3:7.2f}")
# CTA 0 walks tiles 0, 296, 592, ...
for cta in (0, 1, 295):
    tiles = list(range(cta, nt, 296))
    print(f"CTA {cta}: tile (start, end) us:", [(round((a[t, 0] - t0) / 1e3, 1), round((a[t, 11] - t0) / 1e3, 1)) for t in tiles])

ncta = 296
per = []
for cta in range(ncta):
    tiles = list(range(cta, nt, ncta))
    per.append(((a[tiles[-1], 11] - a[tiles[0], 0]) / 1e3 / len(tiles), len(tiles), int(a[cta, 15]), (a[tiles[-1], 11] - t0) / 1e3))
per = np.array(per)
print("per-CTA mean tile time us: min %.2f median %.2f p90 %.2f max %.2f" % (per[:, 0].min(), np.median(per[:, 0]), np.percentile(per[:, 0], 90), per[:, 0].max()))
print("CTA finish time us: min %.1f median %.1f max %.1f" % (per[:, 3].min(), np.median(per[:, 3]), per[:, 3].max()))
slow = np.argsort(-per[:, 0])[:12]
print("slowest CTAs (cta, sm, mean tile us, tiles):", [(int(c), int(per[c, 2]), round(per[c, 0], 2), int(per[c, 1])) for c in slow])
sm_of = per[:, 2].astype(int)
import collections
cnt = collections.Counter(sm_of.tolist())
print("CTAs per SM histogram:", collections.Counter(cnt.values()))
by_sm = collections.defaultdict(list)
for c in range(ncta):
    by_sm[sm_of[c]].append(round(per[c, 0], 2))
print("pairs on the SMs of the slowest CTAs:", {int(sm_of[c]): by_sm[sm_of[c]] for c in slow[:6]})
