import sys, os, torch
sys.path.insert(0, "gsoc-wav2vec2_b200"); sys.path.insert(0, ".")
from wav2vec2 import Wav2Vec2Config, Wav2Vec2ForCTC, ops
from wav2vec2.ops import Pair
from oracle import w2v2_oracle as O
DEV = "cuda:0"
torch.manual_seed(9)
B, C, L = 4, 512, 16000
x = torch.randn(B, L).to(DEV)
kd = (torch.randn(10, C) * 0.3).to(DEV)
gamma, beta = (1 + 0.1 * torch.randn(C)).to(DEV), (0.1 * torch.randn(C)).to(DEV)
def run(xd):
    b = xd.shape[0]
    T0 = 1 + (L - 10) // 5
    stats = torch.empty(b, 65, dtype=torch.float64, device=DEV)
    fs, fb = torch.empty(b, C, device=DEV), torch.empty(b, C, device=DEV)
    ops.wave_stats(xd, stats)
    ops.conv0_fold(kd, gamma, beta, stats, b, L, None, fb, scale=fs)
    hi = torch.zeros(b, T0, C, dtype=torch.bfloat16, device=DEV)
    ops.conv0_gn_gelu(xd, kd, fs, fb, Pair(hi, None), 1)
    torch.cuda.synchronize()
    return hi, fs, fb, stats
h4, fs4, fb4, st4 = run(x)
h2a, fs2, fb2, st2 = run(x[:2].contiguous())
h2b = run(x[2:].contiguous())[0]
print("stats eq", torch.equal(st4[:2], st2), "fs eq", torch.equal(fs4[:2], fs2), "fb eq", torch.equal(fb4[:2], fb2))
print("conv0 eq", torch.equal(h4[:2], h2a), torch.equal(h4[2:], h2b), (h4[:2].float() - h2a.float()).abs().max().item())
h4b = run(x)[0]
print("repeat eq", torch.equal(h4, h4b))
cfg = Wav2Vec2Config(num_layers=2)
params = O.random_params(cfg, seed=5)
m = Wav2Vec2ForCTC(cfg, input_shape=(2, 16000), precision="bf16", device=DEV)
m.set_variables(params)
xx = torch.randn(4, 16000, generator=torch.Generator().manual_seed(3)).cuda()
f1 = m(xx).clone(); f2 = m(xx).clone()
p1 = torch.cat([m(xx[:2].contiguous()).clone(), m(xx[2:].contiguous()).clone()])
print("model repeat eq", torch.equal(f1, f2), "shard eq", torch.equal(f1, p1), (f1 - p1).abs().max().item())
A = m._arena
print([k for k in A.__dict__])
