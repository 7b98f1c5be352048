import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K, lib
lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
B, nh, hd, L = 3, 1, 64, 40
C = nh * hd
g = torch.Generator().manual_seed(1)
q = torch.randn(B * L, C, generator=g).to(dev).to(torch.bfloat16)
k = torch.randn(B * L, C, generator=g).to(dev).to(torch.bfloat16)
scale = 1 / math.sqrt(hd)
def run(v, mode):
    lib.set_option("attn_sk", mode)
    o, lse = K.attn_fwd(q, k, v, nh, hd, scale, groups=B, lq=L, lk=L)
    lib.set_option("attn_sk", -1)
    return o.float(), lse
# 1) V = ones -> o == 1 everywhere if P rows sum to 1 over the right keys
v1 = torch.ones(B * L, C, device=dev, dtype=torch.bfloat16)
o, lse = run(v1, 5)
o0, lse0 = run(v1, 0)
print("V=1: o per group mean", [round(o[i*L:(i+1)*L].mean().item(), 3) for i in range(B)], "lse err", (lse - lse0).abs().max().item())
# 2) V = key index in column 0 -> o[:,0] = expected key index under P
v2 = torch.zeros(B * L, C, device=dev, dtype=torch.bfloat16)
v2[:, 0] = (torch.arange(B * L, device=dev) % L).to(torch.bfloat16)
v2[:, 1] = (torch.arange(B * L, device=dev) // L).to(torch.bfloat16)
o, _ = run(v2, 5); o0, _ = run(v2, 0)
for i in range(B):
    print("group", i, "E[key] sk", o[i*L:(i+1)*L, 0][:4].tolist(), "ref", o0[i*L:(i+1)*L, 0][:4].tolist(), "| E[group]", o[i*L:(i+1)*L, 1][:4].tolist())
# 3) random V: error per group and per hd column block
v3 = torch.randn(B * L, C, generator=g).to(dev).to(torch.bfloat16)
o, _ = run(v3, 5); o0, _ = run(v3, 0)
e = (o - o0).abs()
print("rand V: err per group", [round(e[i*L:(i+1)*L].max().item(), 3) for i in range(B)], "per 16-col block", [round(e[:, j:j+16].max().item(), 3) for j in range(0, 64, 16)])
print("rows with err>0.1 in group 1:", (e[L:2*L].max(1).values > 0.1).nonzero().flatten().tolist())
