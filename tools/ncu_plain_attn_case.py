"""A forward + backward launch of one plain-attention shape for an `ncu --set full` capture.
    python tools/ncu_plain_attn_case.py i2t|self [samples]"""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K, lib
lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
case = sys.argv[1] if len(sys.argv) > 1 else "i2t"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nh, hd, Lq, Lk, drop = (16, 32, 576, 40, 0.0) if case == "i2t" else (12, 64, 40, 40, 0.1)
C = nh * hd
q = torch.randn(B * Lq, C, device=dev).to(torch.bfloat16)
kv = torch.randn(B * Lk, 2 * C, device=dev).to(torch.bfloat16)
d_o = torch.randn(B * Lq, C, device=dev).to(torch.bfloat16)
mask = torch.zeros(B, Lk, device=dev); mask[::2, Lk - 9:] = -10000.0
kw = dict(groups=B, lq=Lq, lk=Lk, key_mask=mask, drop_p=drop, seed=7)
for _ in range(2):
    o, lse = K.attn_fwd(q, kv[:, :C], kv[:, C:], nh, hd, 1 / math.sqrt(hd), **kw)
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    K.attn_bwd(d_o, q, kv[:, :C], kv[:, C:], o, lse, nh, hd, 1 / math.sqrt(hd), dq, dkv[:, :C], dkv[:, C:], **kw)
torch.cuda.synchronize()
print("done")
