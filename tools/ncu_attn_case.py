"""One window-attention forward + backward launch for an `ncu --set full` capture.
    python tools/ncu_attn_case.py H C heads shift images [window]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K, lib
lib.check(lib.load().fiber_init(), "init")
H, C, nh, shift, B = (int(a) for a in sys.argv[1:6])
ws = int(sys.argv[6]) if len(sys.argv) > 6 else 12
dev = torch.device("cuda:0")
qkv = torch.randn(B * H * H, 3 * C, device=dev).to(torch.bfloat16)
d_o = torch.randn(B * H * H, C, device=dev).to(torch.bfloat16)
tab = (torch.randn((2 * ws - 1) ** 2, nh, device=dev) * 0.5)
win = (B, H, H, ws, shift)
q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
for _ in range(2):
    o, lse = K.attn_fwd(q, k, v, nh, 32, 32 ** -0.5, window=win, bias_table=tab)
    dqkv = torch.empty_like(qkv); dt = torch.zeros_like(tab)
    K.attn_bwd(d_o, q, k, v, o, lse, nh, 32, 32 ** -0.5, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:],
               dbias_table=dt, window=win, bias_table=tab)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
o, lse = K.attn_fwd(q, k, v, nh, 32, 32 ** -0.5, window=win, bias_table=tab)
e1.record()
K.attn_bwd(d_o, q, k, v, o, lse, nh, 32, 32 ** -0.5, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:],
           dbias_table=dt, window=win, bias_table=tab)
e2.record()
torch.cuda.synchronize()
print("done: forward %.3f ms, backward %.3f ms" % (e0.elapsed_time(e1), e1.elapsed_time(e2)))
