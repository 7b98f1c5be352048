"""Small window-attention forward + backward launches for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K, lib
lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
for B, H, ws, shift, nh in ((1, 24, 12, 6, 2), (2, 14, 7, 3, 1), (40, 24, 12, 0, 1)):
    C = nh * 32
    qkv = torch.randn(B * H * H, 3 * C, device=dev).to(torch.bfloat16)
    tab = torch.randn((2 * ws - 1) ** 2, nh, device=dev)
    win = (B, H, H, ws, shift)
    o, lse = K.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], nh, 32, 0.17, window=win, bias_table=tab)
    dqkv = torch.empty_like(qkv); dt = torch.zeros_like(tab)
    K.attn_bwd(o, qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, lse, nh, 32, 0.17, dqkv[:, :C], dqkv[:, C:2 * C],
               dqkv[:, 2 * C:], dbias_table=dt, window=win, bias_table=tab)
K.colsum(qkv); K.colsum(torch.randn(1000, 128, device=dev).to(torch.bfloat16)); K.colsum(torch.randn(333, 3072, device=dev).to(torch.bfloat16))
torch.cuda.synchronize()
print("sanitize winattn done", float(dqkv.float().abs().sum()), float(dt.abs().sum()))
