"""A few LayerNorm forward / backward launches for an `ncu --set full` capture.
    python tools/ncu_ln_case.py ROWS C"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K, lib
lib.check(lib.load().fiber_init(), "init")
rows, C = (int(a) for a in sys.argv[1:3]) if len(sys.argv) > 2 else (147456, 512)
dev = torch.device("cuda:0")
x = torch.randn(rows, C, device=dev).to(torch.bfloat16)
dy = torch.randn(rows, C, device=dev).to(torch.bfloat16)
dres = torch.randn(rows, C, device=dev).to(torch.bfloat16)
g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
for _ in range(3):
    y, m, r, _ = K.layernorm_fwd(x, g, b, 1e-5)
    K.layernorm_bwd(dy, x, m, r, g, dres=dres, dgamma=torch.zeros_like(g), dbeta=torch.zeros_like(g))
torch.cuda.synchronize()
