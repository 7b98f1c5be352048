import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K, lib
lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
M, N, Kd = 147456, 512, 128
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
x = torch.randn(M, Kd, device=dev).to(torch.bfloat16)
w = torch.randn(N, Kd, device=dev).to(torch.bfloat16)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
bias = torch.randn(N, device=dev)
res = torch.randn(M, N, device=dev).to(torch.bfloat16)
pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    if mode == "plain":
        K.gemm(x, w, out=out)
    else:
        K.gemm(x, w, bias=bias, residual=res, preact=pre, act=K.ACT_GELU, out=out)
torch.cuda.synchronize()
