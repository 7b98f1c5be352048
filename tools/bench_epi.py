import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K, lib
from tools.bench_gemm import timeit
lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
M, N, Kd = 147456, 1024, 256
x = torch.randn(M, Kd, device=dev).to(torch.bfloat16)
w = (torch.randn(N, Kd, device=dev) * 0.05).to(torch.bfloat16)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
bias = torch.randn(N, device=dev)
pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
aux = torch.randn(M, N, device=dev).to(torch.bfloat16)
rs = torch.rand(M // 2304, device=dev)
for name, kw in (("plain", {}), ("bias", {"bias": bias}), ("gelu", {"act": K.ACT_GELU}), ("preact", {"preact": pre}),
                 ("bias+gelu", {"bias": bias, "act": K.ACT_GELU}), ("gelu+preact", {"act": K.ACT_GELU, "preact": pre}),
                 ("gelugrad", {"act": K.ACT_GELU_GRAD, "aux": aux}), ("rowscale", {"row_scale": rs, "rows_per_scale": 2304}),
                 ("all", {"bias": bias, "act": K.ACT_GELU, "preact": pre})):
    us = timeit(lambda: K.gemm(x, w, out=out, **kw))
    print("%-14s %8.1f us" % (name, us))
