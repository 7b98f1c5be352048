import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K, lib
from tools.bench_gemm import timeit
lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
for rows, C in ((589824, 128), (147456, 256), (36864, 512), (9216, 1024), (2560, 768)):
    x = torch.randn(rows, C, device=dev).to(torch.bfloat16)
    dy = torch.randn(rows, C, device=dev).to(torch.bfloat16)
    g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    y, mean, rstd, _ = K.layernorm_fwd(x, g, b, 1e-5)
    dg, db = torch.zeros_like(g), torch.zeros_like(g)
    t1 = timeit(lambda: K.layernorm_fwd(x, g, b, 1e-5))
    t2 = timeit(lambda: K.layernorm_bwd(dy, x, mean, rstd, g, dres=dy, dgamma=dg, dbeta=db))
    print("rows %7d C %4d  fwd %7.1f us %5.2f TB/s   bwd(+dres) %7.1f us %5.2f TB/s" %
          (rows, C, t1, rows * C * 4 / t1 / 1e6, t2, rows * C * 8 / t2 / 1e6))

for rows, N in ((589824, 128), (589824, 512), (147456, 1024), (36864, 2048), (36864, 512), (9216, 4096)):
    x = torch.randn(rows, N, device=dev).to(torch.bfloat16)
    out = torch.zeros(N, device=dev)
    t = timeit(lambda: K.colsum(x, out=out))
    print("colsum rows %7d N %4d  %7.1f us %5.2f TB/s" % (rows, N, t, rows * N * 2 / t / 1e6))
