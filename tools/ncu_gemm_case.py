"""A few launches of one GEMM shape / epilogue for an `ncu --set full` capture.
    python tools/ncu_gemm_case.py MODE [M N K]      MODE in plain | bias | gelu | gelu_grad | res | wgrad | gelu_cache | mul_aux | res_pf"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K, lib
lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
M, N, Kd = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (147456, 2048, 512)
x = torch.randn(M, Kd, device=dev).to(torch.bfloat16)
w = torch.randn(N, Kd, device=dev).to(torch.bfloat16)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
bias = torch.randn(N, device=dev)
res = torch.randn(M, N, device=dev).to(torch.bfloat16)
pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
dw = torch.zeros(N, Kd, device=dev)
db = torch.zeros(N, device=dev)
for _ in range(3):
    if mode == "plain":
        K.gemm(x, w, out=out)
    elif mode == "bias":
        K.gemm(x, w, bias=bias, out=out)
    elif mode == "res":
        K.gemm(x, w, bias=bias, residual=res, out=out)
    elif mode == "gelu":
        K.gemm(x, w, bias=bias, preact=pre, act=K.ACT_GELU, out=out)
    elif mode == "gelu_grad":
        K.gemm(x, w, aux=res, act=K.ACT_GELU_GRAD, out=out)
    elif mode == "gelu_cache":   # act 3: GELU + GELU' in one pass (the default fc1 epilogue)
        K.gemm(x, w, bias=bias, preact=pre, act=K.ACT_GELU_CACHE, out=out)
    elif mode == "mul_aux":      # act 4: acc * aux (the default fc2-dgrad epilogue)
        K.gemm(x, w, aux=res, act=K.ACT_MUL_AUX, out=out)
    elif mode == "res_pf":       # act 6: residual epilogue on the two-box path
        K.gemm(x, w, bias=bias, residual=res, act=K.ACT_RES_PF, out=out)
    elif mode == "wgrad":  # dW[N, Kd] = dY[M, N]^T X[M, Kd] with the fused bias gradient
        K.gemm(res, x, mn_major=True, accumulate=True, out=dw, colsum=db)
torch.cuda.synchronize()
