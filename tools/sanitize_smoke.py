"""Tiny invocation of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K, lib
lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
bf = torch.bfloat16
a = torch.randn(300, 128, device=dev).to(bf); w = torch.randn(264, 128, device=dev).to(bf)
bias = torch.randn(264, device=dev); res = torch.randn(300, 264, device=dev).to(bf)
pre = torch.empty(300, 264, device=dev, dtype=bf)
K.gemm(a, w, bias=bias, residual=res, preact=pre, act=K.ACT_GELU)
K.gemm(a, w, aux=res, act=K.ACT_GELU_GRAD)
K.gemm(res, a, mn_major=True, accumulate=True)
K.gemm(a, w, out_dtype=torch.float32)
B, H, ws, nh = 2, 14, 7, 2
C = nh * 32
qkv = torch.randn(B * H * H, 3 * C, device=dev).to(bf)
tab = torch.randn((2 * ws - 1) ** 2, nh, device=dev)
o, lse = K.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], nh, 32, 0.17, window=(B, H, H, ws, 3), bias_table=tab)
dqkv = torch.empty_like(qkv); dt = torch.zeros_like(tab)
K.attn_bwd(o, qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, lse, nh, 32, 0.17, dqkv[:, :C], dqkv[:, C:2 * C],
           dqkv[:, 2 * C:], dbias_table=dt, window=(B, H, H, ws, 3), bias_table=tab)
q = torch.randn(B * 40, 768, device=dev).to(bf); kv = torch.randn(B * 150, 1536, device=dev).to(bf)
o2, lse2 = K.attn_fwd(q, kv[:, :768], kv[:, 768:], 12, 64, 0.125, groups=B, lq=40, lk=150, drop_p=0.1, seed=3)
dq = torch.empty_like(q); dkv = torch.empty_like(kv)
K.attn_bwd(o2, q, kv[:, :768], kv[:, 768:], o2, lse2, 12, 64, 0.125, dq, dkv[:, :768], dkv[:, 768:], groups=B, lq=40,
           lk=150, drop_p=0.1, seed=3)
x = torch.randn(77, 512, device=dev).to(bf); g = torch.ones(512, device=dev); b = torch.zeros(512, device=dev)
y, m, r, _ = K.layernorm_fwd(x, g, b, 1e-5, add=x)
K.layernorm_bwd(y, x, m, r, g, add=x, dres=y, dgamma=torch.zeros_like(g), dbeta=torch.zeros_like(g))
xm = torch.randn(2 * 8 * 8, 64, device=dev).to(bf); g4 = torch.ones(256, device=dev); b4 = torch.zeros(256, device=dev)
ym, mm, rm, _ = K.layernorm_fwd(xm, g4, b4, 1e-5, merge=(2, 8, 8))
K.layernorm_bwd(ym, xm, mm, rm, g4, merge=(2, 8, 8), dgamma=torch.zeros_like(g4), dbeta=torch.zeros_like(g4))
K.colsum(x); K.dot(x, x); K.dropout(x, 0.1, 1); K.scale_rows(x, torch.ones(7, device=dev), 11); K.axpy(x, x, torch.ones(1, device=dev))
K.cast_bf16(torch.randn(1001, device=dev)); K.patch_gather(torch.randn(1, 3, 32, 32, device=dev))
ids = torch.randint(3, 50, (2, 10), device=dev); ids[0, 6:] = 1
wd = torch.randn(50, 64, device=dev); ps = torch.randn(14, 64, device=dev); ty = torch.randn(1, 64, device=dev)
e = K.embed_gather(ids, wd, ps, ty); K.embed_scatter(ids, e, torch.zeros_like(wd), torch.zeros_like(ps))
torch.cuda.synchronize()
print("sanitize smoke done, launches", lib.launch_count())
