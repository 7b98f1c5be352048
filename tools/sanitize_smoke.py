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
# round 2: single-pass GELU epilogues (two store boxes), aux / residual epilogues with the coalesced loader, small-key
# tcgen05 attention forward (with and without dropout), fused AdamW
a2 = torch.randn(256, 128, device=dev).to(bf); w2 = torch.randn(288, 128, device=dev).to(bf)
aux2 = torch.randn(256, 288, device=dev).to(bf); pre2 = torch.empty(256, 288, device=dev, dtype=bf)
K.gemm(a2, w2, bias=torch.randn(288, device=dev), act=K.ACT_GELU_CACHE, preact=pre2)
K.gemm(a2, w2, aux=aux2, act=K.ACT_MUL_AUX)
K.gemm(a2, w2, bias=torch.randn(288, device=dev), residual=aux2, act=K.ACT_RES_PF, row_scale=torch.ones(2, device=dev), rows_per_scale=128)
# tcgen05 plain attention, forward + backward: i2t (many queries), unpacked short, packed self-attention (with dropout)
for nhh, hd, lq, lk, dp in ((2, 32, 200, 40, 0.0), (2, 64, 130, 50, 0.2), (2, 64, 40, 40, 0.1), (2, 64, 33, 47, 0.0)):
    Cc = nhh * hd
    qq = torch.randn(4 * lq, Cc, device=dev).to(bf); kk = torch.randn(4 * lk, 2 * Cc, device=dev).to(bf)
    dd = torch.randn(4 * lq, Cc, device=dev).to(bf)
    mk = torch.zeros(4, lk, device=dev); mk[1, lk - 5:] = -10000.0
    lib.set_option("attn_sk", 31)
    kw2 = dict(groups=4, lq=lq, lk=lk, key_mask=mk, drop_p=dp, seed=5)
    oo, ll = K.attn_fwd(qq, kk[:, :Cc], kk[:, Cc:], nhh, hd, hd ** -0.5, **kw2)
    dq2, dkv2 = torch.empty_like(qq), torch.empty_like(kk)
    K.attn_bwd(dd, qq, kk[:, :Cc], kk[:, Cc:], oo, ll, nhh, hd, hd ** -0.5, dq2, dkv2[:, :Cc], dkv2[:, Cc:], **kw2)
    lib.set_option("attn_sk", -1)
# CTA-pair GEMM (2-CTA clusters) and the fused MLM decoder + cross-entropy
a3 = torch.randn(512, 1024, device=dev).to(bf); w3 = torch.randn(384, 1024, device=dev).to(bf)
K.gemm(a3, w3, bias=torch.randn(384, device=dev), residual=torch.randn(512, 384, device=dev).to(bf))
from fiber_b200 import ops
hh = torch.randn(200, 768, device=dev).to(bf).requires_grad_(True)
ww = (torch.randn(1000, 768, device=dev) * 0.05).requires_grad_(True); bb = torch.zeros(1000, device=dev, requires_grad=True)
lab = torch.randint(0, 1000, (200,), device=dev); lab[::3] = -100
loss_ce, _ = ops.MlmDecoderCEFn.apply(hh, ww, bb, lab)
loss_ce.backward()
from fiber_b200.optim import FusedAdamW
ps = [torch.nn.Parameter(torch.randn(n, device=dev)) for n in (5, 1023, 70000)]
for p_ in ps:
    p_.grad = torch.randn_like(p_)
FusedAdamW(ps, lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01).step()
torch.cuda.synchronize()
print("sanitize smoke done, launches", lib.launch_count())
