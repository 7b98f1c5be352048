"""Plain-attention micro-benchmark at FIBER's shapes (4B-sample pass of config 1): text self-attention, i2t at stages 2 / 3,
t2i.  Forward and backward time per launch for the current "attn_sk" option.
    python tools/bench_attn_plain.py [samples]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K, lib  # noqa: E402

lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


print("samples:", B, "| attn_sk option:", lib.get_option("attn_sk"))
for name, nh, hd, Lq, Lk, masked, drop in (("text self (L=40)", 12, 64, 40, 40, True, 0.1), ("i2t stage 2", 16, 32, 576, 40, True, 0.0),
                                           ("i2t stage 3", 32, 32, 144, 40, True, 0.0), ("t2i stage 2", 12, 64, 40, 576, False, 0.1)):
    C = nh * hd
    g = torch.Generator(device="cpu").manual_seed(Lq + Lk)
    q = torch.randn(B * Lq, C, generator=g).to(dev).to(torch.bfloat16)
    kv = torch.randn(B * Lk, 2 * C, generator=g).to(dev).to(torch.bfloat16)
    d_o = torch.randn(B * Lq, C, generator=g).to(dev).to(torch.bfloat16)
    mask = None
    if masked:
        mask = torch.zeros(B, Lk, device=dev)
        mask[::2, Lk - 9:] = -10000.0
    kw = dict(groups=B, lq=Lq, lk=Lk, key_mask=mask, drop_p=drop, seed=7)
    scale = 1.0 / math.sqrt(hd)
    o, lse = K.attn_fwd(q, kv[:, :C], kv[:, C:], nh, hd, scale, **kw)
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    t_f = timeit(lambda: K.attn_fwd(q, kv[:, :C], kv[:, C:], nh, hd, scale, **kw))
    t_b = timeit(lambda: K.attn_bwd(d_o, q, kv[:, :C], kv[:, C:], o, lse, nh, hd, scale, dq, dkv[:, :C], dkv[:, C:], **kw))
    by = (2 * q.numel() + kv.numel()) * 2
    print("%-18s nh=%2d hd=%2d Lq=%4d Lk=%4d | fwd %7.3f ms (%5.0f GB/s) | bwd %7.3f ms" % (name, nh, hd, Lq, Lk, t_f, by / t_f / 1e6, t_b))
