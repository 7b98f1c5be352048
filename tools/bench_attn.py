"""Window-attention kernel micro-benchmark at FIBER's four Swin stage shapes (384 px): forward and
backward time per launch, (window, head) problems per microsecond, and the implied mma.sync TFLOP/s.

    python tools/bench_attn.py [images]          # default 64 images

Parity is tests/test_attention_gpu.py's job; this only measures.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K, lib  # noqa: E402

lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
STAGES = [(96, 128, 4), (48, 256, 8), (24, 512, 16), (12, 1024, 32)]
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


print("images:", B, "| winattn_tc option:", lib.get_option("winattn_tc"), "(0 = mma.sync, 3 = tcgen05 + cp.async, 15 = tcgen05 + TMA quadrant tiles)")
for H, C, nh in STAGES:
    for shift in (0, 6):
        if H == 12 and shift:
            continue
        T = H * H
        g = torch.Generator(device="cpu").manual_seed(H + shift)
        qkv = torch.randn(B * T, 3 * C, generator=g).to(dev).to(torch.bfloat16)
        d_o = torch.randn(B * T, C, generator=g).to(dev).to(torch.bfloat16)
        table = (torch.randn(23 * 23, nh, generator=g) * 0.5).to(dev)
        win = (B, H, H, 12, shift)
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        o, lse = K.attn_fwd(q, k, v, nh, 32, 32 ** -0.5, window=win, bias_table=table)
        dqkv = torch.empty_like(qkv)
        dtab = torch.zeros_like(table)
        t_f = timeit(lambda: K.attn_fwd(q, k, v, nh, 32, 32 ** -0.5, window=win, bias_table=table))
        t_b = timeit(lambda: K.attn_bwd(d_o, q, k, v, o, lse, nh, 32, 32 ** -0.5, dqkv[:, :C], dqkv[:, C:2 * C],
                                        dqkv[:, 2 * C:], dbias_table=dtab, window=win, bias_table=table))
        wh = B * (H // 12) ** 2 * nh
        ff, fb = 4 * 144 * 144 * 32, 10 * 144 * 144 * 32
        print("H=%3d C=%4d nh=%2d shift=%d | fwd %7.3f ms %6.1f wh/us %6.1f TF/s | bwd %7.3f ms %6.1f wh/us %6.1f TF/s"
              % (H, C, nh, shift, t_f, wh / t_f / 1e3, wh * ff / t_f / 1e9, t_b, wh / t_b / 1e3, wh * fb / t_b / 1e9))
