"""Per-shape timing of the tcgen05 GEMM on FIBER's layer shapes (run on the GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K  # noqa: E402
from fiber_b200 import lib  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    dev = torch.device("cuda:0")
    lib.check(lib.load().fiber_init(), "init")
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    stages = [("s0", B * 9216, 128), ("s1", B * 2304, 256), ("s2", B * 576, 512), ("s3", B * 144, 1024)]
    print("%-28s %9s %6s %6s %9s %8s %8s" % ("case", "M", "N", "K", "us", "TFLOP/s", "GB/s"))
    flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)

    def report(name, M, N, K_, us, bytes_):
        print("%-28s %9d %6d %6d %9.1f %8.1f %8.0f" % (name, M, N, K_, us, 2.0 * M * N * K_ / us / 1e6, bytes_ / us / 1e3))

    for tag, M, C in stages:
        x = torch.randn(M, C, device=dev).to(torch.bfloat16)
        x4 = torch.randn(M, 4 * C, device=dev).to(torch.bfloat16)
        res = torch.randn(M, C, device=dev).to(torch.bfloat16)
        for name, a, N, kw in (
            ("qkv", x, 3 * C, {}),
            ("proj+res", x, C, {"residual": res}),
            ("fc1+gelu+preact", x, 4 * C, {"act": K.ACT_GELU, "preact": True}),
            ("fc1+gelu+gelu' (act3)", x, 4 * C, {"act": K.ACT_GELU_CACHE, "preact": True}),
            ("fc2+res", x4, C, {"residual": res}),
            ("plain N=4C", x, 4 * C, {}),
        ):
            Kd = a.shape[1]
            w = torch.randn(N, Kd, device=dev).to(torch.bfloat16)
            bias = torch.randn(N, device=dev)
            kw = dict(kw)
            extra = 0
            if kw.get("preact"):
                kw["preact"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
                extra += M * N * 2
            if "residual" in kw:
                extra += M * N * 2
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            us = timeit(lambda: K.gemm(a, w, bias=bias if name != "plain N=4C" else None, out=out, **kw))
            report(tag + " " + name, M, N, Kd, us, M * Kd * 2 + N * Kd * 2 + M * N * 2 + extra)
        # fc2 dgrad through the cached GELU' (act 4): dh[M, 4C] = (dz[M, C] W2^T) * aux
        w2 = torch.randn(4 * C, C, device=dev).to(torch.bfloat16)
        out4 = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
        us = timeit(lambda: K.gemm(x, w2, aux=x4, act=K.ACT_MUL_AUX, out=out4))
        report(tag + " fc2 dgrad * aux (act4)", M, 4 * C, C, us, M * C * 2 + 4 * C * C * 2 + 2 * M * 4 * C * 2)
        del w2, out4
        # wgrad: dW[N, K] = dY[M, N]^T X[M, K]
        for name, N, Kd in (("wgrad qkv", 3 * C, C), ("wgrad fc1", 4 * C, C), ("wgrad fc2", C, 4 * C)):
            dy = torch.randn(M, N, device=dev).to(torch.bfloat16)
            xx = x if Kd == C else x4
            us = timeit(lambda: K.gemm(dy, xx, mn_major=True, accumulate=True))
            report(tag + " " + name, N, Kd, M, us, M * N * 2 + M * Kd * 2)
        del x, x4, res
    M = B * 40
    x = torch.randn(M, 768, device=dev).to(torch.bfloat16)
    for name, N, Kd in (("text qkv", 2304, 768), ("text dense", 768, 768), ("text fc1", 3072, 768)):
        w = torch.randn(N, Kd, device=dev).to(torch.bfloat16)
        us = timeit(lambda: K.gemm(x, w))
        report(name, M, N, Kd, us, M * Kd * 2 + N * Kd * 2 + M * N * 2)
    # square reference point
    a = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
    b = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
    us = timeit(lambda: K.gemm(a, b), 5)
    report("square 8192^3", 8192, 8192, 8192, us, 3 * 8192 * 8192 * 2)
    us = timeit(lambda: torch.matmul(a, b.t()), 5)
    report("cuBLAS 8192^3 (reference)", 8192, 8192, 8192, us, 3 * 8192 * 8192 * 2)


if __name__ == "__main__":
    main()
