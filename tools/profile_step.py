"""Kernel timeline of ONE steady-state bench step via torch.profiler (CUPTI): per-kernel totals in real
(non-serialised, warm) conditions, the sum of kernel time vs the step's wall time, and the idle gaps
between consecutive kernels on the compute stream.

    python tools/profile_step.py [out.txt]
"""
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fiber_b200 import lib, ops  # noqa: E402
from fiber_b200.modules import FIBERTransformerSS, fiber_utils  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else None
dev = torch.device("cuda:0")
lib.check(lib.load().fiber_init(), "init")
B, R, L = 64, 384, 40
torch.manual_seed(1234)
model = FIBERTransformerSS(bench.config(["itm", "itc", "mlm"], R, L)).to(dev)
with torch.no_grad():
    for n, p in model.named_parameters():
        if n.endswith(("alpha_i2t", "alpha_t2i")):
            p.fill_(0.5)
bench.fill_queues(model)
model.train()
fiber_utils.set_task(model)
ops.set_dropout_seed(1234)
torch.backends.cuda.matmul.allow_tf32 = True
batch = bench.to_device(bench.make_batch(B, R, L), dev)


def step():
    for p in model.parameters():  # same order as bench.py: gradients cleared while the GPU is still busy
        p.grad = None
    out = model(batch)
    loss = sum(v for k, v in out.items() if "loss" in k)
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
agg = defaultdict(lambda: [0, 0.0])
for s, e, n in ks:
    n = n.split("(")[0].replace("void ", "")
    agg[n][0] += 1
    agg[n][1] += e - s
t0, t1 = ks[0][0], max(k[1] for k in ks)
busy = 0.0
gaps = []
big = []
cur_end = ks[0][0]
prev = ""
for s, e, n in ks:
    if s > cur_end:
        gaps.append(s - cur_end)
        big.append((s - cur_end, prev, n, (s - t0) / 1e3))
    if e > cur_end:
        busy += e - max(s, cur_end)
        cur_end = e
    prev = n
lines = ["span %.3f ms, %d kernels, busy %.3f ms (%.1f%%), idle %.3f ms in %d gaps (median %.2f us, >10us: %d totalling %.3f ms)"
         % ((t1 - t0) / 1e3, len(ks), busy / 1e3, 100 * busy / (t1 - t0), sum(gaps) / 1e3, len(gaps),
            sorted(gaps)[len(gaps) // 2] if gaps else 0, sum(g > 10 for g in gaps), sum(g for g in gaps if g > 10) / 1e3)]
lines.append("%-86s %6s %10s %7s %9s" % ("kernel", "count", "ms", "share", "avg us"))
tot = sum(v[1] for v in agg.values())
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    lines.append("%-86s %6d %10.3f %6.1f%% %9.1f" % (n[:86], c, us / 1e3, 100 * us / tot, us / c))
lines.append("largest idle gaps: us | at ms | after kernel -> before kernel")
for gap, a, b, at in sorted(big, reverse=True)[:25]:
    lines.append("%8.1f | %7.2f | %s -> %s" % (gap, at, a.split("(")[0][-60:], b.split("(")[0][-60:]))
text = "\n".join(lines)
print(text)
if out_path:
    with open(out_path, "w") as f:
        f.write(text + "\n")
