"""Kernel timeline of one warm fwd+bwd step of a bench.py extra config (vqa576 / itc384).
    python tools/profile_extra.py vqa576 out.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fiber_b200 import lib, ops  # noqa: E402
from fiber_b200.modules import FIBERTransformerSS, fiber_utils  # noqa: E402

name, out = sys.argv[1], sys.argv[2]
dev = torch.device("cuda:0")
lib.check(lib.load().fiber_init(), "init")
tasks, R, L, B, flops = bench.EXTRA_CONFIGS[name]
torch.manual_seed(1234)
model = FIBERTransformerSS(bench.config(tasks, R, L)).to(dev)
with torch.no_grad():
    for n, p in model.named_parameters():
        if n.endswith(("alpha_i2t", "alpha_t2i")):
            p.fill_(0.5)
bench.fill_queues(model)
model.train()
fiber_utils.set_task(model)
ops.set_dropout_seed(1234)
batch = bench.make_batch(B, R, L, seed=1234)
if "vqa" in tasks:
    g = torch.Generator(device="cpu").manual_seed(99)
    batch["vqa_labels"] = [[int(torch.randint(0, 3129, (1,), generator=g))] for _ in range(B)]
    batch["vqa_scores"] = [[1.0] for _ in range(B)]
batch = bench.to_device(batch, dev, non_blocking=False)


def step(b):
    for p in model.parameters():
        p.grad = None
    o = model(b)
    loss = sum(v for k, v in o.items() if "loss" in k)
    loss.backward()
    return loss


for _ in range(3):
    step(batch)
torch.cuda.synchronize()
bench.profile_timeline(step, batch, out)
print(open(out).read())
