"""ONE steady-state bench step between cudaProfilerStart / Stop, for the ncu launch list:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python tools/ncu_step.py
(the same model, batch, seeds and step as bench.py / tools/profile_step.py; three warm steps run unprofiled)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fiber_b200 import lib, ops  # noqa: E402
from fiber_b200.modules import FIBERTransformerSS, fiber_utils  # noqa: E402

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
    for p in model.parameters():
        p.grad = None
    out = model(batch)
    loss = sum(v for k, v in out.items() if "loss" in k)
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
n0 = lib.launch_count()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step:", lib.launch_count() - n0, "launches of this library")
