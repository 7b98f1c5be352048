"""Diagnostic: gradients of the scalar gates (alpha_i2t / alpha_t2i) — CUDA path vs fp32 oracle vs oracle under bf16 autocast,
for the hard-negative ITM objective at 224 px (the setting of tests/test_model_gpu.py::test_itm_hardneg_*)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from oracle import fiber_oracle as O, synth
from fiber_b200 import lib
from fiber_b200.modules import FIBERTransformerSS, objectives as OBJ
lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
cfg = bench.config(["itm", "mlm", "itc"], 224, 40)
model = FIBERTransformerSS(cfg)
shapes = {k: (tuple(v.shape), v.dtype) for k, v in model.state_dict().items() if not k.startswith("rank_output")}
sd = synth.synth_state_dict(shapes)
model.load_state_dict(sd, strict=False)
model.to(dev).train()
for m in model.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
    if hasattr(m, "drop_prob"): m.drop_prob = 0.0
sd = {k: v.to(dev) for k, v in sd.items()}
batch = bench.to_device(synth.synth_batch(2, 224, 40, seed=1234, false_image=True), dev, non_blocking=False)
ineg, tneg, mneg = batch["image"][0].roll(1, 0), batch["text_ids"].roll(1, 0), batch["text_masks"].roll(1, 0)
model.zero_grad()
OBJ.compute_itm_hardneg(model, dict(batch), ineg, tneg, mneg)["itm_loss"].backward()
def oracle(autocast):
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        loss = O.compute_itm_hardneg(sdg, cfg, batch, ineg, tneg, mneg)["itm_loss"]
    loss.backward()
    return sdg
r32, r16 = oracle(False), oracle(True)
print("%-58s %12s %12s %12s" % ("gate", "ours", "oracle fp32", "oracle bf16-autocast"))
for n, p in model.named_parameters():
    if n.endswith(("alpha_i2t", "alpha_t2i")) and p.grad is not None and r32[n].grad is not None:
        print("%-58s %12.4e %12.4e %12.4e" % (n, float(p.grad), float(r32[n].grad), float(r16[n].grad)))
