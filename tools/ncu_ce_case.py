"""Fused MLM decoder + cross-entropy forward / backward at the bench shape (2560 rows, 15 % labelled) for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import ops, lib
lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
rows, V = 2560, 50265
h = (torch.randn(rows, 768, device=dev) * 1.5).to(torch.bfloat16).requires_grad_(True)
w = (torch.randn(V, 768, device=dev) * 0.05).requires_grad_(True)
b = torch.zeros(V, device=dev, requires_grad=True)
lab = torch.randint(0, V, (rows,), device=dev); lab[torch.rand(rows, device=dev) > 0.15] = -100
for _ in range(2):
    loss, _ = ops.MlmDecoderCEFn.apply(h, w, b, lab)
    loss.backward()
torch.cuda.synchronize()
print("done", float(loss))
