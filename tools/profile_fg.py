"""Kernel timeline of one fwd+bwd step of the fine-grained fused backbone (bench.py fg800 extra config).
    python tools/profile_fg.py out.txt [B H W L]"""
import os, sys, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fiber_b200 import lib
from fiber_b200.modules import fusion_swin_fg as M
lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
out = sys.argv[1]
B, Hi, Wi, L = [int(x) for x in sys.argv[2:6]] if len(sys.argv) > 5 else (8, 800, 1344, 256)
g = torch.Generator().manual_seed(77)
img = torch.randn(B, 3, Hi, Wi, generator=g).to(dev)
ids = torch.randint(3, 50265, (B, L), generator=g); ids[:, 0] = 0
mask = torch.ones(B, L, dtype=torch.long)
for b in range(B):
    n = int(torch.randint(20, 60, (1,), generator=g)); ids[b, n - 1] = 2; ids[b, n:] = 1; mask[b, n:] = 0
tok = {"input_ids": ids.to(dev), "attention_mask": mask.to(dev)}
torch.manual_seed(1234)
model = M.FusionSwinTransformer(M.SwinTransformer(drop_path_rate=0.2)).to(dev).train()
with torch.no_grad():
    for n, p in model.named_parameters():
        if n.endswith(("alpha_i2t", "alpha_t2i")):
            p.fill_(0.5)

def step(_=None):
    for p in model.parameters():
        p.grad = None
    outs, lang, _x = model(tok, types.SimpleNamespace(tensors=img))
    loss = sum(o.float().mean() for o in outs) + lang["hidden"].float().mean()
    loss.backward()

for _ in range(2):
    step()
torch.cuda.synchronize()
bench.profile_timeline(step, None, out)
print(open(out).read()[:6000])
