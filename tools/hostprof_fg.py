"""Host-side profile (cProfile) of one fwd+bwd step of the fine-grained fused backbone."""
import cProfile, os, pstats, sys, time, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import lib
from fiber_b200.modules import fusion_swin_fg as M
lib.check(lib.load().fiber_init(), "init")
dev = torch.device("cuda:0")
B, Hi, Wi, L = 8, 800, 1344, 256
g = torch.Generator().manual_seed(77)
img = torch.randn(B, 3, Hi, Wi, generator=g).to(dev)
ids = torch.randint(3, 50265, (B, L), generator=g); ids[:, 0] = 0
mask = torch.ones(B, L, dtype=torch.long)
tok = {"input_ids": ids.to(dev), "attention_mask": mask.to(dev)}
torch.manual_seed(1234)
model = M.FusionSwinTransformer(M.SwinTransformer(drop_path_rate=0.2)).to(dev).train()

def step():
    for p in model.parameters():
        p.grad = None
    outs, lang, _x = model(tok, types.SimpleNamespace(tensors=img))
    loss = sum(o.float().mean() for o in outs) + lang["hidden"].float().mean()
    loss.backward()

for _ in range(2):
    step()
torch.cuda.synchronize()
n0 = lib.launch_count()
t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host time of one step %.1f ms, until GPU done %.1f ms, library launches %d" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3, lib.launch_count() - n0))
pr = cProfile.Profile(); pr.enable(); step(); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
st0 = torch.cuda.memory_stats()
step(); torch.cuda.synchronize()
st1 = torch.cuda.memory_stats()
for k in ("num_alloc_retries", "num_device_alloc", "num_device_free", "num_ooms", "allocation.all.allocated", "segment.all.allocated"):
    print(k, st1.get(k, 0) - st0.get(k, 0))
print("reserved GiB", torch.cuda.memory_reserved() / 2**30, "allocated GiB", torch.cuda.memory_allocated() / 2**30)
t0 = time.perf_counter()
xs = [torch.empty((548352, 384), dtype=torch.bfloat16, device=dev) for _ in range(20)]
print("20 x torch.empty(421 MB) with an idle GPU: %.2f ms each" % ((time.perf_counter() - t0) * 1e3 / 20))
