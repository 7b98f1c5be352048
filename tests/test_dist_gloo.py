"""world_size-2 gloo test (CPU) of the only cross-rank host logic on the path besides DDP's
gradient all-reduce: concat_all_gather + the ITC queue update (fiber_module.py:12-24,181-222)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from fiber_b200.modules import FIBERTransformerSS
    from fiber_b200.modules.fiber_module import concat_all_gather
    g = concat_all_gather(torch.full((2, 3), float(rank)))
    assert g.shape == (4, 3) and g[:, 0].tolist() == [0, 0, 1, 1]
    m = FIBERTransformerSS(bench.config(["itc"], 32))
    n = 3
    f = torch.full((n, 768), float(rank + 1))
    m._dequeue_and_enqueue(f, -f, torch.full((n, 3, 32, 32), float(rank + 1)), torch.full((n, 40), rank + 1),
                           torch.ones(n, 40, dtype=torch.long))
    # every rank holds the identical queue: rank-0 entries then rank-1 entries
    assert int(m.queue_ptr) == 6 and int(m.queue_total) == 6
    assert m.image_queue[0, :6].tolist() == [1, 1, 1, 2, 2, 2]
    assert m.text_input_queue[:6, 0].tolist() == [1, 1, 1, 2, 2, 2]
    # DDP-style gradient averaging of an fp32 bucket (what the NCCL all-reduce does on the GPU)
    grad = torch.full((5,), float(rank + 1))
    dist.all_reduce(grad)
    assert (grad / world).tolist() == [1.5] * 5
    out.put(rank)
    dist.destroy_process_group()


def test_two_rank_gather_and_queue():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs)
    assert sorted(out.get() for _ in range(2)) == [0, 1]
