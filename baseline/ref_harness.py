"""Drive the UNMODIFIED reference (baseline/_ref/fiber, or /root/reference in the build container)
through its own public API — FIBERTransformerSS(config).training_step(batch, 0) + backward — for
bench.py's reference arms.  None of this repo's kernels, modules or objectives are on that path.

    step = RefStep(tasks, image_size, L, device, autocast=torch.bfloat16 | None)
    loss = step(batch)          # fiber_utils.set_task -> forward -> objectives.compute_* -> sum of losses -> backward
"""
import os

import torch

from . import ref_shims


def available():
    return ref_shims.available()


def ensure_group(device):
    """objectives.compute_itc -> _dequeue_and_enqueue -> concat_all_gather (fiber_module.py:12-24) needs a process
    group even on one GPU: a private 1-rank group (HashStore: no port, no rendezvous with other ranks)."""
    import torch.distributed as dist
    if dist.is_initialized():
        return False
    backend = "nccl" if device.type == "cuda" else "gloo"
    dist.init_process_group(backend, store=dist.HashStore(), rank=0, world_size=1)
    return True


def fill_queues(model, seed=4321):
    """Steady-state ITC queues (the state every run reaches after 4096 / global-batch steps): unit-norm feature
    columns, N(0,1) images, valid token ids, queue_total = queue_size.  Works on the reference module and on
    fiber_b200's (same buffer names, fiber_module.py:61-70)."""
    if not hasattr(model, "image_queue"):
        return
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for q in (model.image_queue, model.text_queue):
            v = torch.randn(q.shape, generator=g)
            q.copy_(v / v.norm(dim=0, keepdim=True))
        n, L = model.text_input_queue.shape
        ids = torch.randint(3, 50264, (n, L), generator=g)
        ids[:, 0], ids[:, -1] = 0, 2
        model.text_input_queue.copy_(ids)
        model.text_input_mask_queue.fill_(1)
        model.queue_total.fill_(n)
        model.queue_ptr.fill_(0)


class RefStep:
    def __init__(self, tasks, image_size, L, device, autocast=None, full_queue=True):
        ref_shims.install()
        from fiber.modules import FIBERTransformerSS  # the reference's class, unmodified
        self.device = torch.device(device)
        self.autocast = autocast
        cfg = ref_shims.default_config(tasks=list(tasks), image_size=image_size, max_text_len=L)
        torch.manual_seed(1234)
        model = FIBERTransformerSS(cfg)
        with torch.no_grad():  # same gate values as our arm: the cross-attention branches carry signal
            for n, p in model.named_parameters():
                if n.endswith(("alpha_i2t", "alpha_t2i")):
                    p.fill_(0.5)
        if full_queue:
            fill_queues(model)
        self.model = model.to(self.device).train()
        self.owns_group = ensure_group(self.device) if "itc" in tasks else False

    def __call__(self, batch):
        m = self.model
        batch = dict(batch)  # the reference's objectives rebind keys of the dict they are given (objectives.py:88-95)
        for p in m.parameters():
            p.grad = None
        if self.autocast is not None:
            with torch.autocast(self.device.type, dtype=self.autocast):
                loss = m.training_step(batch, 0)
        else:
            loss = m.training_step(batch, 0)
        loss.backward()
        return loss

    def close(self):
        self.model = None
        if self.owns_group:
            torch.distributed.destroy_process_group()


def cpu_threads():
    return int(os.environ.get("FIBER_REF_THREADS", os.cpu_count() or 1))
