"""Fused multi-tensor AdamW: the optimizer of fiber_utils.set_schedule
(coarse_grained/fiber/modules/fiber_utils.py:156-252 — transformers.AdamW over six name-selected parameter
groups, betas (0.9, 0.98), eps 1e-8) as ONE kernel launch per step (csrc/optim.cu, fiber_adamw_multi).

HF 4.6 semantics, bit for bit in operation order: m, v updated; p -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps);
then p -= lr * wd * p.  Parameter groups survive as two scalars per tensor (lr, weight_decay), so LR schedulers that
rewrite group["lr"] keep working.  There is no CPU path: stepping CPU parameters raises."""
import ctypes as C

import numpy as np
import torch

from . import lib as _lib

CHUNK = 65536  # elements per CTA (a multiple of 1024)

_TENSOR_DTYPE = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("p_bf16", "<u8"), ("n", "<i8"),
                          ("lr", "<f4"), ("wd", "<f4")])
assert _TENSOR_DTYPE.itemsize == 56  # struct fiber_adamw_tensor


def build_tables(entries, chunk=CHUNK):
    """entries: iterable of (p_ptr, g_ptr, m_ptr, v_ptr, bf16_ptr_or_0, numel, lr, wd) -> (tensor table, chunk table) as
    numpy arrays with the layouts of include/fiber_b200.h (host logic, testable without a GPU)."""
    entries = list(entries)
    t = np.zeros(len(entries), dtype=_TENSOR_DTYPE)
    chunks = []
    for i, (p, g, m, v, b, n, lr, wd) in enumerate(entries):
        t[i] = (p, g, m, v, b, n, lr, wd)
        chunks.extend((i, c) for c in range((n + chunk - 1) // chunk))
    return t, np.asarray(chunks, dtype=np.int32).reshape(-1, 2)


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._step = 0
        self._stage = [None, None]  # two pinned staging buffers (+ the event of their last copy), used alternately

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        entries, keep = [], []
        betas = eps = None
        dev = None
        for group in self.param_groups:
            if betas is None:
                betas, eps = group["betas"], group["eps"]
            elif (betas, eps) != (group["betas"], group["eps"]):
                raise RuntimeError("FusedAdamW: betas / eps must be the same in every group (they are in FIBER's)")
            for p in group["params"]:
                if p.grad is None:
                    continue  # e.g. the grad-less parameters of a task mix
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdamW: parameters must be contiguous fp32 CUDA tensors (no CPU path exists)")
                g = p.grad if (p.grad.dtype == torch.float32 and p.grad.is_contiguous()) else p.grad.float().contiguous()
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                keep.append(g)
                dev = p.device
                entries.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), 0,
                                p.numel(), group["lr"], group["weight_decay"]))
        if not entries:
            return loss
        self._step += 1
        t, ch = build_tables(entries)
        raw = torch.from_numpy(np.concatenate([t.view(np.uint8).reshape(-1), ch.view(np.uint8).reshape(-1)]))
        # ~50 KB of tables per step (the scheduler rewrites lr).  A pageable .to(device) makes the host wait for the whole
        # stream — i.e. for the step's backward — every step, and the next step's first launches then start on an idle GPU
        # (+14..20 ms per step measured).  Stage through pinned memory instead, two buffers used alternately: a buffer is
        # rewritten only after the copy issued from it two steps ago has completed.
        nbytes = raw.numel()
        slot = self._step & 1
        stage = self._stage[slot]
        if stage is None or stage[0].numel() < nbytes:
            stage = (torch.empty(nbytes, dtype=torch.uint8, pin_memory=True), torch.cuda.Event())
            self._stage[slot] = stage
        else:
            stage[1].synchronize()
        stage[0][:nbytes].copy_(raw)
        d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        d.copy_(stage[0][:nbytes], non_blocking=True)
        stage[1].record(torch.cuda.current_stream())
        off = t.nbytes
        _lib.check(_lib.load().fiber_adamw_multi(C.c_void_p(d.data_ptr()), C.c_void_p(d.data_ptr() + off), len(ch), CHUNK,
                                                 float(betas[0]), float(betas[1]), float(eps), self._step,
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)), "adamw_multi")
        d.record_stream(torch.cuda.current_stream())
        return loss
