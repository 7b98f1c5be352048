"""Use pytorch_lightning when it is installed; otherwise a minimal stand-in that provides what
FIBERTransformerSS and the objectives touch (hparams.config, log, device, trainer, global_step)."""
import torch
import torch.nn as nn

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as pl
    LightningModule = pl.LightningModule
    HAVE_PL = True
except Exception:  # noqa: BLE001
    HAVE_PL = False

    class _HParams(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

    class LightningModule(nn.Module):
        def __init__(self):
            super().__init__()
            self.hparams = _HParams()
            self.trainer = None
            self.global_step = 0
            self.logged = {}

        def save_hyperparameters(self, **kw):
            self.hparams.update(kw)

        @property
        def device(self):
            return next(self.parameters()).device

        def log(self, name, value, *a, **k):
            self.logged[name] = value.detach() if isinstance(value, torch.Tensor) else value


class Metric(nn.Module):
    """Tiny accumulate-and-average metric (pytorch_lightning.metrics.Metric stand-in)."""

    def __init__(self):
        super().__init__()
        self.register_buffer("acc", torch.tensor(0.0), persistent=False)
        self.register_buffer("total", torch.tensor(0.0), persistent=False)

    def compute(self):
        return self.acc / self.total

    def reset(self):
        self.acc.zero_()
        self.total.zero_()

    def forward(self, *a):
        self.update(*a)
        return self.compute()


class Accuracy(Metric):  # gadgets/my_metrics.py:5-29
    def update(self, logits, target):
        logits, target = logits.detach(), target.detach()
        preds = logits.argmax(dim=-1)
        keep = target != -100
        # same sums as the reference's boolean-mask indexing, without its two device->host syncs per update
        # (the mask is applied arithmetically; an all-ignored batch adds 0 / 0 exactly like the early return)
        self.acc += ((preds == target) & keep).sum()
        self.total += keep.sum()


class Scalar(Metric):  # gadgets/my_metrics.py:32-47
    def update(self, scalar):
        self.acc += scalar.detach().float() if isinstance(scalar, torch.Tensor) else float(scalar)
        self.total += 1


class VQAScore(Metric):  # gadgets/my_metrics.py:50-69
    def update(self, logits, target):
        logits, target = logits.detach().float(), target.detach().float()
        idx = logits.max(1)[1]
        one_hots = torch.zeros_like(target).scatter_(1, idx.view(-1, 1), 1)
        self.acc += (one_hots * target).sum()
        self.total += len(idx)
