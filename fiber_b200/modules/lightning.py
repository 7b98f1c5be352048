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

    def batch_state(self, *a):
        """(numerator, denominator) contributed by one batch."""
        raise NotImplementedError

    def update(self, *a):
        num, den = self.batch_state(*a)
        self.acc += num
        self.total += den

    def forward(self, *a):
        """PL Metric.forward semantics (compute_on_step): accumulate the batch into the epoch state and return the value
        of THIS batch — what the objectives log per step; the epoch value comes from compute() in epoch_wrapup."""
        num, den = self.batch_state(*a)
        self.acc += num
        self.total += den
        return num / den


class Accuracy(Metric):  # gadgets/my_metrics.py:5-29
    def batch_state(self, logits, target):
        logits, target = logits.detach(), target.detach()
        # an integer tensor of the target's shape is taken as the predictions themselves (the fused MLM decoder +
        # cross-entropy returns the arg-max, not the logits)
        preds = logits if (not logits.is_floating_point() and logits.shape == target.shape) else logits.argmax(dim=-1)
        keep = target != -100
        # same sums as the reference's boolean-mask indexing, without its two device->host syncs per update
        # (the mask is applied arithmetically; an all-ignored batch adds 0 / 0 exactly like the early return)
        return ((preds == target) & keep).sum().float(), keep.sum().float()


class Scalar(Metric):  # gadgets/my_metrics.py:32-47
    def batch_state(self, scalar):
        v = scalar.detach().float() if isinstance(scalar, torch.Tensor) else torch.tensor(float(scalar), device=self.acc.device)
        return v, torch.ones((), device=self.acc.device)


class VQAScore(Metric):  # gadgets/my_metrics.py:50-69
    def batch_state(self, logits, target):
        logits, target = logits.detach().float(), target.detach().float()
        idx = logits.max(1)[1]
        one_hots = torch.zeros_like(target).scatter_(1, idx.view(-1, 1), 1)
        return (one_hots * target).sum(), torch.tensor(float(len(idx)), device=self.acc.device)
