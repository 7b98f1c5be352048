"""Mirror of coarse_grained/fiber/modules/fiber_utils.py for the parts a training step needs:
metric registration, task selection and the optimizer's name-substring parameter groups."""
import torch

from .lightning import Accuracy, Scalar, VQAScore


def set_metrics(pl_module):  # fiber_utils.py:14-41
    for split in ["train", "val"]:
        for k, v in pl_module.hparams.config["loss_names"].items():
            if v <= 0:
                continue
            if k == "vqa":
                setattr(pl_module, f"{split}_vqa_score", VQAScore())
                setattr(pl_module, f"{split}_{k}_loss", Scalar())
            elif k == "itc":
                setattr(pl_module, f"{split}_{k}_i2t_accuracy", Accuracy())
                setattr(pl_module, f"{split}_{k}_t2i_accuracy", Accuracy())
                setattr(pl_module, f"{split}_{k}_loss", Scalar())
                setattr(pl_module, f"{split}_{k}_logit_scale", Scalar())
            else:
                setattr(pl_module, f"{split}_{k}_accuracy", Accuracy())
                setattr(pl_module, f"{split}_{k}_loss", Scalar())


# metrics that make up "<phase>/the_metric" (fiber_utils.py:44-140): (attribute suffix, logged name, counts towards the metric)
_EPOCH_METRICS = {
    "vqa": (("vqa_score", "score_epoch", True), ("vqa_loss", "loss_epoch", False)),
    "itm": (("itm_accuracy", "accuracy_epoch", True), ("itm_loss", "loss_epoch", False)),
    "mlm": (("mlm_accuracy", "accuracy_epoch", True), ("mlm_loss", "loss_epoch", False)),
    # the reference sums only the LAST value it computed for itc, the t2i accuracy (:118-131)
    "itc": (("itc_i2t_accuracy", "i2t_accuracy_epoch", False), ("itc_t2i_accuracy", "t2i_accuracy_epoch", True),
            ("itc_loss", "loss_epoch", False)),
}


def epoch_wrapup(pl_module):
    """fiber_utils.py:44-140 for the tasks in scope: log every epoch metric, reset it, and log '<phase>/the_metric'
    (what run.py's ModelCheckpoint monitors).  Retrieval recall (get_recall_metric) is an evaluation pass of its own
    and is out of scope here."""
    phase = "train" if pl_module.training else "val"
    the_metric = 0
    for loss_name, v in pl_module.hparams.config["loss_names"].items():
        if v <= 0 or loss_name not in _EPOCH_METRICS:
            continue
        for attr, logged, counts in _EPOCH_METRICS[loss_name]:
            metric = getattr(pl_module, f"{phase}_{attr}")
            value = metric.compute()
            pl_module.log(f"{loss_name}/{phase}/{logged}", value)
            metric.reset()
            if counts:
                the_metric = the_metric + value
    pl_module.log(f"{phase}/the_metric", the_metric)


def set_task(pl_module):  # fiber_utils.py:151-153
    pl_module.current_tasks = [k for k, v in pl_module.hparams.config["loss_names"].items() if v > 0]


NO_DECAY = ["bias", "LayerNorm.bias", "LayerNorm.weight", "norm.bias", "norm.weight", "norm1.bias", "norm1.weight",
            "norm2.bias", "norm2.weight"]
HEAD_NAMES = ["vqa_classifier", "nlvr2_classifier", "mlm_score", "itm_score", "snli_classifier"]
CROSS_MODAL_NAMES = ["cross_modal", "i2t", "t2i"]


def param_groups(pl_module):
    """The six AdamW groups of fiber_utils.set_schedule (:156-245), selected by name substrings."""
    cfg = pl_module.hparams.config
    lr, wd = cfg["learning_rate"], cfg["weight_decay"]
    groups = []
    for head, cross, mult in ((False, False, 1.0), (True, False, cfg["lr_mult_head"]),
                              (False, True, cfg["lr_mult_cross_modal"])):
        for decay in (True, False):
            ps = [p for n, p in pl_module.named_parameters()
                  if (not any(nd in n for nd in NO_DECAY)) == decay
                  and any(bb in n for bb in HEAD_NAMES) == head
                  and any(ht in n for ht in CROSS_MODAL_NAMES) == cross]
            groups.append({"params": ps, "weight_decay": wd if decay else 0.0, "lr": lr * mult})
    return groups


def resolve_max_steps(pl_module):
    """fiber_utils.py:254-262: trainer.max_steps, or — every fine-tuning config sets max_steps=None — the number of
    optimizer steps in max_epochs passes over the training dataloader.  Without a trainer (this repo's bench / tests
    drive training_step directly) the config's own max_steps is used."""
    cfg = pl_module.hparams.config
    trainer = getattr(pl_module, "trainer", None)
    max_steps = getattr(trainer, "max_steps", None) if trainer is not None else cfg.get("max_steps")
    if max_steps is None or max_steps < 0:  # PL >= 1.5 spells "unset" as -1
        if trainer is None or getattr(trainer, "datamodule", None) is None:
            raise ValueError("max_steps is None and there is no trainer.datamodule to derive it from "
                             "(len(train_dataloader) * max_epochs // accumulate_grad_batches)")
        max_steps = (len(trainer.datamodule.train_dataloader()) * trainer.max_epochs
                     // trainer.accumulate_grad_batches)
    return int(max_steps)


def set_schedule(pl_module):
    """fiber_utils.py:156-287.  Optimizer: on the GPU, fiber_b200.optim.FusedAdamW — transformers.AdamW (4.6.0)'s update
    in its exact operation order as one multi-tensor kernel.  For CPU-resident modules (host-side tests) torch.optim.AdamW
    stands in; the two differ in second-order details only (HF adds eps to sqrt(v) before the bias correction and
    decays after the Adam update, torch adds eps after the correction and decays first)."""
    import math
    cfg = pl_module.hparams.config
    groups = param_groups(pl_module)
    lr = cfg["learning_rate"]
    if cfg["optim_type"] == "adamw":
        on_gpu = all(p.is_cuda for g in groups for p in g["params"]) and any(len(g["params"]) for g in groups)
        if on_gpu:  # one fused launch per step with HF AdamW's exact update order (fiber_b200/optim.py, csrc/optim.cu)
            from ..optim import FusedAdamW
            optimizer = FusedAdamW(groups, lr=lr, eps=1e-8, betas=(0.9, 0.98))
        else:       # host-side construction (tests, group bookkeeping): same groups, torch's AdamW
            optimizer = torch.optim.AdamW(groups, lr=lr, eps=1e-8, betas=(0.9, 0.98))
    elif cfg["optim_type"] == "adam":
        optimizer = torch.optim.Adam(groups, lr=lr)
    elif cfg["optim_type"] == "sgd":
        optimizer = torch.optim.SGD(groups, lr=lr, momentum=0.9)
    else:
        raise ValueError("unknown optim_type %r" % cfg["optim_type"])
    max_steps = resolve_max_steps(pl_module)
    warmup = cfg["warmup_steps"]
    if isinstance(warmup, float):
        warmup = int(max_steps * warmup)
    end_lr, power = cfg["end_lr"], cfg["decay_power"]

    def cosine(step):  # transformers.get_cosine_schedule_with_warmup (num_cycles = 0.5)
        if step < warmup:
            return step / max(1, warmup)
        progress = (step - warmup) / max(1, max_steps - warmup)
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * progress)))

    def poly(step):  # transformers.get_polynomial_decay_schedule_with_warmup
        if step < warmup:
            return step / max(1, warmup)
        if step > max_steps:
            return end_lr / lr
        remaining = 1 - (step - warmup) / (max_steps - warmup)
        return ((lr - end_lr) * remaining ** power + end_lr) / lr

    scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, cosine if power == "cosine" else poly)
    return [optimizer], [{"scheduler": scheduler, "interval": "step"}]
