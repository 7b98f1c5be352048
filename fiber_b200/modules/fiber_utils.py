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


def set_schedule(pl_module):
    cfg = pl_module.hparams.config
    groups = param_groups(pl_module)
    lr = cfg["learning_rate"]
    if cfg["optim_type"] == "adamw":
        optimizer = torch.optim.AdamW(groups, lr=lr, eps=1e-8, betas=(0.9, 0.98))
    elif cfg["optim_type"] == "adam":
        optimizer = torch.optim.Adam(groups, lr=lr)
    else:
        optimizer = torch.optim.SGD(groups, lr=lr, momentum=0.9)
    max_steps = cfg["max_steps"]
    warmup = cfg["warmup_steps"]
    if isinstance(warmup, float):
        warmup = int(max_steps * warmup)
    end_lr, power = cfg["end_lr"], cfg["decay_power"]

    def poly(step):  # transformers.get_polynomial_decay_schedule_with_warmup
        if step < warmup:
            return step / max(1, warmup)
        if step > max_steps:
            return end_lr / lr
        remaining = 1 - (step - warmup) / (max_steps - warmup)
        return ((lr - end_lr) * remaining ** power + end_lr) / lr

    scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, poly)
    return [optimizer], [{"scheduler": scheduler, "interval": "step"}]
