"""Mirror of the objectives a FIBER training step calls (coarse_grained/fiber/modules/objectives.py):
compute_mlm :17, compute_itm :44, compute_itm_hardneg :78, compute_itc :119, compute_vqa :182,
init_weights :502.  They are callers of infer(); losses are ordinary torch ops on the fp32 features
infer() returns.  One deliberate change: the 2*B `.item()` host syncs of the
hard-negative sampling loop (:154-165) are replaced by one batched torch.multinomial per direction
(same per-row distribution)."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F


def init_weights(module):
    if isinstance(module, (nn.Linear, nn.Embedding)):
        module.weight.data.normal_(mean=0.0, std=0.02)
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)
    if isinstance(module, nn.Linear) and module.bias is not None:
        module.bias.data.zero_()


# On the GPU the MLM loss comes from the fused decoder + cross-entropy (heads.MLMHead.loss_and_pred: the [B L, 50265] fp32
# logits never reach memory); the returned dict then carries "mlm_pred" (the arg-max the accuracy metric needs) instead
# of "mlm_logits".  FIBER_MLM_FUSED_CE=0 / set_fused_mlm_ce(False) restore the reference's logits -> F.cross_entropy form.
FUSED_MLM_CE = os.environ.get("FIBER_MLM_FUSED_CE", "1") != "0"


def set_fused_mlm_ce(on):
    global FUSED_MLM_CE
    FUSED_MLM_CE = bool(on)


def _mlm_tail(pl_module, text_feats, mlm_labels, mlm_ids):
    """objectives.py:19-41 of the reference on the text features of the masked pairs."""
    if FUSED_MLM_CE and text_feats.is_cuda and hasattr(pl_module.mlm_score, "loss_and_pred"):
        mlm_loss, mlm_pred = pl_module.mlm_score.loss_and_pred(text_feats, mlm_labels)
        ret = {"mlm_loss": mlm_loss, "mlm_pred": mlm_pred, "mlm_labels": mlm_labels, "mlm_ids": mlm_ids}
        acc_in = mlm_pred
    else:
        mlm_logits = pl_module.mlm_score(text_feats)
        mlm_loss = F.cross_entropy(mlm_logits.view(-1, pl_module.hparams.config["vocab_size"]).float(),
                                   mlm_labels.view(-1), ignore_index=-100)
        ret = {"mlm_loss": mlm_loss, "mlm_logits": mlm_logits, "mlm_labels": mlm_labels, "mlm_ids": mlm_ids}
        acc_in = mlm_logits
    phase = _phase(pl_module)
    loss = getattr(pl_module, f"{phase}_mlm_loss")(ret["mlm_loss"])
    acc = getattr(pl_module, f"{phase}_mlm_accuracy")(acc_in, ret["mlm_labels"])
    pl_module.log(f"mlm/{phase}/loss", loss)
    pl_module.log(f"mlm/{phase}/accuracy", acc)
    return ret


def _phase(pl_module):
    return "train" if pl_module.training else "val"


def compute_mlm(pl_module, batch):
    infer = pl_module.infer(batch, mask_text=True, mask_image=False)
    return _mlm_tail(pl_module, infer["text_feats"], infer["text_labels"], infer["text_ids"])


def _itm_tail(pl_module, infer, itm_labels):
    itm_logits = pl_module.itm_score(infer["cls_feats"])
    itm_loss = F.cross_entropy(itm_logits.float(), itm_labels.long())
    ret = {"itm_loss": itm_loss, "itm_logits": itm_logits, "itm_labels": itm_labels}
    phase = _phase(pl_module)
    loss = getattr(pl_module, f"{phase}_itm_loss")(ret["itm_loss"])
    acc = getattr(pl_module, f"{phase}_itm_accuracy")(ret["itm_logits"], ret["itm_labels"])
    pl_module.log(f"itm/{phase}/loss", loss)
    pl_module.log(f"itm/{phase}/accuracy", acc)
    return ret


def compute_itm(pl_module, batch, itm_labels=None):
    n = len(batch["text"])
    pos_len = n // 2
    if itm_labels is None:
        itm_labels = torch.cat([torch.ones(pos_len), torch.zeros(n - pos_len)])
        itm_labels = itm_labels[torch.randperm(n)]
    itm_labels = itm_labels.to(pl_module.device)
    pick = (itm_labels == 1).view(-1, 1, 1, 1)
    itm_images = [torch.where(pick, bti, bfi) for bti, bfi in zip(batch["image"], batch["false_image_0"])]
    batch = {k: v for k, v in batch.items()}
    batch["image"] = itm_images
    infer = pl_module.infer(batch, mask_text=False, mask_image=False)
    return _itm_tail(pl_module, infer, itm_labels)


def compute_itm_hardneg(pl_module, batch, image_neg, text_neg, text_mask_neg):
    pos_len = len(batch["text"])
    itm_labels = torch.cat([torch.ones(pos_len), torch.zeros(2 * pos_len)]).to(pl_module.device)
    batch = {k: v for k, v in batch.items()}
    batch["image"] = [torch.cat([batch["image"][0], batch["image"][0], image_neg], dim=0)]
    batch["text_masks"] = torch.cat([batch["text_masks"], text_mask_neg, batch["text_masks"]], dim=0)
    batch["text_ids"] = torch.cat([batch["text_ids"], text_neg, batch["text_ids"]], dim=0)
    batch["text_labels"] = torch.cat([batch["text_labels"]] * 3, dim=0)
    infer = pl_module.infer(batch, mask_text=False, mask_image=False)
    return _itm_tail(pl_module, infer, itm_labels)


def compute_mlm_itm_hardneg_merged(pl_module, batch, image_neg, text_neg, text_mask_neg):
    """compute_mlm + compute_itm_hardneg through ONE backbone pass of 4B samples.

    The reference runs infer() on the B masked-text pairs (objectives.py:18) and, separately, on the
    3B ITM pairs (:85-97).  Samples are independent through the whole backbone, so concatenating the
    two batches gives the same features; on the GPU it means a third fewer (and larger) kernel launches
    and one gradient accumulation less per parameter.  Returns the union of both functions' dicts."""
    B = len(batch["text"])
    img = batch["image"][0]
    merged = {k: v for k, v in batch.items()}
    merged["image"] = [torch.cat([img, img, img, image_neg], dim=0)]
    merged["text_ids"] = torch.cat([batch["text_ids_mlm"], batch["text_ids"], text_neg, batch["text_ids"]], dim=0)
    merged["text_masks"] = torch.cat([batch["text_masks"], batch["text_masks"], text_mask_neg, batch["text_masks"]], dim=0)
    merged["text_labels"] = torch.cat([batch["text_labels_mlm"]] + [batch["text_labels"]] * 3, dim=0)
    infer = pl_module.infer(merged, mask_text=False, mask_image=False)
    # ---- MLM on the first B samples (objectives.py:19-41) ----
    ret = _mlm_tail(pl_module, infer["text_feats"][:B], batch["text_labels_mlm"], batch["text_ids_mlm"])
    # ---- ITM on the remaining 3B samples (objectives.py:99-116) ----
    itm_labels = torch.cat([torch.ones(B), torch.zeros(2 * B)]).to(pl_module.device)
    ret.update(_itm_tail(pl_module, {"cls_feats": infer["cls_feats"][B:]}, itm_labels))
    return ret


def _rows_of_cat(first, second, idx):
    """torch.cat([first, second], 0)[idx] without the concatenation (no host sync: one gather per source + where)."""
    n = first.shape[0]
    if second.shape[0] == 0:
        return first[idx]
    pick_first = (idx < n).view((-1,) + (1,) * (first.dim() - 1))
    return torch.where(pick_first, first[idx.clamp(max=n - 1)], second[(idx - n).clamp(min=0)].to(first.dtype))


def compute_itc(pl_module, batch):
    if hasattr(pl_module, "queue_sync"):
        pl_module.queue_sync()  # the previous step's queue update may still be running on its side stream
    with torch.no_grad():
        pl_module.temp.clamp_(0.001, 1.0)
    infer_image = pl_module.infer(batch, mask_image=False, mask_text=False, image_only=True)
    infer_text = pl_module.infer(batch, mask_image=False, mask_text=False, text_only=True)
    image_feat, text_feat = infer_image["cls_feats"], infer_text["cls_feats"]
    image_feat_all = torch.cat([image_feat.t().detach(), pl_module.image_queue.clone().detach()], dim=1)
    text_feat_all = torch.cat([text_feat.t().detach(), pl_module.text_queue.clone().detach()], dim=1)
    sim_i2t = image_feat @ text_feat_all / pl_module.temp
    sim_t2i = text_feat @ image_feat_all / pl_module.temp
    sim_targets = torch.zeros_like(sim_i2t)
    sim_targets.fill_diagonal_(1)
    loss_i2t = -torch.sum(F.log_softmax(sim_i2t, dim=1) * sim_targets, dim=1).mean()
    loss_t2i = -torch.sum(F.log_softmax(sim_t2i, dim=1) * sim_targets, dim=1).mean()
    loss_itc = (loss_i2t + loss_t2i) / 2.0
    bs = image_feat.size(0)
    qt = pl_module.queue_counters()[1]  # host mirror of queue_total (no device sync in steady state)
    with torch.no_grad():
        weights_i2t = F.softmax(sim_i2t[:, :bs + qt], dim=1)
        weights_t2i = F.softmax(sim_t2i[:, :bs + qt], dim=1)
        weights_i2t.fill_diagonal_(0)
        weights_t2i.fill_diagonal_(0)
        img_idx = torch.multinomial(weights_t2i + 1e-9, 1).view(-1)
        txt_idx = torch.multinomial(weights_i2t + 1e-9, 1).view(-1)
        # rows img_idx / txt_idx of cat([batch, queue[:qt]]) (objectives.py:142-166) without materialising the
        # concatenation: with a full queue the reference copies 4096 raw images (7.2 GB at 384 px) per step
        image_neg = _rows_of_cat(batch["image"][0], pl_module.image_input_queue[:qt], img_idx)
        text_neg = _rows_of_cat(batch["text_ids"], pl_module.text_input_queue[:qt], txt_idx)
        text_mask_neg = _rows_of_cat(batch["text_masks"], pl_module.text_input_mask_queue[:qt], txt_idx)
    if pl_module.training:
        pl_module._dequeue_and_enqueue(image_feat.detach().clone(), text_feat.detach().clone(),
                                       batch["image"][0].clone(), batch["text_ids"].clone(),
                                       batch["text_masks"].clone())
    ret = {"itc_loss": loss_itc}
    phase = _phase(pl_module)
    loss = getattr(pl_module, f"{phase}_itc_loss")(ret["itc_loss"])
    pl_module.log(f"itc/{phase}/loss", loss)
    return ret, image_neg, text_neg, text_mask_neg


def compute_vqa(pl_module, batch):
    infer = pl_module.infer(batch, mask_text=False, mask_image=False)
    vqa_logits = pl_module.vqa_classifier(infer["cls_feats"])
    vqa_targets = torch.zeros(len(vqa_logits), pl_module.hparams.config["vqav2_label_size"])
    for i, (_label, _score) in enumerate(zip(batch["vqa_labels"], batch["vqa_scores"])):
        for l, s in zip(_label, _score):
            vqa_targets[i, l] = s
    vqa_targets = vqa_targets.to(pl_module.device)
    vqa_loss = F.binary_cross_entropy_with_logits(vqa_logits.float(), vqa_targets) * vqa_targets.shape[1]
    ret = {"vqa_loss": vqa_loss, "vqa_logits": vqa_logits, "vqa_targets": vqa_targets,
           "vqa_labels": batch["vqa_labels"], "vqa_scores": batch["vqa_scores"]}
    phase = _phase(pl_module)
    loss = getattr(pl_module, f"{phase}_vqa_loss")(ret["vqa_loss"])
    score = getattr(pl_module, f"{phase}_vqa_score")(ret["vqa_logits"], ret["vqa_targets"])
    pl_module.log(f"vqa/{phase}/loss", loss)
    pl_module.log(f"vqa/{phase}/score", score)
    return ret


def vqa_test_step(pl_module, batch, output):
    """objectives.py:652-668: argmax answers of a test batch (the id -> answer table lives in the datamodule)."""
    try:
        id2answer = pl_module.trainer.datamodule.dm_dicts["vqa_trainval"].id2answer if "vqa_trainval" in \
            pl_module.trainer.datamodule.dm_dicts else pl_module.trainer.datamodule.dm_dicts["vqa"].id2answer
    except Exception:  # noqa: BLE001  (no datamodule attached: return the label ids)
        id2answer = None
    vqa_preds = output["vqa_logits"].argmax(dim=-1)
    vqa_preds = [id2answer[p.item()] if id2answer is not None else p.item() for p in vqa_preds]
    return {"qids": batch.get("qid"), "preds": vqa_preds}
