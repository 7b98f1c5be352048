"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only (`python tools/make_golden.py`): the reference is imported through
baseline/ref_shims.py, its parameters are overwritten with oracle/synth.py's name-seeded values, and
small slices / statistics of its outputs and gradients are saved.  tests/test_oracle_golden.py
then pins oracle/fiber_oracle.py against these files on CPU, anywhere.
"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from baseline import ref_shims  # noqa: E402
from oracle import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
ref_swin, ref_roberta = ref_shims.install()
from fiber.modules import FIBERTransformerSS  # noqa: E402


def fill(module, prefix=""):
    """Overwrite every floating parameter/buffer of `module` with the synth recipe (by name)."""
    sd = module.state_dict()
    shapes = {prefix + k: (tuple(v.shape), v.dtype) for k, v in sd.items() if not k.startswith("rank_output")}
    new = synth.synth_state_dict(shapes)
    module.load_state_dict({k[len(prefix):]: v for k, v in new.items()}, strict=False)
    return new


def no_dropout(module):
    for m in module.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if m.__class__.__name__ == "DropPath":
            m.drop_prob = 0.0


def grad_stats(module, prefix=""):
    out = {}
    for n, p in module.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.double()
        out[prefix + n] = (float(g.sum()), float(g.norm()), p.grad.flatten()[:8].clone())
    return out


def probe(name, shape):
    return synth.synth_tensor("probe." + name, shape) * 1.0  # fixed random projection for scalar losses


def gold_blocks():
    """Block-level fixtures at small widths (full outputs stored)."""
    out = {}
    B, H, ws, C, nh, L, Ct = 2, 14, 7, 64, 2, 6, 48
    x = synth.synth_tensor("in.x", (B, H * H, C))
    y = synth.synth_tensor("in.y", (B, L, Ct))
    ymask = torch.zeros(B, 1, 1, L)
    ymask[1, :, :, 4:] = -10000.0
    for tag, shift, fused in (("plain", 0, False), ("shift", 3, False), ("fused", 0, True), ("fused_shift", 3, True)):
        blk = ref_swin.SwinTransformerBlock(C, (H, H), nh, window_size=ws, shift_size=shift,
                                            dim_text=Ct if fused else None)
        pre = "vit_model.layers.0.blocks.%s" % tag
        sd = fill(blk, pre + ".")
        xi = x.clone().requires_grad_(True)
        yi = y.clone().requires_grad_(True)
        o = blk(xi, yi, ymask) if fused else blk(xi)
        (o * probe("swin", o.shape)).sum().backward()
        out["swin_" + tag] = {"out": o.detach(), "dx": xi.grad, "dy": yi.grad if fused else None,
                              "grads": grad_stats(blk, pre + "."), "shift": shift, "fused": fused}
    # stage-3 style block: one window, resolution == window (shift forced to 0, :304-307)
    blk = ref_swin.SwinTransformerBlock(C, (ws, ws), nh, window_size=ws, shift_size=3, dim_text=Ct)
    pre = "vit_model.layers.0.blocks.onewin"
    fill(blk, pre + ".")
    x1 = x[:, : ws * ws].clone().requires_grad_(True)
    o = blk(x1, y, ymask)
    (o * probe("swin1", o.shape)).sum().backward()
    out["swin_onewin"] = {"out": o.detach(), "dx": x1.grad, "grads": grad_stats(blk, pre + ".")}

    pm = ref_swin.PatchMerging((H, H), C)
    fill(pm, "vit_model.layers.0.downsample.")
    xi = x.clone().requires_grad_(True)
    o = pm(xi)
    (o * probe("pm", o.shape)).sum().backward()
    out["patch_merging"] = {"out": o.detach(), "dx": xi.grad, "grads": grad_stats(pm, "vit_model.layers.0.downsample.")}

    from timm.models.layers import PatchEmbed
    pe = PatchEmbed(img_size=56, patch_size=4, in_chans=3, embed_dim=C, norm_layer=nn.LayerNorm)
    fill(pe, "vit_model.patch_embed.")
    img = synth.synth_tensor("in.img", (B, 3, 56, 56))
    o = pe(img)
    (o * probe("pe", o.shape)).sum().backward()
    out["patch_embed"] = {"out": o.detach(), "grads": grad_stats(pe, "vit_model.patch_embed.")}

    # RoBERTa layers at width 64
    from transformers import RobertaConfig
    cfg = RobertaConfig(vocab_size=100, hidden_size=64, num_hidden_layers=12, num_attention_heads=4,
                        intermediate_size=128, max_position_embeddings=30, type_vocab_size=1, layer_norm_eps=1e-5,
                        pad_token_id=1, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    cfg.position_embedding_type = "absolute"
    cfg.chunk_size_feed_forward = 0
    cfg.is_decoder = False
    cfg.add_cross_attention = False
    ref_roberta.DIM_IMG = 64
    ref_roberta.NUM_FUSE_BLOCK = 6
    Lt = 10
    h = synth.synth_tensor("in.h", (B, Lt, 64))
    tm = torch.ones(B, Lt, dtype=torch.long)
    tm[1, 7:] = 0
    em = (1.0 - tm[:, None, None, :].float()) * -10000.0
    img32 = synth.synth_tensor("in.img32", (B, 20, 32))
    img64 = synth.synth_tensor("in.img64", (B, 9, 64))
    for tag, li, image, last_norm in (("plain", 2, None, True), ("fused", 7, img32, True), ("fused_nonorm", 11, img64, False)):
        layer = ref_roberta.RobertaLayer(cfg, layer_index=li)
        pre = "text_transformer.encoder.layer.%d." % li
        fill(layer, pre)
        hi = h.clone().requires_grad_(True)
        ii = image.clone().requires_grad_(True) if image is not None else None
        o = layer(hi, em, encoder_hidden_states=ii, last_norm=last_norm)[0]
        (o * probe("rl", o.shape)).sum().backward()
        out["roberta_" + tag] = {"out": o.detach(), "dh": hi.grad, "dimg": ii.grad if ii is not None else None,
                                 "grads": grad_stats(layer, pre), "layer": li, "last_norm": last_norm}
    emb = ref_roberta.RobertaEmbeddings(cfg)
    fill(emb, "text_transformer.embeddings.")
    ids = torch.tensor([[0, 5, 6, 7, 2, 1, 1, 1, 1, 1], [0, 9, 8, 7, 6, 5, 4, 3, 11, 2]])
    o = emb(input_ids=ids)
    (o * probe("emb", o.shape)).sum().backward()
    out["roberta_embeddings"] = {"out": o.detach(), "ids": ids, "grads": grad_stats(emb, "text_transformer.embeddings.")}
    torch.save(out, os.path.join(GOLD, "blocks.pt"))
    print("blocks.pt:", {k: tuple(v["out"].shape) for k, v in out.items()})


def model_shapes(model):
    return {k: (tuple(v.shape), v.dtype) for k, v in model.state_dict().items()}


def gold_model(tag, image_size, tasks, B, L, train_step=True, infer_modes=("fused",)):
    torch.manual_seed(0)
    cfg = ref_shims.default_config(tasks=tasks, image_size=image_size, max_text_len=L,
                                   draw_false_image=1 if ("itm" in tasks and "itc" not in tasks) else 0)
    model = FIBERTransformerSS(cfg)
    fill(model)
    batch = synth.synth_batch(B, image_size, L, seed=1234, false_image=True, vqa="vqa" in tasks)
    out = {"cfg": cfg, "B": B, "L": L, "state_keys": {k: tuple(s) for k, (s, _) in model_shapes(model).items()}}
    model.eval()
    with torch.no_grad():
        for mode in infer_modes:
            r = model.infer(batch, image_only=(mode == "image_only"), text_only=(mode == "text_only"))
            out["infer_" + mode] = {
                "cls_feats": r["cls_feats"].clone(),
                "text_feats": None if r["text_feats"] is None else r["text_feats"][:, :3, :64].clone(),
                "image_feats": None if r["image_feats"] is None else r["image_feats"][:, :5, :64].clone(),
            }
    if train_step:
        model.train()
        no_dropout(model)
        model.zero_grad()
        torch.manual_seed(77)
        itm_labels = None
        if "itm" in tasks and "itc" not in tasks:
            n = len(batch["text"])
            lab = torch.cat([torch.ones(n // 2), torch.zeros(n - n // 2)])
            st = torch.random.get_rng_state()
            itm_labels = lab[torch.randperm(n)]  # what compute_itm will draw first (objectives.py:47-48)
            torch.random.set_rng_state(st)
        loss = model.training_step({k: v for k, v in batch.items()}, 0)
        loss.backward()
        out["train"] = {"loss": float(loss), "itm_labels": itm_labels, "grads": grad_stats(model),
                        "logged": {k: float(v) for k, v in model.logged.items()}}
    torch.save(out, os.path.join(GOLD, "model_%s.pt" % tag))
    print("model_%s.pt" % tag, "loss", out.get("train", {}).get("loss"), "keys", len(out["state_keys"]))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["blocks", "cfg0", "fused384", "vqa"]
    if "blocks" in which:
        gold_blocks()
    if "cfg0" in which:   # BASELINE.json configs[0]: ITM+MLM, 224 px / 40 tok, B=2
        gold_model("cfg0_224_itm_mlm", 224, ["itm", "mlm"], 2, 40, infer_modes=("fused",))
    if "fused384" in which:  # the three infer() variants at the north-star resolution
        gold_model("384_infer", 384, ["itm", "mlm", "itc"], 2, 40, train_step=False,
                   infer_modes=("fused", "image_only", "text_only"))
    if "vqa" in which:    # VQA head path at 224 px / 50 tok (576 px is covered by the window-324 block test)
        gold_model("224_vqa", 224, ["vqa"], 2, 50, infer_modes=())


def gold_itc():
    """objectives.compute_itc of the unmodified reference (BASELINE configs[1] / [3] objective) at 224 px, B = 3:
    the contrastive loss against the (synthetic) queues, its gradients, and the queue state after
    _dequeue_and_enqueue.  The hard-negative draws are RNG-dependent and not part of the fixture."""
    import torch.distributed as dist
    from fiber.modules import objectives as ref_obj
    if not dist.is_initialized():  # concat_all_gather (fiber_module.py:12-24) needs a process group: 1-rank gloo
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29641")
        dist.init_process_group("gloo", rank=0, world_size=1)
    torch.manual_seed(0)
    B, L, size = 3, 40, 224
    cfg = ref_shims.default_config(tasks=["itm", "mlm", "itc"], image_size=size, max_text_len=L)
    model = FIBERTransformerSS(cfg)
    fill(model)
    batch = synth.synth_batch(B, size, L, seed=1234, false_image=True)
    model.train()
    no_dropout(model)
    model.zero_grad()
    torch.manual_seed(5)
    ret, image_neg, text_neg, text_mask_neg = ref_obj.compute_itc(model, {k: v for k, v in batch.items()})
    ret["itc_loss"].backward()
    out = {"cfg": cfg, "B": B, "L": L, "state_keys": {k: tuple(s) for k, (s, _) in model_shapes(model).items()},
           "itc_loss": float(ret["itc_loss"]), "grads": grad_stats(model),
           "queue_ptr": int(model.queue_ptr), "queue_total": int(model.queue_total),
           "image_queue_head": model.image_queue[:, :B].clone(), "text_queue_head": model.text_queue[:, :B].clone(),
           "text_input_queue_head": model.text_input_queue[:B].clone(),
           "neg_shapes": (tuple(image_neg.shape), tuple(text_neg.shape), tuple(text_mask_neg.shape))}
    torch.save(out, os.path.join(GOLD, "model_224_itc.pt"))
    print("model_224_itc.pt", "itc_loss", out["itc_loss"], "queue", out["queue_ptr"], out["queue_total"],
          "grads", len(out["grads"]))


def gold_itm_hardneg():
    """objectives.compute_itm_hardneg (:78-116), the ITM objective of BASELINE configs[1] (3B-sample pass: positives,
    (image, negative text), (negative image, text)) with deterministic negatives (the batch rolled by one)."""
    from fiber.modules import objectives as ref_obj
    torch.manual_seed(0)
    B, L, size = 2, 40, 224
    cfg = ref_shims.default_config(tasks=["itm", "mlm", "itc"], image_size=size, max_text_len=L)
    model = FIBERTransformerSS(cfg)
    fill(model)
    batch = synth.synth_batch(B, size, L, seed=1234, false_image=True)
    image_neg = batch["image"][0].roll(1, 0)
    text_neg, text_mask_neg = batch["text_ids"].roll(1, 0), batch["text_masks"].roll(1, 0)
    model.train()
    no_dropout(model)
    model.zero_grad()
    ret = ref_obj.compute_itm_hardneg(model, {k: v for k, v in batch.items()}, image_neg, text_neg, text_mask_neg)
    ret["itm_loss"].backward()
    out = {"cfg": cfg, "B": B, "L": L, "state_keys": {k: tuple(s) for k, (s, _) in model_shapes(model).items()},
           "itm_loss": float(ret["itm_loss"].detach()), "itm_logits": ret["itm_logits"].detach().clone(),
           "grads": grad_stats(model)}
    torch.save(out, os.path.join(GOLD, "model_224_itm_hardneg.pt"))
    print("model_224_itm_hardneg.pt", "itm_loss", out["itm_loss"], "grads", len(out["grads"]))


if __name__ == "__main__" and "itc" in sys.argv[1:]:
    gold_itc()
if __name__ == "__main__" and "hardneg" in sys.argv[1:]:
    gold_itm_hardneg()


def gold_schedule():
    """Parameter-group sizes of the reference's fiber_utils.set_schedule (name-substring rules)."""
    import types
    from fiber.modules import fiber_utils as ref_utils
    out = {}
    for tag, tasks, size in (("cfg0", ["itm", "mlm"], 224), ("cfg1", ["itm", "mlm", "itc"], 384), ("vqa", ["vqa"], 224)):
        cfg = ref_shims.default_config(tasks=tasks, image_size=size)
        model = FIBERTransformerSS(cfg)
        model.trainer = types.SimpleNamespace(max_steps=1000)
        opt = ref_utils.set_schedule(model)[0][0]
        names = {id(p): n for n, p in model.named_parameters()}
        out[tag] = [{"lr": g.get("initial_lr", g["lr"]), "weight_decay": g["weight_decay"], "n": len(g["params"]),
                     "numel": sum(p.numel() for p in g["params"]),
                     "first": sorted(names[id(p)] for p in g["params"])[:3]} for g in opt.param_groups]
    torch.save(out, os.path.join(GOLD, "schedule.pt"))
    print("schedule.pt", {k: [g["n"] for g in v] for k, v in out.items()})


if __name__ == "__main__" and "schedule" in sys.argv[1:]:
    gold_schedule()


def gold_train384():
    """BASELINE configs[1] objective mix at the north-star resolution (384 px / 40 tok, B = 2), dropout off, through the
    unmodified reference: compute_itc (queues empty) + compute_itm_hardneg with deterministic negatives (the batch
    rolled by one) + compute_mlm, summed and back-propagated.  Stores the three losses, the ITM logits, the MLM logits
    at the labelled positions and their log-sum-exp, and every parameter-gradient statistic."""
    import torch.distributed as dist
    from fiber.modules import objectives as ref_obj
    if not dist.is_initialized():
        dist.init_process_group("gloo", store=dist.HashStore(), rank=0, world_size=1)
    torch.manual_seed(0)
    B, L, size = 2, 40, 384
    cfg = ref_shims.default_config(tasks=["itm", "mlm", "itc"], image_size=size, max_text_len=L)
    model = FIBERTransformerSS(cfg)
    fill(model)
    batch = synth.synth_batch(B, size, L, seed=1234, false_image=True)
    image_neg = batch["image"][0].roll(1, 0)
    text_neg, text_mask_neg = batch["text_ids"].roll(1, 0), batch["text_masks"].roll(1, 0)
    model.train()
    no_dropout(model)
    model.zero_grad()
    torch.manual_seed(5)
    itc, _, _, _ = ref_obj.compute_itc(model, dict(batch))
    itm = ref_obj.compute_itm_hardneg(model, dict(batch), image_neg, text_neg, text_mask_neg)
    mlm = ref_obj.compute_mlm(model, dict(batch))
    loss = itc["itc_loss"] + itm["itm_loss"] + mlm["mlm_loss"]
    loss.backward()
    lab = batch["text_labels_mlm"]
    pos = (lab != -100).nonzero()
    ml = mlm["mlm_logits"].detach()
    out = {"cfg": cfg, "B": B, "L": L, "state_keys": {k: tuple(s) for k, (s, _) in model_shapes(model).items()},
           "loss": float(loss), "itc_loss": float(itc["itc_loss"]),
           "itm_loss": float(itm["itm_loss"]), "mlm_loss": float(mlm["mlm_loss"]),
           "itm_logits": itm["itm_logits"].detach().clone(),
           "mlm_pos": pos, "mlm_logit_at_label": ml[pos[:, 0], pos[:, 1], lab[pos[:, 0], pos[:, 1]]].clone(),
           "mlm_lse": torch.logsumexp(ml[pos[:, 0], pos[:, 1]], dim=-1), "mlm_logits_head": ml[:, :, :128].clone(),
           "grads": grad_stats(model)}
    torch.save(out, os.path.join(GOLD, "model_384_train.pt"))
    print("model_384_train.pt loss", out["loss"], out["itc_loss"], out["itm_loss"], out["mlm_loss"], "grads", len(out["grads"]))


def gold_vqa576():
    """BASELINE configs[2]: VQA at 576 px / 50 tokens (18 x 18 = 324-token windows), B = 2: infer() features, the VQA
    logits and loss, gradient statistics."""
    from fiber.modules import objectives as ref_obj
    torch.manual_seed(0)
    B, L, size = 2, 50, 576
    cfg = ref_shims.default_config(tasks=["vqa"], image_size=size, max_text_len=L)
    model = FIBERTransformerSS(cfg)
    fill(model)
    batch = synth.synth_batch(B, size, L, seed=1234, vqa=True)
    model.train()
    no_dropout(model)
    model.zero_grad()
    ret = ref_obj.compute_vqa(model, dict(batch))
    ret["vqa_loss"].backward()
    model.eval()
    with torch.no_grad():
        r = model.infer(dict(batch))
    out = {"cfg": cfg, "B": B, "L": L, "state_keys": {k: tuple(s) for k, (s, _) in model_shapes(model).items()},
           "vqa_loss": float(ret["vqa_loss"]), "vqa_logits": ret["vqa_logits"].detach().clone(),
           "cls_feats": r["cls_feats"].clone(), "text_feats": r["text_feats"][:, :3, :64].clone(),
           "image_feats": r["image_feats"][:, :5, :64].clone(), "grads": grad_stats(model)}
    torch.save(out, os.path.join(GOLD, "model_576_vqa.pt"))
    print("model_576_vqa.pt vqa_loss", out["vqa_loss"], "grads", len(out["grads"]))


if __name__ == "__main__" and "train384" in sys.argv[1:]:
    gold_train384()
if __name__ == "__main__" and "vqa576" in sys.argv[1:]:
    gold_vqa576()
