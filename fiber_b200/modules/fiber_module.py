"""FIBERTransformerSS — host-side mirror of coarse_grained/fiber/modules/fiber_module.py:26-523.

Same constructor argument (the sacred `_config` dict), sub-module names, state_dict keys and
`infer` / `forward` / `training_step` contracts as the reference, so coarse_grained/run.py and the
ITM / ITC / MLM / VQA objectives sit on top unchanged; the fused backbone underneath runs on the
sm_100a kernels of this package (no eager or CPU fallback: calling it without the CUDA library or
with CPU tensors raises)."""
import os

import torch
import torch.nn as nn

from . import fiber_utils, heads, objectives, roberta, swin_transformer
from .lightning import LightningModule
from .roberta import RobertaModel
from .swin_transformer import FLinear


@torch.no_grad()
def concat_all_gather(tensor):
    """fiber_module.py:12-24; a single process is its own world."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return tensor
    # same result as the reference's list all_gather + cat (rank-major concatenation along dim 0), without the
    # world_size ones_like fills and the extra cat pass over the gathered data (906 MB of raw images at 8 ranks)
    world = torch.distributed.get_world_size()
    tensor = tensor.contiguous()
    out = torch.empty((world * tensor.shape[0],) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    torch.distributed.all_gather_into_tensor(out, tensor)
    return out


def _flinear_fp32(i, o):
    m = FLinear(i, o)
    m.out_fp32 = True
    return m


class FIBERTransformerSS(LightningModule):
    def __init__(self, config):
        super().__init__()
        self.save_hyperparameters(config=config) if not hasattr(self.hparams, "config") else None
        self.hparams["config"] = config
        self.config = config
        hs = config["hidden_size"]
        self.num_fuse_block = config["num_fuse_block"]
        self.num_text_layer = config["num_layers"]
        roberta.NUM_FUSE_BLOCK = swin_transformer.NUM_FUSE_BLOCK = self.num_fuse_block
        roberta.DIM_IMG = config["input_image_embed_size"]

        self.cross_modal_text_transform = _flinear_fp32(config["input_text_embed_size"], hs)
        self.cross_modal_image_transform = _flinear_fp32(config["input_image_embed_size"], hs)
        self.cross_modal_text_transform_itc = _flinear_fp32(config["input_text_embed_size"], hs)
        self.cross_modal_image_transform_itc = _flinear_fp32(config["input_image_embed_size"], hs)
        for m in (self.cross_modal_text_transform, self.cross_modal_image_transform,
                  self.cross_modal_text_transform_itc, self.cross_modal_image_transform_itc):
            m.apply(objectives.init_weights)

        if config["loss_names"]["itc"] > 0:  # ALBEF queues, fiber_module.py:61-70
            self.temp = nn.Parameter(torch.ones([]) * 0.07)
            self.queue_size = 4096
            self.register_buffer("image_queue", torch.randn(hs, self.queue_size))
            self.register_buffer("text_queue", torch.randn(hs, self.queue_size))
            self.register_buffer("image_input_queue",
                                 torch.randn(self.queue_size, 3, config["image_size"], config["image_size"]))
            self.register_buffer("text_input_queue", torch.zeros(self.queue_size, config["max_text_len"], dtype=torch.long))
            self.register_buffer("text_input_mask_queue",
                                 torch.zeros(self.queue_size, config["max_text_len"], dtype=torch.long))
            self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))
            self.register_buffer("queue_total", torch.zeros(1, dtype=torch.long))

        self.vit_model = getattr(swin_transformer, config["vit"])(pretrained=config["pretrained_vit"], config=config)
        self.avgpool = nn.AdaptiveAvgPool1d(1)
        self.text_transformer = RobertaModel.from_pretrained(config["tokenizer"])

        self.cross_modal_image_pooler = heads.Pooler(hs)
        self.cross_modal_text_pooler = heads.Pooler(hs)
        self.cross_modal_image_pooler.apply(objectives.init_weights)
        self.cross_modal_text_pooler.apply(objectives.init_weights)
        self.itc_pooler = config["itc_pooler"]
        if self.itc_pooler:
            self.cross_modal_image_pooler_itc = heads.Pooler(hs)
            self.cross_modal_text_pooler_itc = heads.Pooler(hs)
            self.cross_modal_image_pooler_itc.apply(objectives.init_weights)
            self.cross_modal_text_pooler_itc.apply(objectives.init_weights)

        ln = config["loss_names"]
        for k in ("caption_mle", "caption_gold", "caption_cider", "nlvr2"):
            if ln.get(k, 0) > 0:
                raise NotImplementedError("%s is outside the B200 hot-path scope (SURVEY.md §8)" % k)
        if ln["mlm"] > 0:
            self.mlm_score = heads.MLMHead(hs, config["vocab_size"])
            self.mlm_score.apply(objectives.init_weights)
        if ln["itm"] > 0:
            self.itm_score = heads.ITMHead(hs * 2)
            self.itm_score.apply(objectives.init_weights)
            self.rank_output = nn.Linear(hs, 1)  # aliases itm_score.fc row 1 (fiber_module.py:112-114)
            self.rank_output.weight.data = self.itm_score.fc.weight.data[1:, :]
            self.rank_output.bias.data = self.itm_score.fc.bias.data[1:]

        if config["load_path"] != "" and not config["test_only"]:
            self._load(config["load_path"], adapt=True)
        if ln["vqa"] > 0:
            vs = config["vqav2_label_size"]
            self.vqa_classifier = nn.Sequential(nn.Linear(hs * 2, hs * 2), nn.LayerNorm(hs * 2), nn.GELU(),
                                                nn.Linear(hs * 2, vs))
            self.vqa_classifier.apply(objectives.init_weights)
        fiber_utils.set_metrics(self)
        self.current_tasks = list()
        self.merge_mlm_itm_pass = os.environ.get("FIBER_MERGE_PASSES", "1") != "0"
        if config["load_path"] != "" and config["test_only"]:
            self._load(config["load_path"])

    def _load(self, path, adapt=False):
        """fiber_module.py:139-148 (fine-tuning start: relative-position tables resized from resolution_before to
        image_size, e.g. the 384-px pre-training checkpoint at 576 px for VQA) and :174-180 (test_only: as is)."""
        state_dict = torch.load(path, map_location="cpu")["state_dict"]
        for key in ["image_queue", "text_queue", "queue_ptr", "queue_total", "image_input_queue", "text_input_queue",
                    "text_input_mask_queue"]:
            state_dict.pop(key, None)
        if adapt:
            state_dict = swin_transformer.swin_adapt_position_encoding(
                state_dict, before=self.config.get("resolution_before", self.config["image_size"]),
                after=self.config["image_size"])
        self.load_state_dict(state_dict, strict=False)

    @torch.no_grad()
    def _dequeue_and_enqueue(self, image_feat, text_feat, image_input, text_input, text_input_mask):
        """fiber_module.py:181-222 (ring buffer with wrap-around).

        Multi-GPU: the reference runs its five all_gathers synchronously in the middle of the step — at 8 ranks the raw
        images alone are 0.9 GB gathered per step.  Nothing reads the queues again before the NEXT step's compute_itc, so
        with FIBER_ITC_ASYNC_QUEUE=1 the gathers and the queue writes run on a side stream while the compute stream goes
        on with the hard-negative ITM pass and the backward; `queue_sync()` — called by forward(), compute_itc,
        queue_counters() and state_dict() — makes the compute stream wait for the update before any read.  Opt-in: at
        N = 8 it measured 2386 pairs/s against 2388 for the synchronous update (profiles/r2_scaling_ab.txt) — the
        gathers take ~1.5 ms on NVSwitch; the scaling loss is the spread between the GPUs (bench.py gpu_speed_probe).
        Every rank issues the collectives at the same point of its step, so their order on the communicator is the same
        everywhere.  The queue contents are bit-identical to the synchronous update."""
        world = torch.distributed.get_world_size() if (torch.distributed.is_available()
                                                       and torch.distributed.is_initialized()) else 1
        overlap = (world > 1 and image_feat.is_cuda and os.environ.get("FIBER_ITC_ASYNC_QUEUE", "0") == "1")
        self.queue_sync()  # a previous update still in flight writes the same buffers
        ptr, total = self.queue_counters()
        n = image_feat.shape[0] * world
        if overlap:
            cur = torch.cuda.current_stream()
            side = self.__dict__.get("_queue_stream")
            if side is None:
                side = self.__dict__["_queue_stream"] = torch.cuda.Stream()
            side.wait_stream(cur)
            for t in (image_feat, text_feat, image_input, text_input, text_input_mask):
                t.record_stream(side)
            ctx = torch.cuda.stream(side)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            image_feats = concat_all_gather(image_feat)
            text_feats = concat_all_gather(text_feat)
            image_input = concat_all_gather(image_input)
            text_input = concat_all_gather(text_input)
            text_input_mask = concat_all_gather(text_input_mask)
            idx = (ptr + torch.arange(n, device=image_feats.device)) % self.queue_size
            self.image_queue[:, idx] = image_feats.T.float()
            self.text_queue[:, idx] = text_feats.T.float()
            self.image_input_queue[idx] = image_input
            self.text_input_queue[idx] = text_input
            self.text_input_mask_queue[idx] = text_input_mask
            self.queue_ptr[0] = (ptr + n) % self.queue_size
            self.queue_total[0] = total + n
            if overlap:
                self.__dict__["_queue_event"] = side.record_event()
        self._queue_host = ((ptr + n) % self.queue_size, total + n, self.queue_ptr._version, self.queue_total._version)

    def queue_sync(self):
        """Order the current stream after an ITC queue update that is still running on the side stream."""
        ev = self.__dict__.pop("_queue_event", None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def state_dict(self, *a, **k):
        self.queue_sync()
        return super().state_dict(*a, **k)

    def queue_counters(self):
        """(queue_ptr, queue_total) as Python ints.  The reference reads both device buffers with int(...) every
        step (fiber_module.py:205, objectives.py:137): three host syncs.  Here the values written by the last
        _dequeue_and_enqueue are mirrored on the host and re-read from the device only when somebody else
        modified the buffers (load_state_dict, manual reset), detected through their version counters."""
        h = getattr(self, "_queue_host", None)
        if h is None:
            self.queue_sync()
        if h is None or h[2] != self.queue_ptr._version or h[3] != self.queue_total._version:
            h = (int(self.queue_ptr), int(self.queue_total), self.queue_ptr._version, self.queue_total._version)
            self._queue_host = h
        return h[0], h[1]

    def infer(self, batch, mask_text=False, mask_image=False, image_token_type_idx=1, img=None, text_only=False,
              image_only=False):
        if not text_only and img is None:
            imgkey = f"image_{image_token_type_idx - 1}" if f"image_{image_token_type_idx - 1}" in batch else "image"
            img = batch[imgkey][0]
        text_ids = text_labels = text_masks = None
        if not image_only:
            do_mlm = "_mlm" if mask_text else ""
            text_ids = batch[f"text_ids{do_mlm}"]
            text_labels = batch[f"text_labels{do_mlm}"]
            text_masks = batch["text_masks"]
        tt, vit = self.text_transformer, self.vit_model

        if text_only:  # fiber_module.py:249-276
            text_embeds = tt.embeddings(input_ids=text_ids)
            ext = tt.get_extended_attention_mask(text_masks, text_masks.size(), text_embeds.device)
            for layer in tt.encoder.layer:
                text_embeds = layer(text_embeds, ext)[0]
            text_embeds = self.cross_modal_text_transform_itc(text_embeds)
            cls = self.cross_modal_text_pooler_itc(text_embeds) if self.itc_pooler else text_embeds[:, 0]
            cls = cls / cls.norm(dim=-1, keepdim=True)
            return {"text_feats": text_embeds, "image_feats": None, "cls_feats": cls, "text_labels": text_labels,
                    "text_ids": text_ids, "text_masks": text_masks, "image": None}

        vit.draw_droppath(img.shape[0], img.device)
        image_embeds = vit.pos_drop(vit.patch_embed(img))
        if image_only:  # fiber_module.py:278-308
            for layer in vit.layers:
                image_embeds = layer(image_embeds)
            image_embeds = self.cross_modal_image_transform_itc(vit.norm(image_embeds))
            avg = image_embeds.mean(dim=1, keepdim=True)
            cls = self.cross_modal_image_pooler_itc(avg) if self.itc_pooler else avg[:, 0]
            cls = cls / cls.norm(dim=-1, keepdim=True)
            return {"text_feats": None, "image_feats": image_embeds, "cls_feats": cls, "text_labels": None,
                    "text_ids": None, "text_masks": None, "image": None}

        # fused pass, fiber_module.py:310-367
        for layer in vit.layers[:2]:
            image_embeds = layer(image_embeds)
        text_embeds = tt.embeddings(input_ids=text_ids)
        ext = tt.get_extended_attention_mask(text_masks, text_masks.size(), text_embeds.device)
        num_pre_text = self.num_text_layer - self.num_fuse_block
        for layer in tt.encoder.layer[:num_pre_text]:
            text_embeds = layer(text_embeds, ext)[0]
        num_pre_block = 8 + num_pre_text
        for blk_cnt, blk in enumerate(vit.layers[2].blocks):
            if blk_cnt < num_pre_block:
                image_embeds = blk(image_embeds)
            else:  # the two towers exchange their PREVIOUS states (:330-334)
                fuse_image_embeds = blk(image_embeds, text_embeds, ext)
                text_embeds = tt.encoder.layer[blk_cnt - 8](text_embeds, ext, encoder_hidden_states=image_embeds)[0]
                image_embeds = fuse_image_embeds
        if vit.layers[2].downsample is not None:
            image_embeds = vit.layers[2].downsample(image_embeds)
        for blk_cnt, blk in enumerate(vit.layers[3].blocks):
            fuse_image_embeds = blk(image_embeds, text_embeds, ext)
            text_embeds = tt.encoder.layer[blk_cnt + 10](text_embeds, ext, encoder_hidden_states=image_embeds,
                                                         last_norm=(blk_cnt == 0))[0]
            image_embeds = fuse_image_embeds
        text_embeds = self.cross_modal_text_transform(text_embeds)
        image_embeds = self.cross_modal_image_transform(image_embeds)
        cls_feats_text = self.cross_modal_text_pooler(text_embeds)
        avg_image_feats = image_embeds.mean(dim=1, keepdim=True)
        cls_feats_image = self.cross_modal_image_pooler(avg_image_feats)
        cls_feats = torch.cat([cls_feats_text, cls_feats_image], dim=-1)
        return {"text_feats": text_embeds, "image_feats": image_embeds, "cls_feats": cls_feats,
                "text_labels": text_labels, "text_ids": text_ids, "text_masks": text_masks, "image": img}

    def forward(self, batch):
        self.queue_sync()
        ret = dict()
        if len(self.current_tasks) == 0:
            ret.update(self.infer(batch))
            return ret
        tasks = self.current_tasks
        # MLM and hard-negative ITM share the fused backbone on independent samples: one 4B pass
        # instead of a B pass and a 3B pass (same losses; FIBER_MERGE_PASSES=0 restores the
        # reference's call-by-call order of fiber_module.py:437-451)
        merge = self.merge_mlm_itm_pass and all(t in tasks for t in ("mlm", "itc", "itm"))
        if "mlm" in tasks and not merge:
            ret.update(objectives.compute_mlm(self, batch))
        if "itc" in tasks:
            ret_itc, image_neg, text_neg, text_mask_neg = objectives.compute_itc(self, batch)
            ret.update(ret_itc)
        if merge:
            ret.update(objectives.compute_mlm_itm_hardneg_merged(self, batch, image_neg, text_neg, text_mask_neg))
        elif "itm" in tasks:
            if "itc" in tasks:
                ret.update(objectives.compute_itm_hardneg(self, batch, image_neg, text_neg, text_mask_neg))
            else:
                ret.update(objectives.compute_itm(self, batch))
        if "vqa" in self.current_tasks:
            ret.update(objectives.compute_vqa(self, batch))
        return ret

    def training_step(self, batch, batch_idx):
        fiber_utils.set_task(self)
        output = self(batch)
        return sum([v for k, v in output.items() if "loss" in k])

    def training_epoch_end(self, outs):
        fiber_utils.epoch_wrapup(self)

    def validation_step(self, batch, batch_idx):
        fiber_utils.set_task(self)
        self(batch)  # returns None like the reference (:483-485): PL keeps step outputs for the whole epoch otherwise

    def validation_epoch_end(self, outs):
        fiber_utils.epoch_wrapup(self)

    def test_step(self, batch, batch_idx):
        """fiber_module.py:490-505 minus the captioning branch (out of scope)."""
        fiber_utils.set_task(self)
        output = self(batch)
        ret = dict()
        if self.config["loss_names"]["vqa"] > 0:
            ret.update(objectives.vqa_test_step(self, batch, output))
        return ret

    def test_epoch_end(self, outs):
        fiber_utils.epoch_wrapup(self)

    def configure_optimizers(self):
        return fiber_utils.set_schedule(self)
