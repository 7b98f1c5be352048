"""ORACLE — test infrastructure, not product code.

A plain-PyTorch fp32 restatement of the reference's fusion-in-the-backbone path
(microsoft/FIBER, coarse_grained/fiber/modules), written functionally over a `state_dict` with the
reference's parameter names.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it, and only as the checker or the timed CPU baseline.

Parity pinning: the reference ships no tests or golden vectors, so this oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF: tools/make_golden.py imports the unmodified
reference in the build container (under the import shims of baseline/ref_shims.py), runs it on
seeded inputs/weights (oracle/synth.py) and commits the results under tests/golden/;
tests/test_oracle_golden.py checks this file against those fixtures on CPU.

Each function cites the reference lines it restates (paths relative to
/root/reference/coarse_grained/fiber/modules/).  The third-party pieces the reference calls and
that are not vendored in it are restated from their published behaviour:
  timm==0.4.12         PatchEmbed (conv k=s=4 -> flatten -> LayerNorm), Mlp (fc1, GELU(erf), fc2)
  transformers==4.6.0  get_extended_attention_mask ((1-m)*-10000), BertPredictionHeadTransform
                       (dense -> gelu -> LayerNorm(eps 1e-12)), ACT2FN["gelu"] (erf form)
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# closed-form index maps (replace the reference's roll / partition / buffers)
# --------------------------------------------------------------------------------------------
def window_token_source(H, W, ws, shift):
    """[nW, N] flat source-token index (h*W + w) for window w, in-window token i.

    Restates roll(-shift) + window_partition (swin_transformer.py:99-110,366-373): window token
    (wh, ww, hi, wi) reads un-shifted location ((wh*ws+hi+shift) % H, (ww*ws+wi+shift) % W);
    window_reverse + roll(+shift) (:379-384) writes back to the same location.
    """
    hh = torch.arange(H).view(H // ws, 1, ws, 1)
    wwi = torch.arange(W).view(1, W // ws, 1, ws)
    src = ((hh + shift) % H) * W + ((wwi + shift) % W)
    return src.reshape((H // ws) * (W // ws), ws * ws)


def relative_position_index(ws):
    """[N, N] index into the (2ws-1)^2 bias table (swin_transformer.py:165-176)."""
    i = torch.arange(ws * ws)
    hi, wi = i // ws, i % ws
    return (hi[:, None] - hi[None, :] + ws - 1) * (2 * ws - 1) + (wi[:, None] - wi[None, :] + ws - 1)


def shift_attn_mask(H, W, ws, shift):
    """[nW, N, N] 0 / -100 mask of SW-MSA (swin_transformer.py:327-350), None when shift == 0."""
    if shift == 0:
        return None

    def region(c, size):
        return (c >= size - ws).long() + (c >= size - shift).long()

    hp = torch.arange(H).view(H // ws, 1, ws, 1)
    wp = torch.arange(W).view(1, W // ws, 1, ws)
    rid = (3 * region(hp, H) + region(wp, W)).reshape((H // ws) * (W // ws), ws * ws)
    return torch.where(rid[:, :, None] == rid[:, None, :], 0.0, -100.0)


# --------------------------------------------------------------------------------------------
# primitive ops
# --------------------------------------------------------------------------------------------
def _ln(x, sd, prefix, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _lin(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def _heads(x, nh):
    b, t, c = x.shape
    return x.view(b, t, nh, c // nh).transpose(1, 2)  # [B, nH, T, d]


def _merge(x):
    b, nh, t, d = x.shape
    return x.transpose(1, 2).reshape(b, t, nh * d)


# --------------------------------------------------------------------------------------------
# Swin tower
# --------------------------------------------------------------------------------------------
def patch_embed(img, sd, prefix="vit_model.patch_embed"):
    """timm PatchEmbed as used at fiber_module.py:311 / swin_transformer.py:588-594.

    A 4x4/stride-4 conv is a per-patch linear map over the 48 values (c, kh, kw)."""
    B, C, H, W = img.shape
    w = sd[prefix + ".proj.weight"]
    p = w.shape[-1]
    x = img.view(B, C, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // p) * (W // p), C * p * p)
    x = F.linear(x, w.reshape(w.shape[0], -1), sd[prefix + ".proj.bias"])
    return _ln(x, sd, prefix + ".norm")


def window_attention(x, sd, prefix, H, W, ws, shift, nh, text=None, text_mask=None):
    """W-MSA / SW-MSA (+ optional image->text cross attention) on image-ordered tokens.

    x: [B, H*W, C] (already LayerNorm'ed).  Restates swin_transformer.py:195-261 together with the
    shift/partition/reverse plumbing of :363-387.  Returns [B, H*W, C] in image order.
    """
    B, T, C = x.shape
    N = ws * ws
    src = window_token_source(H, W, ws, shift).to(x.device)  # [nW, N]
    nW = src.shape[0]
    xw = x[:, src.reshape(-1)].reshape(B * nW, N, C)
    qkv = _lin(xw, sd, prefix + ".qkv").view(B * nW, N, 3, nh, C // nh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    scale = (C // nh) ** -0.5
    s = (q * scale) @ k.transpose(-2, -1)
    table = sd[prefix + ".relative_position_bias_table"]
    bias = table[relative_position_index(ws).to(x.device).reshape(-1)].view(N, N, nh).permute(2, 0, 1)
    s = s + bias.unsqueeze(0)
    m = shift_attn_mask(H, W, ws, shift)
    if m is not None:
        s = (s.view(B, nW, nh, N, N) + m.to(x.device)[None, :, None]).view(B * nW, nh, N, N)
    o = _merge(torch.softmax(s, dim=-1) @ v)
    o = _lin(o, sd, prefix + ".proj")  # window order
    out = torch.empty((B, T, C), dtype=o.dtype, device=x.device)
    out[:, src.reshape(-1)] = o.view(B, nW * N, C)
    if text is not None:
        # i2t (:226-259): per-token query against the sample's own text keys; window layout is
        # irrelevant for a token-wise op, so it is evaluated in image order.
        kv = _lin(text, sd, prefix + ".qkv_text_i2t")
        L = text.shape[1]
        kv = kv.view(B, L, 2, nh, C // nh).permute(2, 0, 3, 1, 4)
        kt, vt = kv[0], kv[1]
        qi = _heads(_lin(_ln(out, sd, prefix + ".norm_i2t_i"), sd, prefix + ".qkv_i2t"), nh) * scale
        s2 = qi @ kt.transpose(-2, -1)
        if text_mask is not None:
            s2 = s2 + text_mask.view(B, 1, 1, L)
        y = _lin(_merge(torch.softmax(s2, dim=-1) @ vt), sd, prefix + ".proj_i2t")
        out = out + sd[prefix + ".alpha_i2t"] * y
    return out


def swin_block(x, sd, prefix, H, W, ws, shift, nh, text=None, text_mask=None):
    """SwinTransformerBlock.forward (swin_transformer.py:356-393), DropPath disabled."""
    if min(H, W) <= ws:  # :304-307
        shift, ws = 0, min(H, W)
    a = window_attention(_ln(x, sd, prefix + ".norm1"), sd, prefix + ".attn", H, W, ws, shift, nh, text, text_mask)
    x = x + a
    h = F.gelu(_lin(_ln(x, sd, prefix + ".norm2"), sd, prefix + ".mlp.fc1"))
    return x + _lin(h, sd, prefix + ".mlp.fc2")


def patch_merging(x, sd, prefix, H, W):
    """PatchMerging.forward (swin_transformer.py:411-432)."""
    B, T, C = x.shape
    x = x.view(B, H // 2, 2, W // 2, 2, C)
    # channel order [x(0,0), x(1,0), x(0,1), x(1,1)]: row-odd before column-odd (:422-426)
    x = x.permute(0, 1, 3, 4, 2, 5).reshape(B, (H // 2) * (W // 2), 4 * C)
    return F.linear(_ln(x, sd, prefix + ".norm"), sd[prefix + ".reduction.weight"])


class SwinSpec:
    """Static structure of swin_base_patch4_window*_in22k as FIBER builds it
    (swin_transformer.py:575,608-632,756-771; BasicLayer :488-506)."""

    def __init__(self, image_size, num_fuse_block=6, embed_dim=128, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32)):
        self.ws = int(image_size / 32)
        self.grid = image_size // 4
        self.depths, self.heads, self.embed_dim = depths, heads, embed_dim
        self.num_fuse_block = num_fuse_block

    def stage(self, s):
        g = self.grid // (2 ** s)
        return g, g, self.embed_dim * (2 ** s), self.heads[s]

    def shift(self, i):
        return 0 if i % 2 == 0 else self.ws // 2

    def fused(self, s, i):
        return s == 3 or (s == 2 and i >= 20 - self.num_fuse_block)


def swin_stage(x, sd, spec, s, text=None, text_mask=None, blocks=None, downsample=True):
    H, W, C, nh = spec.stage(s)
    for i in (range(spec.depths[s]) if blocks is None else blocks):
        x = swin_block(x, sd, "vit_model.layers.%d.blocks.%d" % (s, i), H, W, spec.ws, spec.shift(i), nh, text, text_mask)
    if downsample and s < 3:
        x = patch_merging(x, sd, "vit_model.layers.%d.downsample" % s, H, W)
    return x


# --------------------------------------------------------------------------------------------
# RoBERTa tower
# --------------------------------------------------------------------------------------------
def roberta_position_ids(ids, pad=1):
    """create_position_ids_from_input_ids (roberta.py:877-888)."""
    m = ids.ne(pad).int()
    return (torch.cumsum(m, dim=1).type_as(m) * m).long() + pad


def roberta_embeddings(ids, sd, prefix="text_transformer.embeddings", pad=1):
    """RobertaEmbeddings.forward (roberta.py:169-199), dropout disabled; token type always 0."""
    # nn.Embedding(padding_idx=1) on both tables (:150,165): the pad row receives no gradient
    e = F.embedding(ids, sd[prefix + ".word_embeddings.weight"], padding_idx=pad)
    e = e + sd[prefix + ".token_type_embeddings.weight"][0]
    e = e + F.embedding(roberta_position_ids(ids, pad), sd[prefix + ".position_embeddings.weight"], padding_idx=pad)
    return _ln(e, sd, prefix + ".LayerNorm")


def extended_mask(text_masks):
    """HF 4.6 get_extended_attention_mask: 0 for real tokens, -10000 for padding."""
    return (1.0 - text_masks[:, None, None, :].float()) * -10000.0


def _bert_attention(h_q, h_kv, sd, prefix, nh, mask):
    """RobertaSelfAttention + RobertaSelfOutput.dense (roberta.py:256-326,337-340); scores are
    divided by sqrt(d) AFTER QK^T (:302)."""
    q = _heads(_lin(h_q, sd, prefix + ".self.query"), nh)
    k = _heads(_lin(h_kv, sd, prefix + ".self.key"), nh)
    v = _heads(_lin(h_kv, sd, prefix + ".self.value"), nh)
    s = q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1])
    if mask is not None:
        s = s + mask
    return _lin(_merge(torch.softmax(s, dim=-1) @ v), sd, prefix + ".output.dense")


def roberta_layer(h, ext_mask, sd, i, image=None, last_norm=True, nh=12):
    """RobertaLayer.forward (roberta.py:441-502), dropout disabled."""
    p = "text_transformer.encoder.layer.%d" % i
    a = _bert_attention(h, h, sd, p + ".attention", nh, ext_mask)
    if image is not None:
        c = _bert_attention(a, image, sd, p + ".crossattention_t2i", nh, None)  # no mask on image keys
        a = sd[p + ".alpha_t2i"] * c + a
    a = _ln(a + h, sd, p + ".attention.output.LayerNorm")
    f = _lin(F.gelu(_lin(a, sd, p + ".intermediate.dense")), sd, p + ".output.dense") + a
    return _ln(f, sd, p + ".output.LayerNorm") if last_norm else f


# --------------------------------------------------------------------------------------------
# heads (callers of the path; restated so a full training_step can be checked)
# --------------------------------------------------------------------------------------------
def pooler(x, sd, prefix):
    return torch.tanh(_lin(x[:, 0], sd, prefix + ".dense"))  # heads.py:8-18


def mlm_head(x, sd, prefix="mlm_score"):
    """heads.MLMHead (heads.py:31-43): transform LN eps is the HF default 1e-12."""
    h = _ln(F.gelu(_lin(x, sd, prefix + ".transform.dense")), sd, prefix + ".transform.LayerNorm", eps=1e-12)
    return F.linear(h, sd[prefix + ".decoder.weight"]) + sd[prefix + ".bias"]


# --------------------------------------------------------------------------------------------
# FIBERTransformerSS.infer (fiber_module.py:224-367)
# --------------------------------------------------------------------------------------------
def infer(sd, cfg, img=None, text_ids=None, text_masks=None, image_only=False, text_only=False):
    spec = SwinSpec(cfg["image_size"], cfg["num_fuse_block"])
    nl, nf = cfg["num_layers"], cfg["num_fuse_block"]
    if text_only:  # :249-276
        t = roberta_embeddings(text_ids, sd)
        em = extended_mask(text_masks)
        for i in range(nl):
            t = roberta_layer(t, em, sd, i)
        t = _lin(t, sd, "cross_modal_text_transform_itc")
        c = pooler(t, sd, "cross_modal_text_pooler_itc") if cfg["itc_pooler"] else t[:, 0]
        return {"text_feats": t, "image_feats": None, "cls_feats": c / c.norm(dim=-1, keepdim=True)}
    x = patch_embed(img, sd)
    if image_only:  # :278-308
        for s in range(4):
            x = swin_stage(x, sd, spec, s)
        x = _lin(_ln(x, sd, "vit_model.norm"), sd, "cross_modal_image_transform_itc")
        avg = x.mean(dim=1, keepdim=True)
        c = pooler(avg, sd, "cross_modal_image_pooler_itc") if cfg["itc_pooler"] else avg[:, 0]
        return {"text_feats": None, "image_feats": x, "cls_feats": c / c.norm(dim=-1, keepdim=True)}
    # fused pass :310-367
    x = swin_stage(x, sd, spec, 0)
    x = swin_stage(x, sd, spec, 1)
    t = roberta_embeddings(text_ids, sd)
    em = extended_mask(text_masks)
    n_pre_text = nl - nf
    for i in range(n_pre_text):
        t = roberta_layer(t, em, sd, i)
    n_pre_block = 8 + n_pre_text
    H, W, C, nh = spec.stage(2)
    for b in range(spec.depths[2]):
        pre = "vit_model.layers.2.blocks.%d" % b
        if b < n_pre_block:
            x = swin_block(x, sd, pre, H, W, spec.ws, spec.shift(b), nh)
        else:  # both towers read the OLD state of the other one (:330-334)
            x_new = swin_block(x, sd, pre, H, W, spec.ws, spec.shift(b), nh, t, em)
            t = roberta_layer(t, em, sd, b - 8, image=x)
            x = x_new
    x = patch_merging(x, sd, "vit_model.layers.2.downsample", H, W)
    H, W, C, nh = spec.stage(3)
    for b in range(spec.depths[3]):
        x_new = swin_block(x, sd, "vit_model.layers.3.blocks.%d" % b, H, W, spec.ws, spec.shift(b), nh, t, em)
        t = roberta_layer(t, em, sd, b + 10, image=x, last_norm=(b == 0))
        x = x_new
    t = _lin(t, sd, "cross_modal_text_transform")
    x = _lin(x, sd, "cross_modal_image_transform")
    ct = pooler(t, sd, "cross_modal_text_pooler")
    ci = pooler(x.mean(dim=1, keepdim=True), sd, "cross_modal_image_pooler")
    return {"text_feats": t, "image_feats": x, "cls_feats": torch.cat([ct, ci], dim=-1)}


# --------------------------------------------------------------------------------------------
# objectives (objectives.py) — the parts of a training_step that surround infer()
# --------------------------------------------------------------------------------------------
def compute_mlm(sd, cfg, batch):
    """objectives.compute_mlm (:17-41)."""
    out = infer(sd, cfg, batch["image"][0], batch["text_ids_mlm"], batch["text_masks"])
    logits = mlm_head(out["text_feats"], sd)
    loss = F.cross_entropy(logits.view(-1, cfg["vocab_size"]), batch["text_labels_mlm"].view(-1), ignore_index=-100)
    return {"mlm_loss": loss, "mlm_logits": logits}


def compute_itm(sd, cfg, batch, itm_labels):
    """objectives.compute_itm (:44-75) with the (random) label permutation supplied."""
    pick = itm_labels.view(-1, 1, 1, 1) == 1
    img = torch.where(pick, batch["image"][0], batch["false_image_0"][0])
    out = infer(sd, cfg, img, batch["text_ids"], batch["text_masks"])
    logits = _lin(out["cls_feats"], sd, "itm_score.fc")
    return {"itm_loss": F.cross_entropy(logits, itm_labels.long()), "itm_logits": logits}


def compute_itc(sd, cfg, batch, queue_total=0):
    """objectives.compute_itc (:119-180): loss + hard-negative sampling weights."""
    fi = infer(sd, cfg, img=batch["image"][0], image_only=True)["cls_feats"]
    ft = infer(sd, cfg, text_ids=batch["text_ids"], text_masks=batch["text_masks"], text_only=True)["cls_feats"]
    temp = sd["temp"].clamp(0.001, 1.0)
    fi_all = torch.cat([fi.t().detach(), sd["image_queue"]], dim=1)
    ft_all = torch.cat([ft.t().detach(), sd["text_queue"]], dim=1)
    sim_i2t = fi @ ft_all / temp
    sim_t2i = ft @ fi_all / temp
    tgt = torch.zeros_like(sim_i2t)
    tgt.fill_diagonal_(1)
    loss = (-(F.log_softmax(sim_i2t, 1) * tgt).sum(1).mean() - (F.log_softmax(sim_t2i, 1) * tgt).sum(1).mean()) / 2
    bs = fi.shape[0]
    with torch.no_grad():
        w_i2t = F.softmax(sim_i2t[:, : bs + queue_total], dim=1)
        w_t2i = F.softmax(sim_t2i[:, : bs + queue_total], dim=1)
        w_i2t.fill_diagonal_(0)
        w_t2i.fill_diagonal_(0)
    return {"itc_loss": loss, "image_feat": fi, "text_feat": ft, "weights_i2t": w_i2t, "weights_t2i": w_t2i}


def compute_itm_hardneg(sd, cfg, batch, image_neg, text_neg, text_mask_neg):
    """objectives.compute_itm_hardneg (:78-116): 3B samples = pos, (img, neg text), (neg img, text)."""
    B = batch["text_ids"].shape[0]
    img = torch.cat([batch["image"][0], batch["image"][0], image_neg], 0)
    ids = torch.cat([batch["text_ids"], text_neg, batch["text_ids"]], 0)
    msk = torch.cat([batch["text_masks"], text_mask_neg, batch["text_masks"]], 0)
    out = infer(sd, cfg, img, ids, msk)
    logits = _lin(out["cls_feats"], sd, "itm_score.fc")
    labels = torch.cat([torch.ones(B), torch.zeros(2 * B)]).long().to(logits.device)
    return {"itm_loss": F.cross_entropy(logits, labels), "itm_logits": logits}


def compute_vqa(sd, cfg, batch):
    """objectives.compute_vqa (:182-215)."""
    out = infer(sd, cfg, batch["image"][0], batch["text_ids"], batch["text_masks"])
    h = _lin(out["cls_feats"], sd, "vqa_classifier.0")
    h = F.gelu(_ln(h, sd, "vqa_classifier.1"))
    logits = _lin(h, sd, "vqa_classifier.3")
    tgt = torch.zeros_like(logits)
    for i, (ls, ss) in enumerate(zip(batch["vqa_labels"], batch["vqa_scores"])):
        for l, s in zip(ls, ss):
            tgt[i, l] = s
    return {"vqa_loss": F.binary_cross_entropy_with_logits(logits, tgt) * tgt.shape[1], "vqa_logits": logits}
