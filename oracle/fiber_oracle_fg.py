"""ORACLE (fine-grained fused backbone) — test infrastructure, not product code.

Plain-PyTorch fp32 restatement of the FUSED BACKBONE of FIBER's fine-grained model:
`fine_grained/maskrcnn_benchmark/modeling/backbone/fusion_swin_transformer_v2.py` (FusionSwinTransformer.forward
:817-942 and the Swin pieces it calls) over `language_backbone/roberta_fused_model_v2.py`.  Same math as the coarse
path (oracle/fiber_oracle.py) with these differences, each restated here:
  * dynamic H x W: PatchEmbed pads the image to a multiple of 4 (:551-557); every block pads the LayerNorm'ed tokens
    with zeros to a multiple of the window (12) before the shift / partition and crops after the reverse (:309-345);
    the SW-MSA mask is built on the PADDED grid (BasicLayer.get_attention_mask :471-495);
  * the image->text branch has NO LayerNorm in front of its query projection: q = qkv_i2t(proj(window attention))
    (:199-200) — the coarse model normalises first (norm_i2t_i);
  * stage outputs go through norm0..norm3 and leave as NCHW maps (the FPN inputs, :873-878,905-912,935-945);
  * the text tower always applies the final LayerNorm of a layer (no last_norm flag) and the text is pooled by a
    masked mean (RobertaFusedEncoder.get_aggregated_output, roberta_fused_model_v2.py:86-100).
FPN / DyHead / the detection losses are out of scope.

State-dict names are the reference's with the two towers prefixed as in the coarse oracle, so its functions are
reused: "vit_model." + SwinTransformer keys, "text_transformer." + RobertaModel keys.

Parity pinning: tests/golden/fg_fused_backbone.pt holds outputs of the UNMODIFIED reference modules
(FusionSwinTransformer.forward run in the build container by tools/make_golden_fg.py on synthetic weights);
tests/test_oracle_golden.py checks this file against it on CPU.
"""
import torch
import torch.nn.functional as F

from . import fiber_oracle as O

WS = 12


def patch_embed(img, sd, prefix="vit_model.patch_embed"):
    """PatchEmbed.forward (fusion_swin_transformer_v2.py:548-566): pad right / bottom to a multiple of 4, conv, LN."""
    _, _, H, W = img.shape
    if W % 4 or H % 4:
        img = F.pad(img, (0, (4 - W % 4) % 4, 0, (4 - H % 4) % 4))
    return O.patch_embed(img, sd, prefix), img.shape[2] // 4, img.shape[3] // 4


def window_attention(x, sd, prefix, Hp, Wp, shift, nh, text=None, text_mask=None):
    """WindowAttention.forward (:148-230) on the PADDED grid, tokens in image order [B, Hp*Wp, C]."""
    B, T, C = x.shape
    N = WS * WS
    src = O.window_token_source(Hp, Wp, WS, shift).to(x.device)
    nW = src.shape[0]
    xw = x[:, src.reshape(-1)].reshape(B * nW, N, C)
    qkv = O._lin(xw, sd, prefix + ".qkv").view(B * nW, N, 3, nh, C // nh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    scale = (C // nh) ** -0.5
    s = (q * scale) @ k.transpose(-2, -1)
    table = sd[prefix + ".relative_position_bias_table"]
    s = s + table[O.relative_position_index(WS).to(x.device).reshape(-1)].view(N, N, nh).permute(2, 0, 1).unsqueeze(0)
    m = O.shift_attn_mask(Hp, Wp, WS, shift)
    if m is not None:
        s = (s.view(B, nW, nh, N, N) + m.to(x.device)[None, :, None]).view(B * nW, nh, N, N)
    o = O._lin(O._merge(torch.softmax(s, dim=-1) @ v), sd, prefix + ".proj")
    out = torch.empty((B, T, C), dtype=o.dtype, device=x.device)
    out[:, src.reshape(-1)] = o.view(B, nW * N, C)
    if text is not None:  # :188-228: the query is the projected window-attention output itself
        L = text.shape[1]
        kv = O._lin(text, sd, prefix + ".qkv_text_i2t").view(B, L, 2, nh, C // nh).permute(2, 0, 3, 1, 4)
        qi = O._heads(O._lin(out, sd, prefix + ".qkv_i2t"), nh) * scale
        s2 = qi @ kv[0].transpose(-2, -1)
        if text_mask is not None:
            s2 = s2 + text_mask.view(B, 1, 1, L)
        y = O._lin(O._merge(torch.softmax(s2, dim=-1) @ kv[1]), sd, prefix + ".proj_i2t")
        out = out + sd[prefix + ".alpha_i2t"] * y
    return out


def swin_block(x, sd, prefix, H, W, shift, nh, text=None, text_mask=None):
    """SwinTransformerBlock.forward (:296-347), DropPath disabled."""
    B, T, C = x.shape
    Hp, Wp = -(-H // WS) * WS, -(-W // WS) * WS
    xn = O._ln(x, sd, prefix + ".norm1").view(B, H, W, C)
    xn = F.pad(xn, (0, 0, 0, Wp - W, 0, Hp - H)).reshape(B, Hp * Wp, C)
    a = window_attention(xn, sd, prefix + ".attn", Hp, Wp, shift, nh, text, text_mask)
    x = x + a.view(B, Hp, Wp, C)[:, :H, :W].reshape(B, T, C)
    h = F.gelu(O._lin(O._ln(x, sd, prefix + ".norm2"), sd, prefix + ".mlp.fc1"))
    return x + O._lin(h, sd, prefix + ".mlp.fc2")


def stage_out(x, sd, s, H, W):
    """norm{s} + NCHW (:873-878); norm0 is nn.Identity for the *-RETINANET backbone variants (:703-706)."""
    B, _, C = x.shape
    if "vit_model.norm%d.weight" % s in sd:
        x = O._ln(x, sd, "vit_model.norm%d" % s)
    return x.view(B, H, W, C).permute(0, 3, 1, 2).contiguous()


def fused_backbone(sd, img, ids, mask, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32), num_pre_block=14, num_pre_text=6):
    """FusionSwinTransformer.forward (:817-942) up to the FPN: returns (four NCHW stage maps, text dict)."""
    x, H, W = patch_embed(img, sd)
    t = O.roberta_embeddings(ids, sd)
    em = O.extended_mask(mask)
    outs = []
    for i in range(num_pre_text):
        t = O.roberta_layer(t, em, sd, i)
    for s in range(2):
        for b in range(depths[s]):
            x = swin_block(x, sd, "vit_model.layers.%d.blocks.%d" % (s, b), H, W, 0 if b % 2 == 0 else WS // 2, heads[s])
        outs.append(stage_out(x, sd, s, H, W))
        x = O.patch_merging(x, sd, "vit_model.layers.%d.downsample" % s, H, W)
        H, W = (H + 1) // 2, (W + 1) // 2
    for b in range(depths[2]):
        pre = "vit_model.layers.2.blocks.%d" % b
        shift = 0 if b % 2 == 0 else WS // 2
        if b < num_pre_block:
            x = swin_block(x, sd, pre, H, W, shift, heads[2])
        else:  # the image block and the text layer of a fused pair both read the un-fused other modality (:891-899)
            xf = swin_block(x, sd, pre, H, W, shift, heads[2], t, em)
            t = O.roberta_layer(t, em, sd, b - num_pre_block + num_pre_text, image=x)
            x = xf
    outs.append(stage_out(x, sd, 2, H, W))
    x = O.patch_merging(x, sd, "vit_model.layers.2.downsample", H, W)
    H, W = (H + 1) // 2, (W + 1) // 2
    for b in range(depths[3]):
        xf = swin_block(x, sd, "vit_model.layers.3.blocks.%d" % b, H, W, 0 if b % 2 == 0 else WS // 2, heads[3], t, em)
        t = O.roberta_layer(t, em, sd, 10 + b, image=x)
        x = xf
    outs.append(stage_out(x, sd, 3, H, W))
    m = mask.to(t.dtype)
    embedded = t * m.unsqueeze(-1)
    lang = {"aggregate": embedded.sum(1) / m.sum(-1, keepdim=True), "embedded": embedded, "masks": mask, "hidden": t}
    return outs, lang
