"""Fine-grained fused backbone: the host-side mirror of
fine_grained/maskrcnn_benchmark/modeling/backbone/fusion_swin_transformer_v2.py (FusionSwinTransformer :803-942 and the
Swin pieces it drives, :76-800) over the RoBERTa of language_backbone/roberta_fused_model_v2.py, on the sm_100a kernels
of this package.  Same module tree and parameter names as the reference (`backbone.body.layers.2.blocks.14.attn.
alpha_i2t`, `language_backbone.body.model.encoder.layer.6.crossattention_t2i...`), so its checkpoints load; same
`forward(tokenizer_input, images) -> (stage maps, language dict, None)` protocol.  FPN / DyHead / losses are the
caller's (out of scope: the maps returned are the FPN inputs).

What differs from the coarse path (modules/swin_transformer.py), as in the reference:
  * any H x W: PatchEmbed pads the image to a multiple of 4, every block zero-pads the LayerNorm'ed tokens to a multiple
    of the 12-token window and crops after the attention (:309-345).  The padded grid is what the window kernels see
    (their shift / partition / SW-MSA mask are index math on (Hp, Wp)), so the 12 x 12 tcgen05 + TMA kernels serve every
    resolution; pad and crop are torch copies between the kernel launches (ops.WindowAttnFn, CrossAttnFn, MlpFn);
  * image -> text attention takes its query from the projected window-attention output WITHOUT a LayerNorm (:199);
  * stage outputs pass norm0..norm3 and leave as fp32 NCHW maps; the text is pooled by a masked mean.
There is no CPU path: the Functions call the C-ABI kernels."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from . import roberta as R
from .swin_transformer import DropPath, FLayerNorm, FLinear, _trunc_normal_

BF16 = torch.bfloat16


class WindowAttention(nn.Module):  # :76-146 (parameters; the block runs the math)
    def __init__(self, dim, window_size, num_heads, qkv_bias=True, dim_text=None):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, (window_size, window_size), num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window_size - 1) ** 2, num_heads))
        i = torch.arange(window_size * window_size)
        hi, wi = i // window_size, i % window_size
        self.register_buffer("relative_position_index", (hi[:, None] - hi[None, :] + window_size - 1) * (2 * window_size - 1)
                             + (wi[:, None] - wi[None, :] + window_size - 1))
        self.qkv = FLinear(dim, dim * 3, bias=qkv_bias)
        self.proj = FLinear(dim, dim)
        _trunc_normal_(self.relative_position_bias_table, std=0.02)
        if dim_text is not None:
            self.qkv_text_i2t = FLinear(dim_text, dim * 2, bias=qkv_bias)
            self.qkv_i2t = FLinear(dim, dim, bias=qkv_bias)
            self.proj_i2t = FLinear(dim, dim)
            self.alpha_i2t = nn.Parameter(torch.Tensor([0]))


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = FLinear(dim, hidden)
        self.fc2 = FLinear(hidden, dim)

    def forward(self, x):
        return ops.MlpFn.apply(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)


class SwinTransformerBlock(nn.Module):  # :233-347
    def __init__(self, dim, num_heads, window_size=12, shift_size=0, mlp_ratio=4.0, qkv_bias=True, drop_path=0.0,
                 dim_text=None):
        super().__init__()
        assert dim // num_heads == 32 and window_size == 12 and shift_size in (0, 6), \
            "the window kernels of this path cover Swin-B / Swin-L geometry (head_dim 32, 12 x 12 windows)"
        self.dim, self.num_heads, self.window_size, self.shift_size = dim, num_heads, window_size, shift_size
        self.norm1 = FLayerNorm(dim)
        self.attn = WindowAttention(dim, window_size, num_heads, qkv_bias, dim_text)
        self.drop_path = DropPath(drop_path)
        self.norm2 = FLayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.H = self.W = None

    def forward(self, x, mask_matrix=None, x_text=None, mask_text=None):
        """x [B, H*W, C] bf16; mask_matrix is accepted for protocol compatibility and ignored (the SW-MSA mask is index
        math inside the kernel); x_text [B, L, 768], mask_text additive (B, 1, 1, L)."""
        B, T, C = x.shape
        H, W, ws = self.H, self.W, self.window_size
        assert T == H * W, "input feature has wrong size"
        Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
        padded = Hp != H or Wp != W
        a = self.attn
        xn = self.norm1(x)
        if padded:
            xn = F.pad(xn.view(B, H, W, C), (0, 0, 0, Wp - W, 0, Hp - H))
        qkv = a.qkv(xn.reshape(B * Hp * Wp, C))
        ao = ops.WindowAttnFn.apply(qkv, a.relative_position_bias_table, (B, Hp, Wp, ws, self.shift_size), self.num_heads)
        z = a.proj(ao)
        if x_text is not None:
            L = x_text.shape[1]
            q2 = a.qkv_i2t(z)
            kvt = a.qkv_text_i2t(x_text.reshape(B * L, -1))
            km = None if mask_text is None else mask_text.reshape(B, L).float()
            y = a.proj_i2t(ops.CrossAttnFn.apply(q2, kvt, km, B, Hp * Wp, L, self.num_heads))
            z = (z.float() + a.alpha_i2t * y.float()).to(BF16)
        s = self.drop_path.sample_scale(B, x.device)
        x = ops.CropScaleAddFn.apply(x, z, s, (H, W), (Hp, Wp))   # shortcut + drop_path(z[:, :H, :W]) in one pass
        s = self.drop_path.sample_scale(B, x.device)
        # LN2 -> fc1 -> GELU -> fc2 (+ residual, DropPath scale) as one node (no padding in this half)
        return ops.MlpResidualFn.apply(x, s, self.norm2.weight, self.norm2.bias, self.norm2.eps, self.mlp.fc1.weight,
                                       self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias)


class PatchMerging(nn.Module):  # :348-390
    def __init__(self, dim):
        super().__init__()
        self.reduction = FLinear(4 * dim, 2 * dim, bias=False)
        self.norm = FLayerNorm(4 * dim)

    def forward(self, x, H, W):
        B, T, C = x.shape
        assert T == H * W, "input feature has wrong size"
        if H % 2 or W % 2:  # the reference asserts even sizes (:370) and keeps a padding branch behind it
            x = F.pad(x.view(B, H, W, C), (0, 0, 0, W % 2, 0, H % 2)).reshape(B, -1, C)
            H, W = H + H % 2, W + W % 2
        return ops.PatchMergeFn.apply(x, self.norm.weight, self.norm.bias, self.reduction.weight, H, W)


class BasicLayer(nn.Module):  # :402-525
    def __init__(self, dim, depth, num_heads, window_size, mlp_ratio, qkv_bias, drop_path, downsample, fuse_from=None):
        super().__init__()
        self.window_size, self.shift_size, self.depth = window_size, window_size // 2, depth
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2, mlp_ratio, qkv_bias,
                                 drop_path[i], dim_text=768 if (fuse_from is not None and i >= fuse_from) else None)
            for i in range(depth)])
        self.downsample = PatchMerging(dim) if downsample else None

    def get_attention_mask(self, H, W, device):
        return None  # the reference's (nW, 144, 144) table: here the kernels derive it from (Hp, Wp, shift)

    def forward(self, x, H, W, x_text=None, mask_text=None):
        for blk in self.blocks:
            blk.H, blk.W = H, W
            x = blk(x, None, x_text, mask_text)
        if self.downsample is not None:
            return x, H, W, self.downsample(x, H, W), (H + 1) // 2, (W + 1) // 2
        return x, H, W, x, H, W


class PatchEmbed(nn.Module):  # :527-566
    def __init__(self, patch_size=4, in_chans=3, embed_dim=128):
        super().__init__()
        assert patch_size == 4 and in_chans == 3
        self.embed_dim = embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=4, stride=4)  # parameters only
        self.norm = nn.LayerNorm(embed_dim)

    def tokens(self, img):
        """[B, 3, H, W] -> ([B, Wh*Ww, C] bf16, Wh, Ww)"""
        _, _, H, W = img.shape
        if H % 4 or W % 4:
            img = F.pad(img, (0, (4 - W % 4) % 4, 0, (4 - H % 4) % 4))
        x = ops.PatchEmbedFn.apply(img, self.proj.weight, self.proj.bias, self.norm.weight, self.norm.bias)
        return x, img.shape[2] // 4, img.shape[3] // 4

    def forward(self, img):
        x, Wh, Ww = self.tokens(img)
        return x.transpose(1, 2).reshape(img.shape[0], self.embed_dim, Wh, Ww)


class SwinTransformer(nn.Module):  # :569-800 (Swin-B / Swin-L of the FIBER fine-grained configs)
    def __init__(self, embed_dim=128, depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32), window_size=12, mlp_ratio=4.0,
                 qkv_bias=True, drop_path_rate=0.2, out_features=("stage2", "stage3", "stage4", "stage5"),
                 backbone_arch="SWINT-FPN-RETINANET", num_pre_block=14, **unused):
        super().__init__()
        self.num_layers, self.embed_dim, self.ape = len(depths), embed_dim, False
        self.out_features = list(out_features)
        self.patch_embed = PatchEmbed(4, 3, embed_dim)
        self.pos_drop = nn.Identity()
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        self.num_features = [int(embed_dim * 2 ** i) for i in range(self.num_layers)]
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            fuse_from = num_pre_block if i == 2 else (0 if i == 3 else None)  # :430 `768 if i >= 14 else dim_text`
            self.layers.append(BasicLayer(self.num_features[i], depths[i], num_heads[i], window_size, mlp_ratio, qkv_bias,
                                          dpr[sum(depths[:i]):sum(depths[:i + 1])], i < self.num_layers - 1, fuse_from))
        for i in range(self.num_layers):
            if "stage%d" % (i + 2) in self.out_features:
                ident = i == 0 and backbone_arch.endswith("RETINANET")  # :703-706
                self.add_module("norm%d" % i, nn.Identity() if ident else FLayerNorm(self.num_features[i]))
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)


class _Body(nn.Module):
    def __init__(self, body):
        super().__init__()
        self.body = body


class _LangBody(nn.Module):  # RobertaFusedEncoder (roberta_fused_model_v2.py:69-100) without the HF download
    def __init__(self, model):
        super().__init__()
        self.model = model
        self.language_dim = 768

    def get_aggregated_output(self, features, input_ids, mask):
        embedded = features * mask.unsqueeze(-1).float()
        aggregate = embedded.sum(1) / mask.sum(-1).unsqueeze(-1).float()
        return {"aggregate": aggregate, "embedded": embedded, "masks": mask, "hidden": features}


def build_language_model():
    """The RoBERTa-base of roberta_fused_model_v2.py: cross-attention (and its gate) on layers >= 6 only (:479)."""
    old = R.NUM_FUSE_BLOCK, R.DIM_IMG
    R.NUM_FUSE_BLOCK, R.DIM_IMG = 6, 1024
    try:
        model = R.RobertaModel(R.RobertaConfig(), add_pooling_layer=False)
    finally:
        R.NUM_FUSE_BLOCK, R.DIM_IMG = old
    for i, layer in enumerate(model.encoder.layer):
        if i < 6 and hasattr(layer, "alpha_t2i"):
            del layer.alpha_t2i  # the coarse model creates the gate on every layer; this one only where it is used
        if hasattr(layer, "crossattention_t2i"):
            layer.crossattention_t2i.output.LayerNorm = nn.Identity()  # unused there as well (:314-316)
    return model


class FusionSwinTransformer(nn.Module):  # :803-942
    def __init__(self, vision_backbone=None, language_model=None, num_pre_text=6, num_pre_vision=2, num_pre_block=14):
        super().__init__()
        self.backbone = _Body(vision_backbone if vision_backbone is not None else SwinTransformer(num_pre_block=num_pre_block))
        self.language_backbone = _Body(_LangBody(language_model if language_model is not None else build_language_model()))
        self.num_pre_text, self.num_pre_vision, self.num_pre_block = num_pre_text, num_pre_vision, num_pre_block

    def _stage_out(self, x, i, H, W):
        body = self.backbone.body
        x = getattr(body, "norm%d" % i)(x)
        return x.float().view(-1, H, W, body.num_features[i]).permute(0, 3, 1, 2).contiguous()

    def forward(self, tokenizer_input, images):
        body, lm = self.backbone.body, self.language_backbone.body.model
        img = images.tensors if hasattr(images, "tensors") else images
        x, Wh, Ww = body.patch_embed.tokens(img)
        ids, amask = tokenizer_input["input_ids"], tokenizer_input["attention_mask"]
        t = lm.embeddings(input_ids=ids)
        em = lm.get_extended_attention_mask(amask, amask.size(), amask.device)
        outs = []
        for layer in lm.encoder.layer[:self.num_pre_text]:
            t = layer(t, em)[0]
        for i, layer in enumerate(body.layers[:self.num_pre_vision]):
            x_out, H, W, x, Wh, Ww = layer(x, Wh, Ww)
            if "stage%d" % (i + 2) in body.out_features:
                outs.append(self._stage_out(x_out, i, H, W))
        s = self.num_pre_vision
        for b, blk in enumerate(body.layers[s].blocks):
            blk.H, blk.W = Wh, Ww
            if b < self.num_pre_block:
                x = blk(x)
            else:  # both halves of a fused pair read the other modality BEFORE the pair (:891-899)
                xf = blk(x, None, t, em)
                t = lm.encoder.layer[b - self.num_pre_block + self.num_pre_text](t, em, encoder_hidden_states=x)[0]
                x = xf
        if "stage%d" % (s + 2) in body.out_features:
            outs.append(self._stage_out(x, s, Wh, Ww))
        if body.layers[s].downsample is not None:
            x = body.layers[s].downsample(x, Wh, Ww)
            Wh, Ww = (Wh + 1) // 2, (Ww + 1) // 2
        for b, blk in enumerate(body.layers[s + 1].blocks):
            blk.H, blk.W = Wh, Ww
            xf = blk(x, None, t, em)
            t = lm.encoder.layer[len(lm.encoder.layer) - len(body.layers[s + 1].blocks) + b](t, em, encoder_hidden_states=x)[0]
            x = xf
        if "stage%d" % (s + 3) in body.out_features:
            outs.append(self._stage_out(x, s + 1, Wh, Ww))
        lang = self.language_backbone.body.get_aggregated_output(t.float(), ids, amask)
        return outs, lang, None

    def load_reference_state(self, vision_sd, language_sd):
        """Load the reference's SwinTransformer / RobertaModel state_dicts (e.g. split out of a FIBER fine-grained
        checkpoint by their `backbone.body.` / `language_backbone.body.model.` prefixes)."""
        mv = self.backbone.body.load_state_dict(vision_sd, strict=False)
        ml = self.language_backbone.body.model.load_state_dict(language_sd, strict=False)
        bad = [k for k in list(mv.missing_keys) + list(ml.missing_keys) if "relative_position_index" not in k and "position_ids" not in k]
        if bad:
            raise RuntimeError("fiber_b200 fine-grained backbone: parameters missing from the reference state: %s" % bad[:8])
        return mv, ml
