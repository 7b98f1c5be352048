"""Swin-Base tower with FIBER's image->text cross attention — host-side mirror of the reference
module tree (coarse_grained/fiber/modules/swin_transformer.py): same class names, constructor
arguments, sub-module / parameter / buffer names and shapes (so reference checkpoints load and
fiber_utils.set_schedule's name-substring groups are unchanged), but every forward is a sequence
of sm_100a kernels (fiber_b200.ops).  timm's PatchEmbed / Mlp / DropPath (timm==0.4.12, not
vendored in the reference) are restated here with timm's attribute names.
"""
import torch
import torch.nn as nn

from .. import ops

DIM_TEXT = 768
NUM_FUSE_BLOCK = 6


def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


class FLinear(nn.Linear):
    """nn.Linear whose forward/backward run on the tcgen05 GEMM."""
    out_fp32 = False

    def forward(self, x):
        return ops.LinearFn.apply(x, self.weight, self.bias, self.out_fp32)


class FLayerNorm(nn.LayerNorm):
    def forward(self, x):
        return ops.LayerNormFn.apply(x, self.weight, self.bias, self.eps)


class DropPath(nn.Module):
    """timm DropPath: per-sample stochastic depth.  The block asks for the per-sample scale
    (0 or 1/keep) and fuses it into its GEMM epilogues."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def sample_scale(self, batch, device):
        if self.drop_prob == 0.0 or not self.training:
            return None
        keep = 1.0 - self.drop_prob
        return torch.floor(keep + torch.rand(batch, device=device)) / keep

    def forward(self, x):
        s = self.sample_scale(x.shape[0], x.device)
        return x if s is None else x * s.view(-1, *([1] * (x.dim() - 1))).to(x.dtype)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = FLinear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = FLinear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        assert patch_size == 4 and in_chans == 3, "the B200 path implements FIBER's 4x4 RGB patch embedding"
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        return ops.PatchEmbedFn.apply(x, self.proj.weight, self.proj.bias, self.norm.weight, self.norm.bias)


class WindowAttention(nn.Module):
    """Parameter container of W-MSA/SW-MSA (+ i2t); computed inside SwinBlockFn."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, attn_drop=0.0, proj_drop=0.0, dim_text=None,
                 norm_layer=nn.LayerNorm):
        super().__init__()
        assert attn_drop == 0.0 and proj_drop == 0.0, "FIBER uses attn_drop = proj_drop = 0 in the Swin tower"
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.scale = (dim // num_heads) ** -0.5
        ws = window_size[0]
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), num_heads))
        i = torch.arange(ws * ws)
        hi, wi = i // ws, i % ws
        idx = (hi[:, None] - hi[None, :] + ws - 1) * (2 * ws - 1) + (wi[:, None] - wi[None, :] + ws - 1)
        self.register_buffer("relative_position_index", idx)
        self.qkv = FLinear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = FLinear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        _trunc_normal_(self.relative_position_bias_table)
        self.softmax = nn.Softmax(dim=-1)
        self.has_i2t = dim_text is not None
        if dim_text is not None:
            self.qkv_text_i2t = FLinear(dim_text, dim * 2, bias=qkv_bias)
            self.qkv_i2t = FLinear(dim, dim, bias=qkv_bias)
            self.attn_drop_i2t = nn.Dropout(attn_drop)
            self.proj_i2t = FLinear(dim, dim)
            self.proj_drop_i2t = nn.Dropout(proj_drop)
            self.alpha_i2t = nn.Parameter(torch.Tensor([0]))
            self.norm_i2t_i = norm_layer(dim)


def _shift_mask(H, W, ws, shift):
    def region(c, size):
        return (c >= size - ws).long() + (c >= size - shift).long()
    hp = torch.arange(H).view(H // ws, 1, ws, 1)
    wp = torch.arange(W).view(1, W // ws, 1, ws)
    rid = (3 * region(hp, H) + region(wp, W)).reshape((H // ws) * (W // ws), ws * ws)
    return torch.where(rid[:, :, None] == rid[:, None, :], 0.0, -100.0)


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4.0, qkv_bias=True,
                 drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, dim_text=None):
        super().__init__()
        assert drop == 0.0, "FIBER uses drop = 0 in the Swin tower"
        self.dim, self.input_resolution, self.num_heads = dim, input_resolution, num_heads
        self.window_size, self.shift_size, self.mlp_ratio = window_size, shift_size, mlp_ratio
        if min(self.input_resolution) <= self.window_size:
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, (self.window_size, self.window_size), num_heads, qkv_bias=qkv_bias,
                                    attn_drop=attn_drop, proj_drop=drop, dim_text=dim_text, norm_layer=norm_layer)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        H, W = self.input_resolution
        # kept for state_dict compatibility; the kernels regenerate the mask from region ids
        self.register_buffer("attn_mask", _shift_mask(H, W, self.window_size, self.shift_size)
                             if self.shift_size > 0 else None)

    def _params(self, fused):
        names = ops.SWIN_FUSED if fused else ops.SWIN_PLAIN
        out = []
        for n in names:
            obj = self
            for part in n.split("."):
                obj = getattr(obj, part)
            out.append(obj)
        return out

    def forward(self, x, y=None, y_mask=None):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        fused = y is not None
        if fused:
            assert self.attn.has_i2t, "text given to a block without image-to-text attention"
            assert y.shape[0] == B, "B_ is not a multiplier of B_text in window attention"
        s1 = s2 = None
        if isinstance(self.drop_path, DropPath):
            pre = self.__dict__.pop("_dp_scales", None)  # drawn in bulk by SwinTransformer.draw_droppath
            if pre is not None and pre[0].shape[0] == B and self.training:
                s1, s2 = pre
            else:
                s1 = self.drop_path.sample_scale(B, x.device)
                s2 = self.drop_path.sample_scale(B, x.device)
        mask2d = None
        if fused and y_mask is not None:
            mask2d = y_mask.reshape(B, -1).float()
        geom = (H, W, self.window_size, self.shift_size, self.num_heads)
        return ops.SwinBlockFn.apply(x, y if fused else None, mask2d, s1, s2, geom, *self._params(fused))


class PatchMerging(nn.Module):
    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution, self.dim = input_resolution, dim
        self.reduction = FLinear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
        return ops.PatchMergeFn.apply(x, self.norm.weight, self.norm.bias, self.reduction.weight, H, W)


class BasicLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4.0, qkv_bias=True, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False,
                 dim_text=None, layer_index=0):
        super().__init__()
        self.dim, self.input_resolution, self.depth = dim, input_resolution, depth
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(
                dim=dim, input_resolution=input_resolution, num_heads=num_heads, window_size=window_size,
                shift_size=0 if (i % 2 == 0) else window_size // 2, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop,
                attn_drop=attn_drop, drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                norm_layer=norm_layer,
                dim_text=None if layer_index == 2 and i < 20 - NUM_FUSE_BLOCK else dim_text)
            for i in range(depth)])
        self.downsample = downsample(input_resolution, dim=dim, norm_layer=norm_layer) if downsample is not None else None

    def forward(self, x, y=None, y_mask=None):
        for blk in self.blocks:
            x = blk(x, y, y_mask)
        if self.downsample is not None:
            x = self.downsample(x)
        return x


class SwinTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4.0, qkv_bias=True, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.1, norm_layer=FLayerNorm, ape=False, patch_norm=True,
                 use_checkpoint=False, weight_init="", **kwargs):
        super().__init__()
        window_size = int(img_size / 32)  # swin_transformer.py:575 overrides the factory's value
        self.num_classes, self.num_layers = num_classes, len(depths)
        self.embed_dim, self.ape, self.patch_norm = embed_dim, ape, patch_norm
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.mlp_ratio = mlp_ratio
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      norm_layer=norm_layer if patch_norm else None)
        self.patch_grid = self.patch_embed.grid_size
        assert not ape, "absolute position embedding is unused by FIBER"
        self.absolute_pos_embed = None
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        layers = []
        for i_layer in range(self.num_layers):
            layers.append(BasicLayer(
                dim=int(embed_dim * 2 ** i_layer),
                input_resolution=(self.patch_grid[0] // (2 ** i_layer), self.patch_grid[1] // (2 ** i_layer)),
                depth=depths[i_layer], num_heads=num_heads[i_layer], window_size=window_size, mlp_ratio=mlp_ratio,
                qkv_bias=qkv_bias, drop=drop_rate, attn_drop=attn_drop_rate,
                drop_path=dpr[sum(depths[:i_layer]):sum(depths[:i_layer + 1])], norm_layer=norm_layer,
                downsample=PatchMerging if (i_layer < self.num_layers - 1) else None, use_checkpoint=use_checkpoint,
                dim_text=DIM_TEXT if i_layer >= 2 else None, layer_index=i_layer))
        self.layers = nn.Sequential(*layers)
        self.norm = norm_layer(self.num_features)
        self.avgpool = nn.AdaptiveAvgPool1d(1)
        self.apply(_init_vit_weights)

    def draw_droppath(self, batch, device):
        """Draw the two independent per-sample DropPath scales of every block of one pass in a single
        batched op (same distribution as timm's DropPath: floor(keep + U) / keep per sample)."""
        blocks = [b for layer in self.layers for b in layer.blocks if isinstance(b.drop_path, DropPath)
                  and b.drop_path.drop_prob > 0.0]
        if not self.training or not blocks:
            return
        keep = torch.tensor([1.0 - b.drop_path.drop_prob for b in blocks for _ in range(2)], device=device)
        scales = torch.floor(keep[:, None] + torch.rand(len(keep), batch, device=device)) / keep[:, None]
        for i, b in enumerate(blocks):
            b.__dict__["_dp_scales"] = (scales[2 * i], scales[2 * i + 1])

    def forward_features(self, x, y=None, y_mask=None):
        self.draw_droppath(x.shape[0], x.device)
        x = self.patch_embed(x)
        x = self.pos_drop(x)
        for layer in self.layers:
            x = layer(x, y, y_mask)
        return self.norm(x)

    def forward(self, x, y=None, y_mask=None):
        return self.forward_features(x, y, y_mask)


def _init_vit_weights(m):
    """timm 0.4.12 `_init_vit_weights` (non-jax path): Linear trunc-normal(.02)/zero bias, LN 1/0."""
    if isinstance(m, nn.Linear):
        _trunc_normal_(m.weight)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, nn.LayerNorm):
        nn.init.zeros_(m.bias)
        nn.init.ones_(m.weight)


def swin_adapt_position_encoding(state_dict, before=384, patch_size=32, after=384,
                                 suffix="relative_position_bias_table"):
    """swin_helpers.py:20-44 — a checkpoint trained at `before` pixels loaded at `after` pixels: window = size / 32, so
    every (2w-1)^2-row relative-position table is resized bicubically to the new (2w'-1)^2 grid and the resolution-
    dependent buffers (attn_mask, relative_position_index) are dropped from the dict (the model rebuilds them).
    Mutates and returns `state_dict`."""
    if after == before:
        return state_dict
    side_before, side_after = 2 * int(before / 32) - 1, 2 * int(after / 32) - 1
    tables = [k for k in state_dict if k.endswith(suffix)]
    if not tables:
        raise AssertionError("no %s entries in the checkpoint" % suffix)
    for k in tables:
        t = state_dict[k]                                               # [(2w-1)^2, heads]
        grid = t.t().reshape(1, -1, side_before, side_before)            # heads as channels
        grid = torch.nn.functional.interpolate(grid, size=(side_after, side_after), mode="bicubic")
        state_dict[k] = grid[0].permute(1, 2, 0).reshape(side_after * side_after, -1).contiguous()
    for k in [k for k in state_dict if k.endswith(("attn_mask", "relative_position_index"))]:
        del state_dict[k]
    return state_dict


def _create(pretrained=False, **kwargs):
    config = kwargs.pop("config")
    kwargs.pop("num_classes", None)
    model = SwinTransformer(img_size=config["image_size"], **kwargs)
    if pretrained:
        # timm downloads the ImageNet-22k weights here (swin_helpers.py:183-261); this image has no network, so they
        # must be supplied as a local file (a timm / official Swin checkpoint: {"model": state_dict} or a bare dict)
        import os
        path = os.environ.get("FIBER_SWIN_WEIGHTS", "")
        if not path or not os.path.exists(path):
            raise RuntimeError("pretrained_vit=True needs the Swin-B weights as a local file: set FIBER_SWIN_WEIGHTS=/path/to/"
                               "swin_base_patch4_window12_384_22k.pth (no network access to download them)")
        sd = torch.load(path, map_location="cpu")
        sd = sd.get("model", sd)
        sd = {k: v for k, v in sd.items() if not k.startswith("head.")}
        ckpt_res = 32 * ((int(round(sd["layers.0.blocks.0.attn.relative_position_bias_table"].shape[0] ** 0.5)) + 1) // 2)
        sd = swin_adapt_position_encoding(sd, before=ckpt_res, after=config["image_size"])
        missing, unexpected = model.load_state_dict(sd, strict=False)
        bad = [k for k in missing if "i2t" not in k and not k.endswith(("attn_mask", "relative_position_index"))]
        if bad:
            raise RuntimeError("FIBER_SWIN_WEIGHTS lacks backbone tensors: %s ..." % bad[:5])
    return model


def swin_base_patch4_window12_384_in22k(pretrained=False, **kwargs):
    return _create(pretrained, patch_size=4, window_size=12, embed_dim=128, depths=(2, 2, 18, 2),
                   num_heads=(4, 8, 16, 32), **kwargs)


def swin_base_patch4_window7_224_in22k(pretrained=False, **kwargs):
    return _create(pretrained, patch_size=4, window_size=7, embed_dim=128, depths=(2, 2, 18, 2),
                   num_heads=(4, 8, 16, 32), **kwargs)


swin_base_patch4_window12_384 = swin_base_patch4_window12_384_in22k
swin_base_patch4_window7_224 = swin_base_patch4_window7_224_in22k
