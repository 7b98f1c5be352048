"""Import shims that let the UNMODIFIED reference package (coarse_grained/fiber) be imported in this
image, where timm 0.4.12, pytorch_lightning 1.3.2 and sacred are not installed and transformers is
5.x instead of the pinned 4.6.0.

The package is looked up in baseline/_ref/ first (an untouched copy made by baseline/install_ref.py;
git-ignored, so it never enters history, but it travels to the GPU box with the snapshot) and in
/root/reference/coarse_grained otherwise (build container only).

Users: tools/make_golden.py (golden vectors), bench.py's reference arms (`--impl reference`,
`eager_gpu_baseline`), tests/test_reference_on_top.py (the reference's own objectives driven over
the fiber_b200 module).  Nothing under fiber_b200/ imports this module.  The shims restate the
third-party semantics the reference relies on: timm PatchEmbed/Mlp/DropPath/
trunc_normal_/_init_vit_weights, the LightningModule attributes objectives.py touches, HF 4.6
get_extended_attention_mask (-10000).
"""
import os
import sys
import types

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (os.path.join(_HERE, "_ref"), "/root/reference/coarse_grained")


def ref_root():
    """Directory holding the reference's `fiber` package, or None."""
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "fiber", "modules", "fiber_module.py")):
            return c
    return None


def available():
    return ref_root() is not None


_INSTALLED = None


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def install():
    global _INSTALLED
    if _INSTALLED is not None:
        return _INSTALLED
    REF_ROOT = ref_root()
    if REF_ROOT is None:
        raise RuntimeError("reference package not found (baseline/_ref/fiber or /root/reference/coarse_grained/fiber)")
    import transformers  # noqa: F401  (must be imported before a fake `timm` exists)
    import transformers.modeling_utils as mu
    import transformers.file_utils as fu
    import transformers.optimization as opt

    # ---- transformers 4.6 -> 5.x ---------------------------------------------------------
    def _noop_decorator(*a, **k):
        def deco(fn):
            return fn
        return deco

    for name in ("add_code_sample_docstrings", "add_start_docstrings",
                 "add_start_docstrings_to_model_forward", "replace_return_docstrings"):
        setattr(fu, name, _noop_decorator)
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    if not hasattr(mu, "prune_linear_layer"):
        mu.prune_linear_layer = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    if not hasattr(opt, "AdamW"):
        opt.AdamW = torch.optim.AdamW

    # ---- timm 0.4.12 ---------------------------------------------------------------------
    timm = _mod("timm")
    timm_data = _mod("timm.data")
    timm_models = _mod("timm.models")
    timm_helpers = _mod("timm.models.helpers")
    timm_layers = _mod("timm.models.layers")
    timm_registry = _mod("timm.models.registry")
    timm_vit = _mod("timm.models.vision_transformer")
    timm_features = _mod("timm.models.features")
    timm_hub = _mod("timm.models.hub")
    timm.data, timm.models = timm_data, timm_models
    timm_data.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
    timm_data.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

    class PatchEmbed(nn.Module):
        def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True):
            super().__init__()
            img_size, patch_size = to_2tuple(img_size), to_2tuple(patch_size)
            self.img_size, self.patch_size = img_size, patch_size
            self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
            self.num_patches = self.grid_size[0] * self.grid_size[1]
            self.flatten = flatten
            self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
            self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

        def forward(self, x):
            B, C, H, W = x.shape
            assert H == self.img_size[0] and W == self.img_size[1]
            x = self.proj(x)
            if self.flatten:
                x = x.flatten(2).transpose(1, 2)
            return self.norm(x)

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
            super().__init__()
            out_features = out_features or in_features
            hidden_features = hidden_features or in_features
            self.fc1 = nn.Linear(in_features, hidden_features)
            self.act = act_layer()
            self.fc2 = nn.Linear(hidden_features, out_features)
            self.drop = nn.Dropout(drop)

        def forward(self, x):
            return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))

    class DropPath(nn.Module):
        def __init__(self, drop_prob=None):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            shape = (x.shape[0],) + (1,) * (x.ndim - 1)
            rnd = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
            rnd.floor_()
            return x.div(keep) * rnd

    def _init_vit_weights(m, n="", head_bias=0.0, jax_impl=False):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.zeros_(m.bias)
            nn.init.ones_(m.weight)
        elif isinstance(m, nn.Conv2d):
            pass  # timm 0.4.12 leaves Conv2d at its default init in the non-jax path

    timm_layers.PatchEmbed, timm_layers.Mlp, timm_layers.DropPath = PatchEmbed, Mlp, DropPath
    timm_layers.to_2tuple, timm_layers.trunc_normal_ = to_2tuple, trunc_normal_
    timm_layers.Conv2dSame, timm_layers.Linear = nn.Conv2d, nn.Linear
    timm_registry.register_model = lambda fn: fn
    timm_vit.checkpoint_filter_fn = lambda sd, model: sd
    timm_vit._init_vit_weights = _init_vit_weights
    timm_helpers.build_model_with_cfg = None
    timm_helpers.overlay_external_default_cfg = lambda default_cfg, kwargs: None
    for n in ("FeatureListNet", "FeatureDictNet", "FeatureHookNet"):
        setattr(timm_features, n, type(n, (), {}))
    timm_hub.has_hf_hub = lambda *a, **k: False
    timm_hub.download_cached_file = timm_hub.load_state_dict_from_hf = timm_hub.load_state_dict_from_url = None

    # ---- pytorch_lightning 1.3.2 ---------------------------------------------------------
    pl = _mod("pytorch_lightning")
    pl_metrics = _mod("pytorch_lightning.metrics")
    pl.metrics = pl_metrics

    class _HParams(dict):
        __getattr__ = dict.__getitem__

    class LightningModule(nn.Module):
        def __init__(self):
            super().__init__()
            self.hparams = _HParams()
            self.trainer = None
            self.global_step = 0
            self.logged = {}

        def save_hyperparameters(self):
            import inspect
            frame = inspect.currentframe().f_back
            self.hparams["config"] = frame.f_locals["config"]

        @property
        def device(self):
            return next(self.parameters()).device

        def log(self, name, value, *a, **k):
            self.logged[name] = value

    class LightningDataModule:
        pass

    class Metric(nn.Module):
        def __init__(self, dist_sync_on_step=False):
            super().__init__()
            self._defaults = {}

        def add_state(self, name, default, dist_reduce_fx=None):
            self.register_buffer(name, default.clone())
            self._defaults[name] = default.clone()

        def forward(self, *a, **k):
            self.update(*a, **k)
            return self.compute()

        def reset(self):
            for k, v in self._defaults.items():
                getattr(self, k).copy_(v)

    pl.LightningModule, pl.LightningDataModule, pl_metrics.Metric = LightningModule, LightningDataModule, Metric

    # ---- sacred --------------------------------------------------------------------------
    sacred = _mod("sacred")

    class Experiment:
        def __init__(self, *a, **k):
            pass

        def config(self, fn):
            return fn

        def named_config(self, fn):
            return fn

        def automain(self, fn):
            return fn

    sacred.Experiment = Experiment

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    # ---- reference-level patches that restate HF 4.6 behaviour -----------------------------
    import fiber.modules.roberta as ref_roberta
    from transformers import RobertaConfig

    def init_weights(self):
        self.apply(self._init_weights)

    def get_extended_attention_mask(self, attention_mask, input_shape=None, device=None):
        # HF 4.6: (1 - mask[:, None, None, :]) * -10000.0 in the model dtype
        ext = attention_mask[:, None, None, :].to(dtype=torch.float32)
        return (1.0 - ext) * -10000.0

    @classmethod
    def from_pretrained(cls, name, *a, **k):
        cfg = RobertaConfig(vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                            intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
                            attention_probs_dropout_prob=0.1, max_position_embeddings=514, type_vocab_size=1,
                            layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=0, eos_token_id=2)
        cfg.position_embedding_type = "absolute"
        cfg.chunk_size_feed_forward = 0
        cfg.is_decoder = False
        cfg.add_cross_attention = False
        return cls(cfg)

    ref_roberta.RobertaModel.init_weights = init_weights
    ref_roberta.RobertaModel.get_extended_attention_mask = get_extended_attention_mask
    ref_roberta.RobertaModel.from_pretrained = from_pretrained
    # swin_build_model_with_cfg -> plain construction (pretrained=False: no network)
    import fiber.modules.swin_transformer as ref_swin

    def swin_build_model_with_cfg(model_cls, variant, pretrained, default_cfg=None, **kwargs):
        kwargs.pop("pretrained_filter_fn", None)
        kwargs.pop("pretrained_strict", None)
        kwargs.pop("num_classes", None)
        kwargs.pop("config", None)
        assert not pretrained, "no network: pretrained weights unavailable"
        return model_cls(**kwargs)

    ref_swin.swin_build_model_with_cfg = swin_build_model_with_cfg
    _INSTALLED = (ref_swin, ref_roberta)
    return _INSTALLED


def default_config(**over):
    """coarse_grained/fiber/config.py:21-92 defaults (sacred is shimmed, so restated here)."""
    loss_names = {"itm": 0, "mlm": 0, "itc": 0, "vqa": 0, "nlvr2": 0, "caption_mle": 0, "caption_gold": 0,
                  "caption_cider": 0}
    cfg = dict(
        exp_name="fiber", seed=0, loss_names=loss_names, batch_size=4096, image_size=384,
        vit="swin_base_patch4_window12_384_in22k", image_only=False, draw_false_image=0,
        input_image_embed_size=1024, resolution_before=384, pretrained_vit=False, vqav2_label_size=3129,
        max_text_len=40, tokenizer="roberta-base", vocab_size=50265, whole_word_masking=False, mlm_prob=0.15,
        draw_false_text=0, input_text_embed_size=768, hidden_size=768, num_heads=12, num_layers=12, mlp_ratio=4,
        drop_rate=0.1, num_fuse_block=6, itc_pooler=True, optim_type="adamw", learning_rate=1e-5,
        weight_decay=0.01, decay_power=1, max_epoch=100, max_steps=100000, warmup_steps=10000, end_lr=0,
        lr_mult_head=5, lr_mult_cross_modal=5, get_recall_metric=False, get_recall_metric_itc=True,
        cider_path=None, resume_from=None, fast_dev_run=False, val_check_interval=1.0, test_only=False,
        data_root="", log_dir="result", per_gpu_batchsize=0, num_gpus=1, num_nodes=1, load_path="",
        num_workers=8, precision=32,
    )
    tasks = over.pop("tasks", None)
    cfg.update(over)
    if tasks:
        cfg["loss_names"] = dict(loss_names, **{t: 1 for t in tasks})
    return cfg
