"""The UNMODIFIED fine-grained fused backbone of the reference, importable in this image: FusionSwinTransformer of
fine_grained/maskrcnn_benchmark/modeling/backbone/fusion_swin_transformer_v2.py over the RobertaModel of
language_backbone/roberta_fused_model_v2.py, loaded BY PATH (the maskrcnn_benchmark package itself needs its compiled
`_C` extension) under the shims of baseline/ref_shims.py plus two HF 4.6 -> 5.x aliases.

Looked up in baseline/_ref/fine_grained/ first (untouched copies made by baseline/install_ref.py: git-ignored, travel to the
GPU box) and in /root/reference otherwise.  Users: tools/make_golden_fg.py (fixtures) and bench.py's `fg800` extra config
(the eager-GPU reference timed next to the CUDA path).  Nothing under fiber_b200/ imports this module."""
import importlib.util
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_FILES = ("backbone/fusion_swin_transformer_v2.py", "language_backbone/roberta_fused_model_v2.py")
_CANDIDATES = (os.path.join(_HERE, "_ref", "fine_grained"), "/root/reference/fine_grained/maskrcnn_benchmark/modeling")
NS = types.SimpleNamespace


def ref_dir():
    for c in _CANDIDATES:
        if all(os.path.isfile(os.path.join(c, f)) for f in _FILES):
            return c
    return None


def available():
    from . import ref_shims
    return ref_dir() is not None and ref_shims.available()


def load_reference():
    """(vision module, language module) of the reference, imported by path."""
    from . import ref_shims
    ref_shims.install()
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    if not hasattr(mu, "apply_chunking_to_forward"):
        mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    base = ref_dir()
    if base is None:
        raise RuntimeError("fine-grained reference files not found (baseline/_ref/fine_grained or /root/reference)")

    def load(name, rel):
        if name in sys.modules:
            return sys.modules[name]
        spec = importlib.util.spec_from_file_location(name, os.path.join(base, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m
    V = load("fg_swin_v2", _FILES[0])
    Lm = load("fg_roberta_v2", _FILES[1])

    def init_weights(self):
        self.apply(self._init_weights)

    def get_extended_attention_mask(self, attention_mask, input_shape=None, device=None):  # HF 4.6 semantics
        return (1.0 - attention_mask[:, None, None, :].to(dtype=torch.float32)) * -10000.0
    Lm.RobertaModel.init_weights = init_weights
    Lm.RobertaModel.get_extended_attention_mask = get_extended_attention_mask
    return V, Lm


def build(drop_path_rate=0.0):
    """FusionSwinTransformer(Swin-B window 12, roberta-base) of the reference with an identity in place of the FPN
    (out of scope: the forward returns the FPN inputs).  Returns (fusion module, swin, roberta)."""
    V, Lm = load_reference()
    from transformers.models.roberta.configuration_roberta import RobertaConfig
    cfg = RobertaConfig(vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                        hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                        max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=0,
                        eos_token_id=2)
    cfg.position_embedding_type = "absolute"
    cfg.chunk_size_feed_forward = 0
    cfg.is_decoder = False
    cfg.add_cross_attention = False
    rob = Lm.RobertaModel(cfg, add_pooling_layer=False)
    swin = V.SwinTransformer(patch_size=4, in_chans=3, embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32],
                             window_size=12, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0,
                             drop_path_rate=drop_path_rate, norm_layer=torch.nn.LayerNorm, ape=False, patch_norm=True,
                             frozen_stages=-1, backbone_arch="SWINT-FPN-RETINANET", use_checkpoint=False,
                             out_features=["stage2", "stage3", "stage4", "stage5"], max_query_len=256, lang_dim=768)

    class LangBody(torch.nn.Module):
        def __init__(self, model):
            super().__init__()
            self.model = model
            self.cfg = NS(MODEL=NS(DYHEAD=NS(FUSE_CONFIG=NS(USE_DOT_PRODUCT_TOKEN_LOSS=True)),
                                   LANGUAGE_BACKBONE=NS(LANG_DIM=768)))
        get_aggregated_output = Lm.RobertaFusedEncoder.get_aggregated_output

    class Wrap(torch.nn.Module):
        def __init__(self, body):
            super().__init__()
            self.body = body

        def fpn(self, outs):
            return outs
    return V.FusionSwinTransformer(Wrap(swin), Wrap(LangBody(rob))), swin, rob
