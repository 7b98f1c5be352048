"""RoBERTa-Base tower with FIBER's text->image cross attention — host-side mirror of
coarse_grained/fiber/modules/roberta.py (class / parameter names and forward signatures), computed
by fiber_b200.ops.  No dependency on `transformers`: the config is a plain object carrying the
roberta-base values the reference gets from `RobertaModel.from_pretrained("roberta-base")`.
"""
import torch
import torch.nn as nn

from .. import ops
from .swin_transformer import FLayerNorm, FLinear

NUM_FUSE_BLOCK = 6
DIM_IMG = 1024


class RobertaConfig:
    """roberta-base config.json values (transformers 4.6.0) used on the path."""

    def __init__(self, **kw):
        self.vocab_size = 50265
        self.hidden_size = 768
        self.num_hidden_layers = 12
        self.num_attention_heads = 12
        self.intermediate_size = 3072
        self.hidden_act = "gelu"
        self.hidden_dropout_prob = 0.1
        self.attention_probs_dropout_prob = 0.1
        self.max_position_embeddings = 514
        self.type_vocab_size = 1
        self.initializer_range = 0.02
        self.layer_norm_eps = 1e-5
        self.pad_token_id = 1
        self.position_embedding_type = "absolute"
        self.chunk_size_feed_forward = 0
        self.is_decoder = False
        self.add_cross_attention = False
        for k, v in kw.items():
            setattr(self, k, v)


class RobertaEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = FLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))
        self.position_embedding_type = "absolute"
        self.padding_idx = config.pad_token_id
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size,
                                                padding_idx=self.padding_idx)

    def forward(self, input_ids=None, token_type_ids=None, position_ids=None, inputs_embeds=None,
                past_key_values_length=0):
        if input_ids is None or token_type_ids is not None or position_ids is not None or inputs_embeds is not None:
            raise NotImplementedError("the B200 path implements embeddings(input_ids=...) as FIBER calls it "
                                      "(fiber_module.py:250,317)")
        p = self.dropout.p if self.training else 0.0
        return ops.RobertaEmbedFn.apply(input_ids, self.word_embeddings.weight, self.position_embeddings.weight,
                                        self.token_type_embeddings.weight, self.LayerNorm.weight, self.LayerNorm.bias,
                                        self.LayerNorm.eps, p, self.padding_idx)


class RobertaSelfAttention(nn.Module):
    def __init__(self, config, layer_index=None):
        super().__init__()
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = int(config.hidden_size / config.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = FLinear(config.hidden_size, self.all_head_size)
        if layer_index is None:
            kv_in = config.hidden_size
        else:  # roberta.py:235-241: stage-2 image width for layers 6-9, stage-3 width for 10-11
            kv_in = int(DIM_IMG / 2) if layer_index < 10 else DIM_IMG
        self.key = FLinear(kv_in, self.all_head_size)
        self.value = FLinear(kv_in, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)


class RobertaSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = FLinear(config.hidden_size, config.hidden_size)
        self.LayerNorm = FLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class RobertaAttention(nn.Module):
    def __init__(self, config, layer_index=None):
        super().__init__()
        self.self = RobertaSelfAttention(config, layer_index=layer_index)
        self.output = RobertaSelfOutput(config)
        self.pruned_heads = set()


class RobertaIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        assert config.hidden_act == "gelu"
        self.dense = FLinear(config.hidden_size, config.intermediate_size)


class RobertaOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = FLinear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = FLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class RobertaLayer(nn.Module):
    def __init__(self, config, layer_index=None):
        super().__init__()
        self.attention = RobertaAttention(config)
        if layer_index >= 12 - NUM_FUSE_BLOCK:
            self.crossattention_t2i = RobertaAttention(config, layer_index=layer_index)
        self.intermediate = RobertaIntermediate(config)
        self.output = RobertaOutput(config)
        self.alpha_t2i = nn.Parameter(torch.Tensor([0]))
        self.eps = config.layer_norm_eps

    def _params(self, fused):
        out = []
        for n in (ops.ROBERTA_FUSED if fused else ops.ROBERTA_PLAIN):
            obj = self
            for part in n.split("."):
                obj = getattr(obj, part)
            out.append(obj)
        return out

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_value=None, output_attentions=False, last_norm=True):
        if head_mask is not None or encoder_attention_mask is not None or past_key_value is not None \
                or output_attentions:
            raise NotImplementedError("head_mask / encoder_attention_mask / past_key_value / output_attentions "
                                      "are not used on the FIBER path")
        fused = encoder_hidden_states is not None
        if fused:
            assert hasattr(self, "crossattention_t2i"), \
                f"If `encoder_hidden_states` are passed, {self} has to be instantiated with cross-attention layers"
        B, L, C = hidden_states.shape
        mask2d = None
        if attention_mask is not None:
            if attention_mask.numel() != B * L:
                raise NotImplementedError("only (B,1,1,L) additive key masks are supported (no causal decoder mask)")
            mask2d = attention_mask.reshape(B, L).float()
        p_h = self.output.dropout.p if self.training else 0.0
        p_a = self.attention.self.dropout.p if self.training else 0.0
        meta = (self.attention.self.num_attention_heads, bool(last_norm), self.eps, p_h, p_a)
        out = ops.RobertaLayerFn.apply(hidden_states, mask2d, encoder_hidden_states if fused else None, meta,
                                       *self._params(fused))
        return (out,)


class RobertaEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.layer = nn.ModuleList([RobertaLayer(config, layer_index=i) for i in range(config.num_hidden_layers)])

    def forward(self, hidden_states, attention_mask=None, **kw):
        for layer in self.layer:
            hidden_states = layer(hidden_states, attention_mask)[0]
        return (hidden_states,)


class RobertaPooler(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = FLinear(config.hidden_size, config.hidden_size)
        self.dense.out_fp32 = True
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return self.activation(self.dense(hidden_states[:, 0]))


class RobertaModel(nn.Module):
    def __init__(self, config, add_pooling_layer=True):
        super().__init__()
        self.config = config
        self.embeddings = RobertaEmbeddings(config)
        self.encoder = RobertaEncoder(config)
        self.pooler = RobertaPooler(config) if add_pooling_layer else None
        self.apply(self._init_weights)

    def _init_weights(self, module):
        """RobertaPreTrainedModel._init_weights (roberta.py:631-644)."""
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    @classmethod
    def from_pretrained(cls, name, *a, **k):
        """HF `from_pretrained("roberta-base")` (fiber_module.py:88) without network access: the roberta-base weights
        are read from a local file — FIBER_ROBERTA_WEIGHTS=/path/to/pytorch_model.bin (an HF RobertaModel /
        RobertaForMaskedLM state_dict; `name` may also be a directory holding pytorch_model.bin).  Without one the
        model keeps its random initialisation and says so LOUDLY: that is only right when a FIBER checkpoint is
        loaded afterwards (config["load_path"]) or for synthetic benchmarks."""
        import os
        import warnings
        if os.path.basename(str(name).rstrip("/")) not in ("roberta-base",) and not os.path.isdir(str(name)):
            raise ValueError("FIBER-Base pairs Swin-B with roberta-base; got tokenizer/model name %r" % (name,))
        model = cls(RobertaConfig())
        path = os.environ.get("FIBER_ROBERTA_WEIGHTS", "")
        if not path and os.path.isdir(str(name)) and os.path.exists(os.path.join(str(name), "pytorch_model.bin")):
            path = os.path.join(str(name), "pytorch_model.bin")
        if path and os.path.exists(path):
            sd = torch.load(path, map_location="cpu")
            sd = {(k[len("roberta."):] if k.startswith("roberta.") else k): v for k, v in sd.items()}
            sd = {k: v for k, v in sd.items() if not k.startswith("lm_head.") and not k.endswith("position_ids")}
            missing, unexpected = model.load_state_dict(sd, strict=False)
            bad = [k for k in missing if "t2i" not in k and not k.endswith("position_ids")]
            if bad:
                raise RuntimeError("FIBER_ROBERTA_WEIGHTS lacks encoder tensors: %s ..." % bad[:5])
        else:
            warnings.warn("fiber_b200: RobertaModel.from_pretrained(%r) found no local weights (FIBER_ROBERTA_WEIGHTS is "
                          "unset / missing and there is no network): the text tower is RANDOMLY INITIALISED. Load a FIBER "
                          "checkpoint (load_path) or point FIBER_ROBERTA_WEIGHTS at roberta-base's pytorch_model.bin before "
                          "training from scratch." % (name,), RuntimeWarning, stacklevel=2)
        return model

    def get_extended_attention_mask(self, attention_mask, input_shape=None, device=None):
        """transformers 4.6.0 semantics: (1 - mask[:, None, None, :]) * -10000.0"""
        if attention_mask.dim() != 2:
            raise ValueError("expected a (batch, seq_len) attention mask")
        return (1.0 - attention_mask[:, None, None, :].to(dtype=torch.float32)) * -10000.0

    def forward(self, input_ids=None, attention_mask=None, **kw):
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        ext = self.get_extended_attention_mask(attention_mask, input_ids.shape, input_ids.device)
        h = self.encoder(self.embeddings(input_ids=input_ids), ext)[0]
        pooled = self.pooler(h) if self.pooler is not None else None
        return (h, pooled)
