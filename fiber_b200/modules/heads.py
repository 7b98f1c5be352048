"""Task heads that sit on top of infer() — mirror of coarse_grained/fiber/modules/heads.py.
They are callers of the hot path, so they stay ordinary torch modules except the
poolers, whose dense layer is part of infer()'s tail and runs on the tcgen05 GEMM."""
import torch
import torch.nn as nn

from .. import ops
from .swin_transformer import FLinear


class Pooler(nn.Module):
    def __init__(self, hidden_size):
        super().__init__()
        self.dense = FLinear(hidden_size, hidden_size)
        self.dense.out_fp32 = True
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return self.activation(self.dense(hidden_states[:, 0]))


class ITMHead(nn.Module):
    def __init__(self, hidden_size):
        super().__init__()
        self.fc = nn.Linear(hidden_size, 2)

    def forward(self, x):
        return self.fc(x)


class BertPredictionHeadTransform(nn.Module):
    """transformers 4.6.0 BertPredictionHeadTransform: dense -> gelu -> LayerNorm(eps=config.layer_norm_eps)."""

    def __init__(self, hidden_size, layer_norm_eps=1e-12):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.transform_act_fn = nn.GELU()
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)

    def forward(self, x):
        return self.LayerNorm(self.transform_act_fn(self.dense(x)))


class MLMHead(nn.Module):
    def __init__(self, hidden_size, vocab_size, weight=None, layer_norm_eps=1e-12):
        super().__init__()
        self.transform = BertPredictionHeadTransform(hidden_size, layer_norm_eps)
        self.decoder = nn.Linear(hidden_size, vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(vocab_size))
        if weight is not None:
            self.decoder.weight = weight

    def forward(self, x):
        h = self.transform(x)
        if h.is_cuda:  # 768 -> vocab projection (3.1 GFLOP per pair fwd+bwd) on the tcgen05 GEMM
            return ops.VocabDecoderFn.apply(h, self.decoder.weight, self.bias)
        return self.decoder(h) + self.bias

    def loss_and_pred(self, x, labels):
        """cross_entropy(forward(x), labels, ignore_index=-100) and forward(x).argmax(-1) (0 on ignored rows) without the
        logits in memory (ops.MlmDecoderCEFn); what objectives.compute_mlm uses on the GPU."""
        h = self.transform(x)
        return ops.MlmDecoderCEFn.apply(h, self.decoder.weight, self.bias, labels)
