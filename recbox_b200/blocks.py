"""f4: the remaining consumers of the [B, F, D] embedding tensor whose inner work is GEMM shaped, behind the reference's
module API, with every Linear / 1x1 convolution on the tcgen05 GEMM of csrc/gemm.cu (rbx_gemm_f32, 3xTF32 = fp32-level):

  CrossNet / CrossNetV2        ranking/pytorch/layers/interactions/cross_net.py:23-59
  CompressedInteractionNet     ranking/pytorch/layers/interactions/compressed_interaction_net.py:22-48
  ScaledDotProductAttention    ranking/pytorch/layers/attentions/dot_product_attention.py:21-43
  DIN_Attention                ranking/pytorch/layers/attentions/target_attention.py:25-66
  MultiHeadTargetAttention     ranking/pytorch/layers/attentions/target_attention.py:69-121

Same constructors, parameter names and init order as the reference (state_dicts load unchanged).  The element-wise glue
between the GEMMs (Hadamard products, softmax, masks) stays torch: plumbing, not the product."""
import torch
from torch import nn

from ._lib import RbxError
from .layers import MLP_Block, _MLPChainFn


def linear(x, weight, bias=None, relu=False):
    """F.linear (+ ReLU) on the fused GEMM, differentiable; x [..., K] fp32 CUDA."""
    if not x.is_cuda or x.dtype != torch.float32:
        raise RbxError("linear: input must be an fp32 CUDA tensor (recbox_b200 has no CPU path)")
    return _MLPChainFn.apply(x, (bool(relu),), weight, bias)


class CrossInteraction(nn.Module):
    def __init__(self, input_dim):
        super(CrossInteraction, self).__init__()
        self.weight = nn.Linear(input_dim, 1, bias=False)
        self.bias = nn.Parameter(torch.zeros(input_dim))

    def forward(self, X_0, X_i):
        return linear(X_i, self.weight.weight) * X_0 + self.bias


class CrossNet(nn.Module):
    def __init__(self, input_dim, num_layers):
        super(CrossNet, self).__init__()
        self.num_layers = num_layers
        self.cross_net = nn.ModuleList(CrossInteraction(input_dim) for _ in range(self.num_layers))

    def forward(self, X_0):
        X_i = X_0
        for i in range(self.num_layers):
            X_i = X_i + self.cross_net[i](X_0, X_i)
        return X_i


class CrossNetV2(nn.Module):
    """X_{i+1} = X_i + X_0 * (W_i X_i + b_i): one [B, d] x [d, d] GEMM per layer (d = F * D = 624 for the Criteo config)."""

    def __init__(self, input_dim, num_layers):
        super(CrossNetV2, self).__init__()
        self.num_layers = num_layers
        self.cross_layers = nn.ModuleList(nn.Linear(input_dim, input_dim) for _ in range(self.num_layers))

    def forward(self, X_0):
        X_i = X_0
        for i in range(self.num_layers):
            lin = self.cross_layers[i]
            X_i = torch.addcmul(X_i, X_0, linear(X_i, lin.weight, lin.bias))
        return X_i


class CompressedInteractionNet(nn.Module):
    """xDeepFM's CIN.  The reference forms the outer product [B, H*M, D] and runs a 1x1 Conv1d over the H*M channels; here the
    outer product is laid out [B, D, H*M] so that the convolution is ONE GEMM [B*D, H*M] x [H*M, out] on the tensor cores."""

    def __init__(self, num_fields, cin_hidden_units, output_dim=1):
        super(CompressedInteractionNet, self).__init__()
        self.cin_hidden_units = cin_hidden_units
        self.fc = nn.Linear(sum(cin_hidden_units), output_dim)
        self.cin_layer = nn.ModuleDict()
        for i, unit in enumerate(self.cin_hidden_units):
            in_channels = num_fields * self.cin_hidden_units[i - 1] if i > 0 else num_fields ** 2
            self.cin_layer["layer_" + str(i + 1)] = nn.Conv1d(in_channels, unit, kernel_size=1)

    def forward(self, feature_emb):
        pooling_outputs = []
        X_0 = feature_emb
        B, D = X_0.shape[0], X_0.shape[-1]
        X0t = X_0.transpose(1, 2)                                   # [B, D, H]
        Xit = X0t
        for i in range(len(self.cin_hidden_units)):
            conv = self.cin_layer["layer_" + str(i + 1)]
            had = (X0t.unsqueeze(-1) * Xit.unsqueeze(-2)).reshape(B * D, -1)     # [B*D, H*M], channel c = h * M + m as the reference
            out = linear(had, conv.weight.squeeze(-1), conv.bias)               # [B*D, out]
            Xit = out.view(B, D, -1)
            pooling_outputs.append(Xit.sum(dim=1))                              # sum over the embedding dim
        return linear(torch.cat(pooling_outputs, dim=-1), self.fc.weight, self.fc.bias)


class ScaledDotProductAttention(nn.Module):
    def __init__(self, dropout_rate=0.):
        super(ScaledDotProductAttention, self).__init__()
        self.dropout = nn.Dropout(dropout_rate) if dropout_rate > 0 else None

    def forward(self, Q, K, V, scale=None, mask=None):
        scores = torch.matmul(Q, K.transpose(-1, -2))        # [b, heads, 1, len] target attention: a batched row-dot, not a GEMM
        if scale:
            scores = scores / scale
        if mask is not None:
            mask = mask.view_as(scores)
            scores = scores.masked_fill_(mask.float() == 0, -1.e9)
        attention = scores.softmax(dim=-1)
        if self.dropout is not None:
            attention = self.dropout(attention)
        return torch.matmul(attention, V), attention


class DIN_Attention(nn.Module):
    """Target attention of DIN: the [B*L, 4D] -> units -> 1 scoring MLP runs on the fused GEMM chain."""

    def __init__(self, embedding_dim=64, attention_units=[32], hidden_activations="ReLU", output_activation=None, dropout_rate=0,
                 batch_norm=False, use_softmax=False):
        super(DIN_Attention, self).__init__()
        self.embedding_dim = embedding_dim
        self.use_softmax = use_softmax
        if isinstance(hidden_activations, str) and hidden_activations.lower() == "dice":
            raise RbxError("DIN_Attention: the Dice activation is not part of this package; pass module instances instead")
        self.attention_layer = MLP_Block(input_dim=4 * embedding_dim, output_dim=1, hidden_units=attention_units,
                                         hidden_activations=hidden_activations, output_activation=output_activation,
                                         dropout_rates=dropout_rate, batch_norm=batch_norm)

    def forward(self, target_item, history_sequence, mask=None):
        seq_len = history_sequence.size(1)
        target_item = target_item.unsqueeze(1).expand(-1, seq_len, -1)
        attention_input = torch.cat([target_item, history_sequence, target_item - history_sequence,
                                     target_item * history_sequence], dim=-1)
        attention_weight = self.attention_layer(attention_input.view(-1, 4 * self.embedding_dim))
        attention_weight = attention_weight.view(-1, seq_len)
        if mask is not None:
            attention_weight = attention_weight * mask.float()
        if self.use_softmax:
            if mask is not None:
                attention_weight = attention_weight + -1.e9 * (1 - mask.float())
            attention_weight = attention_weight.softmax(dim=-1)
        return (attention_weight.unsqueeze(-1) * history_sequence).sum(dim=1)


class MultiHeadTargetAttention(nn.Module):
    def __init__(self, input_dim=64, attention_dim=64, num_heads=1, dropout_rate=0, use_scale=True, use_qkvo=True):
        super(MultiHeadTargetAttention, self).__init__()
        if not use_qkvo:
            attention_dim = input_dim
        assert attention_dim % num_heads == 0, \
            "attention_dim={} is not divisible by num_heads={}".format(attention_dim, num_heads)
        self.num_heads = num_heads
        self.head_dim = attention_dim // num_heads
        self.scale = self.head_dim ** 0.5 if use_scale else None
        self.use_qkvo = use_qkvo
        if use_qkvo:
            self.W_q = nn.Linear(input_dim, attention_dim, bias=False)
            self.W_k = nn.Linear(input_dim, attention_dim, bias=False)
            self.W_v = nn.Linear(input_dim, attention_dim, bias=False)
            self.W_o = nn.Linear(attention_dim, input_dim, bias=False)
        self.dot_attention = ScaledDotProductAttention(dropout_rate)

    def forward(self, target_item, history_sequence, mask=None):
        if self.use_qkvo:
            query = linear(target_item, self.W_q.weight)
            key = linear(history_sequence, self.W_k.weight)
            value = linear(history_sequence, self.W_v.weight)
        else:
            query, key, value = target_item, history_sequence, history_sequence
        batch_size = query.size(0)
        query = query.view(batch_size, 1, self.num_heads, self.head_dim).transpose(1, 2)
        key = key.view(batch_size, -1, self.num_heads, self.head_dim).transpose(1, 2)
        value = value.view(batch_size, -1, self.num_heads, self.head_dim).transpose(1, 2)
        if mask is not None:
            mask = mask.view(batch_size, 1, 1, -1).expand(-1, self.num_heads, -1, -1)
        output, _ = self.dot_attention(query, key, value, scale=self.scale, mask=mask)
        output = output.transpose(1, 2).contiguous().view(-1, self.num_heads * self.head_dim)
        if self.use_qkvo:
            output = linear(output, self.W_o.weight)
        return output


__all__ = ["linear", "CrossInteraction", "CrossNet", "CrossNetV2", "CompressedInteractionNet", "ScaledDotProductAttention",
           "DIN_Attention", "MultiHeadTargetAttention"]
