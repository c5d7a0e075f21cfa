"""BERT encoder with MLM / next-sentence / QA heads -- the model of the
reference's ``examples/BERT`` workload (``model.py:99-175``: emsize 768,
nhid 3072, 12 layers, 12 heads, untied ``Linear(768, ntoken)`` MLM head,
learned positional + token-type embeddings, post-LayerNorm blocks).

B200-first choices: batch-first ``[N, S, E]`` activations, one fused QKV
projection (a single [E, 3E] GEMM instead of three), attention through
``F.scaled_dot_product_attention`` (flash kernels, bf16), GELU MLP. The
optional fused bias+GELU GEMM epilogue lives in ``adaptdl_b200.ops``.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F

from adaptdl_b200.ops.layer_norm import dropout_add_layer_norm
from adaptdl_b200.ops.linear_act import linear_act
from adaptdl_b200.ops.transformer import (linear, merge_heads, padded_logits,
                                          split_heads)

__all__ = ["BertModel", "MLMTask", "NextSentenceTask", "QuestionAnswerTask",
           "bert_base_mlm"]


class BertEmbedding(nn.Module):
    def __init__(self, ntoken, ninp, max_len=512, type_tokens=2,
                 dropout=0.1):
        super().__init__()
        self.embed = nn.Embedding(ntoken, ninp)
        self.pos_embed = nn.Embedding(max_len, ninp)
        self.tok_type_embed = nn.Embedding(type_tokens, ninp)
        self.norm = nn.LayerNorm(ninp)
        self.dropout = nn.Dropout(dropout)

    def forward(self, src, token_type_input=None):
        n, s = src.shape
        pos = torch.arange(s, device=src.device).unsqueeze(0)
        x = self.embed(src) + self.pos_embed(pos)
        if token_type_input is None:
            x = x + self.tok_type_embed.weight[0]
        else:
            x = x + self.tok_type_embed(token_type_input)
        return self.dropout(self.norm(x))


class EncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward, dropout=0.1,
                 activation="gelu"):
        super().__init__()
        assert d_model % nhead == 0
        self.nhead = nhead
        self.qkv = nn.Linear(d_model, 3 * d_model)
        self.out_proj = nn.Linear(d_model, d_model)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.drop = dropout
        self.activation = F.gelu if activation == "gelu" else F.relu

    def forward(self, x, attn_mask=None, is_causal=False):
        # fused QKV GEMM, then [N, S, 3*E] -> three contiguous [N, H, S, D]
        # (and the inverse pack in backward) with dedicated copy kernels
        qkv = linear(x, self.qkv.weight, self.qkv.bias)
        q, k, v = split_heads(qkv, self.nhead, 3)
        attn = F.scaled_dot_product_attention(
            q, k, v, attn_mask=attn_mask, is_causal=is_causal,
            dropout_p=self.drop if self.training else 0.0)
        attn = merge_heads(attn)                      # [N, S, E]
        # dropout + residual + LayerNorm: one fused kernel per direction
        x = dropout_add_layer_norm(
            x, linear(attn, self.out_proj.weight, self.out_proj.bias),
            self.norm1.weight, self.norm1.bias, self.drop, self.training,
            self.norm1.eps)
        if self.activation is F.gelu:
            # bias + GELU fused into the tcgen05 GEMM epilogue on B200; the
            # hidden dropout belongs to the same op so that its backward and
            # the GELU derivative are one elementwise pass
            h = linear_act(x, self.linear1.weight, self.linear1.bias, "gelu",
                           dropout_p=self.drop, training=self.training)
        else:
            h = F.dropout(self.activation(self.linear1(x)), self.drop,
                          self.training)
        h = linear(h, self.linear2.weight, self.linear2.bias)
        return dropout_add_layer_norm(x, h, self.norm2.weight,
                                      self.norm2.bias, self.drop,
                                      self.training, self.norm2.eps)


class BertModel(nn.Module):
    """Embeddings + a stack of encoder layers. Input ``[N, S]`` token ids."""

    def __init__(self, ntoken, ninp, nhead, nhid, nlayers, dropout=0.1,
                 max_len=512):
        super().__init__()
        self.ninp = ninp
        self.bert_embed = BertEmbedding(ntoken, ninp, max_len,
                                        dropout=dropout)
        self.layers = nn.ModuleList(
            EncoderLayer(ninp, nhead, nhid, dropout) for _ in range(nlayers))
        self.apply(self._init)

    @staticmethod
    def _init(m):
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=0.02)
        if isinstance(m, nn.Linear) and m.bias is not None:
            m.bias.data.zero_()

    def forward(self, src, token_type_input=None):
        x = self.bert_embed(src, token_type_input)
        for layer in self.layers:
            x = layer(x)
        return x


def aligned_linear(x, weight, bias, multiple=64):
    """``F.linear`` for an output width that is not a multiple of 8 (the
    28 996-token vocabulary of the MLM head): cuBLAS falls back to sm_80
    ``mma.sync`` kernels for such shapes (3x slower forward, dgrad and wgrad
    on B200). See :func:`adaptdl_b200.ops.transformer.padded_logits`."""
    return padded_logits(x, weight, bias, multiple)


class MLMTask(nn.Module):
    """Encoder + masked-language-model head (untied)."""

    def __init__(self, ntoken, ninp, nhead, nhid, nlayers, dropout=0.1,
                 max_len=512):
        super().__init__()
        self.bert_model = BertModel(ntoken, ninp, nhead, nhid, nlayers,
                                    dropout, max_len)
        self.mlm_span = nn.Linear(ninp, ninp)
        self.norm_layer = nn.LayerNorm(ninp, eps=1e-12)
        self.mlm_head = nn.Linear(ninp, ntoken)

    def forward(self, src, token_type_input=None):
        out = self.bert_model(src, token_type_input)
        out = self.norm_layer(F.gelu(linear(out, self.mlm_span.weight,
                                            self.mlm_span.bias)))
        return padded_logits(out, self.mlm_head.weight, self.mlm_head.bias)


class NextSentenceTask(nn.Module):
    def __init__(self, bert_model):
        super().__init__()
        self.bert_model = bert_model
        self.linear_layer = nn.Linear(bert_model.ninp, bert_model.ninp)
        self.ns_span = nn.Linear(bert_model.ninp, 2)

    def forward(self, src, token_type_input=None):
        out = self.bert_model(src, token_type_input)
        return self.ns_span(torch.tanh(self.linear_layer(out[:, 0])))


class QuestionAnswerTask(nn.Module):
    def __init__(self, bert_model):
        super().__init__()
        self.bert_model = bert_model
        self.qa_span = nn.Linear(bert_model.ninp, 2)

    def forward(self, src, token_type_input=None):
        out = F.gelu(self.bert_model(src, token_type_input))
        start, end = self.qa_span(out).split(1, dim=-1)
        return start.squeeze(-1), end.squeeze(-1)


def bert_base_mlm(ntoken=28996, max_len=512, dropout=0.1):
    """BERT-base (768 / 3072 / 12 layers / 12 heads) with the MLM head."""
    return MLMTask(ntoken, 768, 12, 3072, 12, dropout, max_len)
