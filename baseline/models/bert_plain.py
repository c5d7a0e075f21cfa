"""BERT-base MLM in stock PyTorch modules, for the reference arm of
``bench.py --workload bert``.

The reference's own ``examples/BERT/model.py`` cannot run here: it imports
``torchtext.nn`` (not installable offline) and its custom encoder layer does
not accept the ``is_causal`` keyword that ``nn.TransformerEncoder`` passes to
its layers since torch 2.0. BASELINE.md section 2 therefore defines the
BERT baseline as "the same model under the reference's AdaptiveDataParallel
with bf16 autocast": this file is that model -- identical dimensions and
parameter count to the reference's ``MLMTask`` (768 / 3072 / 12 layers /
12 heads, learned position + token-type embeddings, post-LayerNorm blocks,
GELU MLP, untied ``Linear(768, ntoken)`` head), written with
``nn.TransformerEncoderLayer`` (whose attention runs through
``F.scaled_dot_product_attention``), i.e. what a reference user would write
on today's PyTorch. Nothing from ``adaptdl_b200`` is imported.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F


class BertEmbedding(nn.Module):
    def __init__(self, ntoken, ninp, max_len=512, dropout=0.1):
        super().__init__()
        self.embed = nn.Embedding(ntoken, ninp)
        self.pos_embed = nn.Embedding(max_len, ninp)
        self.tok_type_embed = nn.Embedding(2, ninp)
        self.norm = nn.LayerNorm(ninp)
        self.dropout = nn.Dropout(dropout)

    def forward(self, src, token_type_input=None):
        pos = torch.arange(src.shape[1], device=src.device).unsqueeze(0)
        if token_type_input is None:
            token_type_input = torch.zeros_like(src)
        x = self.embed(src) + self.pos_embed(pos) \
            + self.tok_type_embed(token_type_input)
        return self.dropout(self.norm(x))


class MLMTask(nn.Module):
    def __init__(self, ntoken, ninp=768, nhead=12, nhid=3072, nlayers=12,
                 dropout=0.1, max_len=512):
        super().__init__()
        self.bert_embed = BertEmbedding(ntoken, ninp, max_len, dropout)
        layer = nn.TransformerEncoderLayer(
            ninp, nhead, nhid, dropout, activation="gelu", batch_first=True,
            norm_first=False)
        self.encoder = nn.TransformerEncoder(layer, nlayers,
                                             enable_nested_tensor=False)
        self.mlm_span = nn.Linear(ninp, ninp)
        self.norm_layer = nn.LayerNorm(ninp, eps=1e-12)
        self.mlm_head = nn.Linear(ninp, ntoken)
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                m.weight.data.normal_(mean=0.0, std=0.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()

    def forward(self, src, token_type_input=None):
        x = self.encoder(self.bert_embed(src, token_type_input))
        x = self.norm_layer(F.gelu(self.mlm_span(x)))
        return self.mlm_head(x)


def bert_base_mlm(ntoken=28996, max_len=512, dropout=0.1):
    return MLMTask(ntoken, 768, 12, 3072, 12, dropout, max_len)
