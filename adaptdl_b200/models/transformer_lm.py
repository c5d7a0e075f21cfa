"""Causal Transformer language model -- the reference's
``examples/transformer`` workload (``transformer.py:59-120``: embedding *
sqrt(d), sinusoidal positions, ``nlayers`` encoder layers with a causal mask,
linear decoder; defaults emsize 200, nhid 200, 2 layers, 2 heads, WikiText-2
vocabulary). Input ``[S, N]`` like the reference (BPTT iterator layout)."""

import math

import torch
import torch.nn as nn

from .bert import EncoderLayer

__all__ = ["TransformerModel"]


class PositionalEncoding(nn.Module):
    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float()
                             * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)

    def forward(self, x):            # [N, S, E]
        return self.dropout(x + self.pe[:x.size(1)])


class TransformerModel(nn.Module):
    def __init__(self, ntoken, ninp=200, nhead=2, nhid=200, nlayers=2,
                 dropout=0.2):
        super().__init__()
        self.ninp = ninp
        self.encoder = nn.Embedding(ntoken, ninp)
        self.pos_encoder = PositionalEncoding(ninp, dropout)
        self.layers = nn.ModuleList(
            EncoderLayer(ninp, nhead, nhid, dropout, activation="relu")
            for _ in range(nlayers))
        self.decoder = nn.Linear(ninp, ntoken)
        self.init_weights()

    def init_weights(self):
        self.encoder.weight.data.uniform_(-0.1, 0.1)
        self.decoder.bias.data.zero_()
        self.decoder.weight.data.uniform_(-0.1, 0.1)

    def forward(self, src):          # [S, N] token ids -> [S, N, ntoken]
        x = self.encoder(src.t()) * math.sqrt(self.ninp)
        x = self.pos_encoder(x)
        for layer in self.layers:
            x = layer(x, is_causal=True)
        return self.decoder(x).transpose(0, 1)
