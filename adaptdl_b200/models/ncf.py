"""Neural Collaborative Filtering (He et al. 2017) -- the reference's
``examples/NCF`` workload (``model.py``): GMF and MLP towers over user/item
embeddings, fused by a final linear layer ("NeuMF-end"). ~1.6 M parameters on
MovieLens-1M shapes: the small-model, latency-bound end of the benchmark
suite."""

import torch
import torch.nn as nn

__all__ = ["NCF"]


class NCF(nn.Module):
    """Arguments: ``user_num``, ``item_num``, ``factor_num`` (predictive
    factors), ``num_layers`` (MLP depth), ``dropout``, ``model`` in
    {"MLP", "GMF", "NeuMF-end", "NeuMF-pre"}; ``GMF_model`` / ``MLP_model``
    are pre-trained towers for "NeuMF-pre"."""

    def __init__(self, user_num, item_num, factor_num=32, num_layers=3,
                 dropout=0.0, model="NeuMF-end", GMF_model=None,
                 MLP_model=None):
        super().__init__()
        self.model = model
        mlp_dim = factor_num * (2 ** (num_layers - 1))
        self.embed_user_GMF = nn.Embedding(user_num, factor_num)
        self.embed_item_GMF = nn.Embedding(item_num, factor_num)
        self.embed_user_MLP = nn.Embedding(user_num, mlp_dim)
        self.embed_item_MLP = nn.Embedding(item_num, mlp_dim)
        layers = []
        for i in range(num_layers):
            width = factor_num * (2 ** (num_layers - i))
            layers += [nn.Dropout(p=dropout), nn.Linear(width, width // 2),
                       nn.ReLU()]
        self.MLP_layers = nn.Sequential(*layers)
        predict = factor_num if model in ("MLP", "GMF") else 2 * factor_num
        self.predict_layer = nn.Linear(predict, 1)
        if model == "NeuMF-pre":
            self._load_pretrained(GMF_model, MLP_model)
        else:
            self._init_weights()

    def _init_weights(self):
        for emb in (self.embed_user_GMF, self.embed_user_MLP,
                    self.embed_item_GMF, self.embed_item_MLP):
            nn.init.normal_(emb.weight, std=0.01)
        for m in self.MLP_layers:
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
        nn.init.kaiming_uniform_(self.predict_layer.weight, a=1,
                                 nonlinearity="sigmoid")
        for m in self.modules():
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()

    def _load_pretrained(self, gmf, mlp):
        self.embed_user_GMF.weight.data.copy_(gmf.embed_user_GMF.weight)
        self.embed_item_GMF.weight.data.copy_(gmf.embed_item_GMF.weight)
        self.embed_user_MLP.weight.data.copy_(mlp.embed_user_MLP.weight)
        self.embed_item_MLP.weight.data.copy_(mlp.embed_item_MLP.weight)
        for m1, m2 in zip(self.MLP_layers, mlp.MLP_layers):
            if isinstance(m1, nn.Linear) and isinstance(m2, nn.Linear):
                m1.weight.data.copy_(m2.weight)
                m1.bias.data.copy_(m2.bias)
        weight = torch.cat([gmf.predict_layer.weight,
                            mlp.predict_layer.weight], dim=1)
        self.predict_layer.weight.data.copy_(0.5 * weight)
        self.predict_layer.bias.data.copy_(
            0.5 * (gmf.predict_layer.bias + mlp.predict_layer.bias))

    def forward(self, user, item):
        parts = []
        if self.model != "MLP":
            parts.append(self.embed_user_GMF(user) * self.embed_item_GMF(item))
        if self.model != "GMF":
            parts.append(self.MLP_layers(torch.cat(
                (self.embed_user_MLP(user), self.embed_item_MLP(item)), -1)))
        return self.predict_layer(torch.cat(parts, -1)).view(-1)
