#!/usr/bin/env python
"""Neural Collaborative Filtering (NeuMF-end) with Adam => AdamScale -- the
reference's examples/NCF/main.py workload: a ~1.6 M parameter model whose
step is dominated by launch latency and small all-reduces (BASELINE config
#5). MovieLens-1M shaped synthetic interactions (6040 users x 3706 items,
4 sampled negatives per positive, re-sampled every epoch).

    python examples/NCF/main.py --epochs 2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _common  # noqa: E402

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from torch.utils.data import Dataset  # noqa: E402

import adaptdl_b200.torch as adl  # noqa: E402
from adaptdl_b200.models import NCF  # noqa: E402


class Interactions(Dataset):
    """Implicit-feedback pairs with on-the-fly negative sampling."""

    def __init__(self, users, items, positives, num_ng, seed=0):
        gen = torch.Generator().manual_seed(seed)
        self.users, self.items, self.num_ng = users, items, num_ng
        self.pos_u = torch.randint(0, users, (positives,), generator=gen)
        self.pos_i = torch.randint(0, items, (positives,), generator=gen)
        self.ng_sample(0)

    def ng_sample(self, epoch):
        gen = torch.Generator().manual_seed(1000 + epoch)
        n = len(self.pos_u)
        neg_i = torch.randint(0, self.items, (n * self.num_ng,),
                              generator=gen)
        self.u = torch.cat([self.pos_u, self.pos_u.repeat(self.num_ng)])
        self.i = torch.cat([self.pos_i, neg_i])
        self.label = torch.cat([torch.ones(n),
                                torch.zeros(n * self.num_ng)])

    def __len__(self):
        return len(self.u)

    def __getitem__(self, idx):
        return self.u[idx], self.i[idx], self.label[idx]


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--lr", type=float, default=0.001)
    parser.add_argument("--dropout", type=float, default=0.0)
    parser.add_argument("--batch_size", type=int, default=256)
    parser.add_argument("--epochs", type=int, default=20)
    parser.add_argument("--factor_num", type=int, default=32)
    parser.add_argument("--num_layers", type=int, default=3)
    parser.add_argument("--num_ng", type=int, default=4)
    parser.add_argument("--positives", type=int, default=100000)
    parser.add_argument("--autoscale-bsz", action="store_true")
    parser.add_argument("--gradient-accumulation", action="store_true")
    args = parser.parse_args()

    adl.init_process_group(_common.backend())
    device = _common.device()
    dataset = Interactions(6040, 3706, args.positives, args.num_ng)
    loader = adl.AdaptiveDataLoader(dataset, batch_size=args.batch_size,
                                    shuffle=True, num_workers=0,
                                    drop_last=True)
    if args.autoscale_bsz:
        loader.autoscale_batch_size(
            8192, local_bsz_bounds=(32, 512),
            gradient_accumulation=args.gradient_accumulation)
    model = NCF(6040, 3706, args.factor_num, args.num_layers, args.dropout,
                "NeuMF-end").to(device)
    loss_fn = nn.BCEWithLogitsLoss()
    optimizer = torch.optim.Adam(model.parameters(), lr=args.lr)
    model = adl.AdaptiveDataParallel(model, optimizer,
                                     find_unused_parameters=True)
    step = adl.GraphedTrainStep(
        model, optimizer, lambda n, u, i, y: loss_fn(n(u, i), y))
    for epoch in adl.remaining_epochs_until(args.epochs):
        model.train()
        dataset.ng_sample(epoch)
        stats = adl.Accumulator()
        for user, item, label in loader:
            loss = step(user, item, label.float())
            stats["loss_sum"] += loss.item() * label.size(0)
            stats["total"] += label.size(0)
        with stats.synchronized():
            print("epoch {} loss {:.4f} gain {:.3f}".format(
                epoch, stats["loss_sum"] / stats["total"], model.gain))


if __name__ == "__main__":
    main()
