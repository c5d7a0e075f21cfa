#!/usr/bin/env python
"""BERT next-sentence-prediction fine-tuning (reference:
examples/BERT/ns_task_adaptdl.py): a BERT encoder with a 2-way head over
sentence pairs, ``autoscale_batch_size(4 * bs)``. Synthetic pairs."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _common  # noqa: E402

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from torch.utils.data import TensorDataset  # noqa: E402

import adaptdl_b200.torch as adl  # noqa: E402
from adaptdl_b200.models import BertModel, NextSentenceTask  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--emsize", type=int, default=768)
    parser.add_argument("--nhid", type=int, default=3072)
    parser.add_argument("--nlayers", type=int, default=12)
    parser.add_argument("--nhead", type=int, default=12)
    parser.add_argument("--lr", type=float, default=0.06)
    parser.add_argument("--epochs", type=int, default=1)
    parser.add_argument("--batch_size", type=int, default=24)
    parser.add_argument("--seq_len", type=int, default=128)
    parser.add_argument("--ntoken", type=int, default=28996)
    parser.add_argument("--pairs", type=int, default=4096)
    parser.add_argument("--checkpoint", default=None,
                        help="pre-trained BertModel state_dict")
    args = parser.parse_args()
    adl.init_process_group(_common.backend())
    device = _common.device()
    cuda = device.type == "cuda"
    gen = torch.Generator().manual_seed(0)
    seqs = torch.randint(2, args.ntoken, (args.pairs, args.seq_len),
                         generator=gen)
    types = torch.zeros_like(seqs)
    types[:, args.seq_len // 2:] = 1
    labels = torch.randint(0, 2, (args.pairs,), generator=gen)
    loader = adl.AdaptiveDataLoader(TensorDataset(seqs, types, labels),
                                    batch_size=args.batch_size, shuffle=True,
                                    drop_last=True)
    loader.autoscale_batch_size(4 * args.batch_size)
    bert = BertModel(args.ntoken, args.emsize, args.nhead, args.nhid,
                     args.nlayers, max_len=args.seq_len)
    if args.checkpoint:
        bert.load_state_dict(torch.load(args.checkpoint))
    model = NextSentenceTask(bert).to(device)
    criterion = nn.CrossEntropyLoss()
    optimizer = torch.optim.SGD(model.parameters(), lr=args.lr)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, 1, gamma=0.1)
    model = adl.AdaptiveDataParallel(model, optimizer, scheduler)
    for epoch in adl.remaining_epochs_until(args.epochs):
        stats = adl.Accumulator()
        for seq, tok_type, label in loader:
            seq, tok_type, label = (t.to(device) for t in
                                    (seq, tok_type, label))
            optimizer.zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=cuda):
                loss = criterion(model(seq, tok_type), label)
            loss.backward()
            optimizer.step()
            stats["loss_sum"] += loss.item() * label.size(0)
            stats["total"] += label.size(0)
        scheduler.step()
        with stats.synchronized():
            print("epoch {} loss {:.4f}".format(
                epoch, stats["loss_sum"] / stats["total"]))


if __name__ == "__main__":
    main()
