#!/usr/bin/env python
"""BERT masked-language-model pre-training with AdaScale -- the reference's
examples/BERT/mlm_task_adaptdl.py workload (emsize 768, nhid 3072, 12 layers,
12 heads, sequences of 128 tokens, plain SGD lr 6 + StepLR, clip 0.1), in
bf16 on the B200-native engine (BASELINE config #3). A "sample" is one
sequence of ``--bptt`` tokens; 15 % of the tokens are masked (80 % [MASK],
10 % random, 10 % kept).

Synthetic token stream unless ``--data`` points to a saved 1-D LongTensor.

    torchrun --nproc-per-node 8 examples/BERT/mlm_task_adaptdl.py
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _common  # noqa: E402

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from torch.utils.data import Dataset  # noqa: E402

import adaptdl_b200.torch as adl  # noqa: E402
from adaptdl_b200 import env  # noqa: E402
from adaptdl_b200.models import MLMTask  # noqa: E402

MASK_ID, PAD_ID = 1, 0


class MaskedSequences(Dataset):
    def __init__(self, tokens, seq_len, ntoken, mask_frac=0.15):
        n = tokens.numel() // seq_len
        self.data = tokens[:n * seq_len].view(n, seq_len)
        self.ntoken, self.mask_frac = ntoken, mask_frac

    def __len__(self):
        return self.data.size(0)

    def __getitem__(self, idx):
        seq = self.data[idx]
        gen = torch.Generator().manual_seed(int(idx))
        roll = torch.rand(seq.shape, generator=gen)
        masked = roll < self.mask_frac
        inputs = seq.clone()
        kind = torch.rand(seq.shape, generator=gen)
        inputs[masked & (kind < 0.8)] = MASK_ID
        rand_pos = masked & (kind >= 0.8) & (kind < 0.9)
        inputs[rand_pos] = torch.randint(2, self.ntoken, (int(rand_pos.sum()),),
                                         generator=gen)
        targets = torch.where(masked, seq, torch.full_like(seq, -100))
        return inputs, targets


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--emsize", type=int, default=768)
    parser.add_argument("--nhid", type=int, default=3072)
    parser.add_argument("--nlayers", type=int, default=12)
    parser.add_argument("--nhead", type=int, default=12)
    parser.add_argument("--lr", type=float, default=6.0)
    parser.add_argument("--clip", type=float, default=0.1)
    parser.add_argument("--epochs", type=int, default=1)
    parser.add_argument("--batch_size", type=int, default=32,
                        help="sequences per global batch")
    parser.add_argument("--bptt", type=int, default=128)
    parser.add_argument("--dropout", type=float, default=0.1)
    parser.add_argument("--ntoken", type=int, default=28996)
    parser.add_argument("--tokens", type=int, default=4_000_000)
    parser.add_argument("--data", default=None)
    parser.add_argument("--autoscale-bsz", action="store_true")
    parser.add_argument("--gradient-accumulation", action="store_true")
    args = parser.parse_args()

    adl.init_process_group(_common.backend())
    device = _common.device()
    cuda = device.type == "cuda"
    if args.data:
        tokens = torch.load(args.data)
    else:
        gen = torch.Generator().manual_seed(0)
        tokens = torch.randint(2, args.ntoken, (args.tokens,), generator=gen)
    dataset = MaskedSequences(tokens, args.bptt, args.ntoken)
    loader = adl.AdaptiveDataLoader(dataset, batch_size=args.batch_size,
                                    shuffle=True, drop_last=True,
                                    num_workers=2, pin_memory=cuda)
    if args.autoscale_bsz:
        base = args.batch_size
        loader.autoscale_batch_size(
            128 * base, local_bsz_bounds=(max(base // 4, 1),
                                          min(2 * base, 64)),
            gradient_accumulation=args.gradient_accumulation)
    model = MLMTask(args.ntoken, args.emsize, args.nhead, args.nhid,
                    args.nlayers, args.dropout, max_len=args.bptt).to(device)
    criterion = nn.CrossEntropyLoss(ignore_index=-100)
    optimizer = torch.optim.SGD(model.parameters(), lr=args.lr)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, 1, gamma=0.1)
    model = adl.AdaptiveDataParallel(model, optimizer, scheduler)

    for epoch in adl.remaining_epochs_until(args.epochs):
        model.train()
        t0, seqs = time.time(), 0
        for i, (inputs, targets) in enumerate(loader):
            inputs = inputs.to(device, non_blocking=True)
            targets = targets.to(device, non_blocking=True)
            optimizer.zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=cuda):
                out = model(inputs)
                loss = criterion(out.view(-1, args.ntoken),
                                 targets.view(-1))
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), args.clip)
            optimizer.step()
            seqs += inputs.size(0)
            if i % 50 == 0 and env.replica_rank() == 0:
                print("epoch {} batch {} loss {:.3f} gain {:.3f} bsz {} "
                      "{:.1f} seq/s/replica".format(
                          epoch, i, loss.item(), model.gain,
                          loader.current_batch_size,
                          seqs / (time.time() - t0)))
        scheduler.step()


if __name__ == "__main__":
    main()
