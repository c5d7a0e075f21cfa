#!/usr/bin/env python
"""Transformer language model on a WikiText-2 shaped token stream with the
elastic BPTT iterator -- the reference's examples/transformer/transformer.py
workload (emsize 200, nhid 200, 2 layers, 2 heads, bptt 35, SGD lr 5.0,
StepLR 0.95, grad-clip 0.5) and BASELINE config #4 (elastic rescale
2->4->8->4 mid-epoch: run it under ``python -m adaptdl_b200.sched.local``).

The token stream is synthetic (Zipf-distributed ids over a 33 278-word
vocabulary, 2.09 M tokens = WikiText-2 train size) unless ``--data`` points
to a 1-D ``torch.save``d LongTensor.

    python -m adaptdl_b200.sched.local --gpus 8 --schedule 2,4,8,4 \
        --interval 20 examples/transformer/transformer.py --epochs 3
"""
import argparse
import faulthandler
import signal
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _common  # noqa: E402

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

import adaptdl_b200.torch as adl  # noqa: E402
from adaptdl_b200 import env  # noqa: E402
from adaptdl_b200.models import TransformerModel  # noqa: E402
from adaptdl_b200.torch.iterator import AdaptiveBPTTIterator  # noqa: E402

NTOKENS = 33278


def token_stream(n, seed):
    gen = torch.Generator().manual_seed(seed)
    # Zipf-like: exponentiate a uniform variate
    u = torch.rand(n, generator=gen)
    return (NTOKENS ** u).long().clamp_(0, NTOKENS - 1)


faulthandler.register(signal.SIGUSR1)   # kill -USR1 <pid>: thread dump


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--bs", type=int, default=20)
    parser.add_argument("--lr", type=float, default=5.0)
    parser.add_argument("--epochs", type=int, default=3)
    parser.add_argument("--bptt", type=int, default=35)
    parser.add_argument("--emsize", type=int, default=200)
    parser.add_argument("--nhid", type=int, default=200)
    parser.add_argument("--nlayers", type=int, default=2)
    parser.add_argument("--nhead", type=int, default=2)
    parser.add_argument("--dropout", type=float, default=0.2)
    parser.add_argument("--tokens", type=int, default=2088628)
    parser.add_argument("--data", default=None)
    parser.add_argument("--autoscale-bsz", action="store_true")
    args = parser.parse_args()

    adl.init_process_group(_common.backend())
    device = _common.device()
    train = torch.load(args.data) if args.data else \
        token_stream(args.tokens, 0)
    valid = token_stream(max(args.tokens // 10, 10 * args.bptt), 1)
    kwargs = {}
    if args.autoscale_bsz:
        kwargs = dict(max_batch_size=1024 * args.bs,
                      local_bsz_bounds=(16, 256))
    train_iter = AdaptiveBPTTIterator(train, batch_size=args.bs,
                                      bptt_len=args.bptt, device=device,
                                      **kwargs)
    valid_iter = AdaptiveBPTTIterator(valid, batch_size=10,
                                      bptt_len=args.bptt, device=device)
    model = TransformerModel(NTOKENS, args.emsize, args.nhead, args.nhid,
                             args.nlayers, args.dropout).to(device)
    criterion = nn.CrossEntropyLoss()
    optimizer = torch.optim.SGD(model.parameters(), lr=args.lr)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, 1, gamma=0.95)
    model = adl.AdaptiveDataParallel(model, optimizer, scheduler)
    cuda = device.type == "cuda"

    for epoch in adl.remaining_epochs_until(args.epochs):
        model.train()
        stats = adl.Accumulator()
        t0, tokens = time.time(), 0
        for i, batch in enumerate(train_iter):
            optimizer.zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=cuda):
                out = model(batch.text)
                loss = criterion(out.reshape(-1, NTOKENS),
                                 batch.target.reshape(-1))
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
            optimizer.step()
            tokens += batch.text.numel()
            stats["loss_sum"] += loss.item()
            stats["batches"] += 1
            if i % 200 == 0 and env.replica_rank() == 0:
                print("epoch {} batch {} replicas {} local_bsz {} loss "
                      "{:.3f} gain {:.3f} {:.0f} tok/s/replica".format(
                          epoch, i, env.num_replicas(),
                          train_iter.current_local_bsz, loss.item(),
                          model.gain, tokens / (time.time() - t0)))
        model.eval()
        vstats = adl.Accumulator()
        with torch.no_grad():
            for batch in valid_iter:
                with torch.autocast("cuda", dtype=torch.bfloat16,
                                    enabled=cuda):
                    out = model(batch.text)
                    vloss = criterion(out.reshape(-1, NTOKENS),
                                      batch.target.reshape(-1))
                vstats["loss_sum"] += vloss.item()
                vstats["batches"] += 1
        scheduler.step()
        with stats.synchronized(), vstats.synchronized():
            tl = stats["loss_sum"] / max(stats["batches"], 1)
            vl = vstats["loss_sum"] / max(vstats["batches"], 1)
            print("| end of epoch {} | train loss {:.3f} | valid loss "
                  "{:.3f} | valid ppl {:.1f}".format(
                      epoch, tl, vl, math.exp(min(vl, 20))))


if __name__ == "__main__":
    main()
