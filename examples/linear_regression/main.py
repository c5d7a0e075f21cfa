#!/usr/bin/env python
"""Tiny polynomial regression -- the plumbing workload (BASELINE config #1):
elastic data parallelism, adaptive batch size and checkpoint-restart on any
backend, in seconds. Same workload as the reference's
examples/linear_regression/main.py.

    python examples/linear_regression/main.py --epochs 5 --autoscale-bsz
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _common  # noqa: E402,F401

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch.optim.lr_scheduler import MultiStepLR  # noqa: E402
from torch.utils.data import TensorDataset  # noqa: E402

import adaptdl_b200.torch as adl  # noqa: E402
from adaptdl_b200.models import LinearRegression  # noqa: E402

POLY_DEGREE = 4


def main():
    parser = argparse.ArgumentParser(description="Linear Regression")
    parser.add_argument("--bs", default=128, type=int)
    parser.add_argument("--lr", default=0.1, type=float)
    parser.add_argument("--epochs", default=90, type=int)
    parser.add_argument("--autoscale-bsz", action="store_true")
    parser.add_argument("--size", default=10000, type=int)
    args = parser.parse_args()

    adl.init_process_group(_common.backend())
    device = _common.device()

    gen = torch.Generator().manual_seed(0)      # same data on every replica
    w_target = torch.randn(POLY_DEGREE, 1, generator=gen) * 5
    b_target = torch.randn(1, generator=gen) * 5
    x = torch.randn(args.size, generator=gen).unsqueeze(1)
    feats = torch.cat([x ** i for i in range(1, POLY_DEGREE + 1)], 1)
    y = feats.mm(w_target) + b_target + 0.25 * torch.randn(1, generator=gen)
    dataloader = adl.AdaptiveDataLoader(
        TensorDataset(feats, y), batch_size=args.bs, shuffle=True,
        num_workers=0, drop_last=True)
    if args.autoscale_bsz:
        dataloader.autoscale_batch_size(8 * args.bs,
                                        local_bsz_bounds=(32, 1024))

    net = LinearRegression(POLY_DEGREE, 1).to(device)
    optimizer = torch.optim.SGD(net.parameters(), lr=args.lr, momentum=0.9,
                                weight_decay=5e-4)
    lr_scheduler = MultiStepLR(optimizer, [30, 45], 0.1)
    net = adl.AdaptiveDataParallel(net, optimizer, lr_scheduler)

    for epoch in adl.remaining_epochs_until(args.epochs):
        stats = adl.Accumulator()
        for inputs, targets in dataloader:
            inputs, targets = inputs.to(device), targets.to(device)
            optimizer.zero_grad()
            loss = F.smooth_l1_loss(net(inputs), targets)
            loss.backward()
            optimizer.step()
            stats["loss_sum"] += loss.item() * targets.size(0)
            stats["total"] += targets.size(0)
        lr_scheduler.step()
        with stats.synchronized():
            print("epoch {} loss {:.4f} bsz {} gain {:.3f}".format(
                epoch, stats["loss_sum"] / stats["total"],
                dataloader._elastic.current_batch_size, net.gain))


if __name__ == "__main__":
    main()
