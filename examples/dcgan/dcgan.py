#!/usr/bin/env python
"""DCGAN with TWO AdaptiveDataParallel instances (generator and
discriminator, each with its own optimizer) -- the reference's
examples/dcgan workload: gradient statistics of both instances are summed
into one goodput model (``_metrics.update_grad_params`` keyed per instance).
64x64 synthetic images.

    python examples/dcgan/dcgan.py --epochs 1 --images 1024
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _common  # noqa: E402

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from torch.utils.data import TensorDataset  # noqa: E402

import adaptdl_b200.torch as adl  # noqa: E402
from adaptdl_b200 import env  # noqa: E402
from adaptdl_b200.models import Discriminator, Generator  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--batch-size", type=int, default=64)
    parser.add_argument("--epochs", type=int, default=5)
    parser.add_argument("--lr", type=float, default=0.0002)
    parser.add_argument("--nz", type=int, default=100)
    parser.add_argument("--images", type=int, default=8192)
    parser.add_argument("--autoscale-bsz", action="store_true")
    args = parser.parse_args()
    adl.init_process_group(_common.backend())
    device = _common.device()
    gen = torch.Generator().manual_seed(0)
    data = torch.randn(args.images, 3, 64, 64, generator=gen).tanh_()
    loader = adl.AdaptiveDataLoader(TensorDataset(data),
                                    batch_size=args.batch_size, shuffle=True,
                                    drop_last=True)
    if args.autoscale_bsz:
        loader.autoscale_batch_size(16 * args.batch_size,
                                    local_bsz_bounds=(16, 512))
    netG, netD = Generator(args.nz).to(device), Discriminator().to(device)
    optD = torch.optim.SGD(netD.parameters(), lr=args.lr * 50,
                           momentum=0.5, nesterov=True)
    optG = torch.optim.SGD(netG.parameters(), lr=args.lr * 50,
                           momentum=0.5, nesterov=True)
    netD = adl.AdaptiveDataParallel(netD, optD, name="adaptdl-discriminator")
    netG = adl.AdaptiveDataParallel(netG, optG, name="adaptdl-generator")
    criterion = nn.BCEWithLogitsLoss()
    for epoch in adl.remaining_epochs_until(args.epochs):
        stats = adl.Accumulator()
        for (real,) in loader:
            real = real.to(device)
            n = real.size(0)
            ones = torch.ones(n, device=device)
            zeros = torch.zeros(n, device=device)
            # (1) discriminator: real + fake in ONE backward pass
            optD.zero_grad()
            fake = netG.module(torch.randn(n, args.nz, 1, 1, device=device))
            lossD = criterion(netD(torch.cat([real, fake.detach()])),
                              torch.cat([ones, zeros]))
            lossD.backward()
            optD.step()
            # (2) generator
            optG.zero_grad()
            fake = netG(torch.randn(n, args.nz, 1, 1, device=device))
            with netD.no_sync():
                lossG = criterion(netD.module(fake), ones)
            lossG.backward()
            optG.step()
            stats["lossD"] += lossD.item()
            stats["lossG"] += lossG.item()
            stats["n"] += 1
        with stats.synchronized():
            if env.replica_rank() == 0:
                print("epoch {} lossD {:.3f} lossG {:.3f} gainD {:.3f} "
                      "gainG {:.3f}".format(
                          epoch, stats["lossD"] / stats["n"],
                          stats["lossG"] / stats["n"], netD.gain, netG.gain))


if __name__ == "__main__":
    main()
