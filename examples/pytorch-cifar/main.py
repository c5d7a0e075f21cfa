#!/usr/bin/env python
"""CIFAR-10 image classification with adaptive batch size -- the
reference's examples/pytorch-cifar/main.py workload (ResNet-18 default, SGD
with one param group per tensor => per-tensor AdaScale, MultiStepLR,
autoscale up to 4096 with local batch 32..1024), on the B200-native engine:
bf16 autocast + channels-last, whole-step CUDA graph
(``adl.GraphedTrainStep``), fused all-reduce + statistics, fused SGD.

Data: real CIFAR-10 if torchvision + the dataset are available under
``$ADAPTDL_SHARE_PATH`` (never downloaded), otherwise ``--synthetic``.

    python examples/pytorch-cifar/main.py --synthetic --epochs 2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _common  # noqa: E402

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from torch.optim.lr_scheduler import MultiStepLR  # noqa: E402
from torch.utils.data import TensorDataset  # noqa: E402

import adaptdl_b200.torch as adl  # noqa: E402
from adaptdl_b200 import env, models  # noqa: E402

MODELS = models.cifar_zoo.MODELS


def load_data(args):
    if not args.synthetic:
        try:
            import torchvision
            import torchvision.transforms as T
            norm = T.Normalize((0.4914, 0.4822, 0.4465),
                               (0.2023, 0.1994, 0.2010))
            train = torchvision.datasets.CIFAR10(
                root=env.share_path() or "./data", train=True,
                download=False, transform=T.Compose([
                    T.RandomCrop(32, padding=4), T.RandomHorizontalFlip(),
                    T.ToTensor(), norm]))
            valid = torchvision.datasets.CIFAR10(
                root=env.share_path() or "./data", train=False,
                download=False, transform=T.Compose([T.ToTensor(), norm]))
            return train, valid
        except Exception as exc:  # noqa: BLE001
            print("CIFAR-10 unavailable ({}); using synthetic data".format(
                exc))
    gen = torch.Generator().manual_seed(0)
    train = TensorDataset(torch.randn(args.synthetic_size, 3, 32, 32,
                                      generator=gen),
                          torch.randint(0, 10, (args.synthetic_size,),
                                        generator=gen))
    valid = TensorDataset(torch.randn(2000, 3, 32, 32, generator=gen),
                          torch.randint(0, 10, (2000,), generator=gen))
    return train, valid


def main():
    parser = argparse.ArgumentParser(description="CIFAR10 training")
    parser.add_argument("--bs", default=128, type=int)
    parser.add_argument("--lr", default=0.1, type=float)
    parser.add_argument("--epochs", default=60, type=int)
    parser.add_argument("--model", default="ResNet18", choices=sorted(MODELS))
    parser.add_argument("--autoscale-bsz", action="store_true")
    parser.add_argument("--mixed-precision", action="store_true",
                        help="fp16 autocast + GradScaler (reference flag); "
                             "default is bf16 autocast on CUDA")
    parser.add_argument("--bf16-params", action="store_true",
                        help="keep conv/linear weights in bf16; the fused "
                             "optimizer holds fp32 masters (no cast kernels)")
    parser.add_argument("--no-graph", action="store_true")
    parser.add_argument("--synthetic", action="store_true")
    parser.add_argument("--synthetic-size", default=12800, type=int)
    args = parser.parse_args()

    adl.init_process_group(_common.backend())
    device = _common.device()
    cuda = device.type == "cuda"
    trainset, validset = load_data(args)
    trainloader = adl.AdaptiveDataLoader(trainset, batch_size=args.bs,
                                         shuffle=True, num_workers=2,
                                         drop_last=True, pin_memory=cuda)
    if args.autoscale_bsz:
        trainloader.autoscale_batch_size(4096, local_bsz_bounds=(32, 1024),
                                         gradient_accumulation=False)
    validloader = adl.AdaptiveDataLoader(validset, batch_size=100,
                                         shuffle=False, num_workers=2)

    net = MODELS[args.model]().to(device)
    if cuda:
        net = net.to(memory_format=torch.channels_last)
        torch.backends.cudnn.benchmark = True
        if args.bf16_params and not args.mixed_precision:
            adl.mixed_precision_params(net)     # before the optimizer exists
    criterion = nn.CrossEntropyLoss()
    optimizer = torch.optim.SGD([{"params": [p]} for p in net.parameters()],
                                lr=args.lr, momentum=0.9, weight_decay=5e-4)
    lr_scheduler = MultiStepLR(optimizer, [30, 45], 0.1)
    scaler = torch.amp.GradScaler("cuda", enabled=True) \
        if (args.mixed_precision and cuda) else None
    net = adl.AdaptiveDataParallel(net, optimizer, lr_scheduler, scaler)
    step = adl.GraphedTrainStep(
        net, optimizer, lambda n, x, y: criterion(n(x), y),
        autocast_dtype=torch.bfloat16 if cuda else None,
        enabled=not args.no_graph and scaler is None, channels_last=cuda)

    def train(epoch, writer):
        net.train()
        stats = adl.Accumulator()
        for inputs, targets in trainloader:
            if scaler is not None:
                inputs, targets = inputs.to(device), targets.to(device)
                optimizer.zero_grad()
                with torch.autocast("cuda", dtype=torch.float16):
                    loss = criterion(net(inputs), targets)
                scaler.scale(loss).backward()
                scaler.step(optimizer)
                scaler.update()
            else:
                loss = step(inputs, targets)
            stats["loss_sum"] += loss.item() * targets.size(0)
            stats["total"] += targets.size(0)
        trainloader.to_tensorboard(writer, epoch, "AdaptDL/Data/")
        net.to_tensorboard(writer, epoch, "AdaptDL/Model/")
        with stats.synchronized():
            stats["loss_avg"] = stats["loss_sum"] / stats["total"]
            writer.add_scalar("Loss/Train", stats["loss_avg"], epoch)
            print("Train:", dict(stats))

    def valid(epoch, writer):
        net.eval()
        stats = adl.Accumulator()
        with torch.no_grad():
            for inputs, targets in validloader:
                inputs, targets = inputs.to(device), targets.to(device)
                if cuda:
                    inputs = inputs.contiguous(
                        memory_format=torch.channels_last)
                with torch.autocast("cuda", dtype=torch.bfloat16,
                                    enabled=cuda):
                    outputs = net(inputs)
                    loss = criterion(outputs, targets)
                stats["loss_sum"] += loss.item() * targets.size(0)
                stats["total"] += targets.size(0)
                stats["correct"] += outputs.argmax(1).eq(targets).sum().item()
        with stats.synchronized():
            stats["loss_avg"] = stats["loss_sum"] / stats["total"]
            stats["accuracy"] = stats["correct"] / stats["total"]
            writer.add_scalar("Loss/Valid", stats["loss_avg"], epoch)
            writer.add_scalar("Accuracy/Valid", stats["accuracy"], epoch)
            print("Valid:", dict(stats))

    with _common.summary_writer(env.job_id() or "cifar") as writer:
        for epoch in adl.remaining_epochs_until(args.epochs):
            train(epoch, writer)
            valid(epoch, writer)
            lr_scheduler.step()


if __name__ == "__main__":
    main()
