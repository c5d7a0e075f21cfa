#!/usr/bin/env python
"""MNIST porting tutorial, with TensorBoard output.

Step 5 plus ``to_tensorboard`` of the data loader and the model (batch size,
gradient noise statistics, gain) under ``$ADAPTDL_TENSORBOARD_LOGDIR``.

(Reference: tutorial/mnist_step_*.py and docs/adaptdl-pytorch.rst. Real MNIST
is used when torchvision finds it under ./data -- it is never downloaded --
otherwise a synthetic stand-in of the same shape.)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F
import torch.optim as optim
from torch.optim.lr_scheduler import StepLR

from adaptdl_b200.models import MnistNet as Net
import adaptdl_b200.torch as adl


def datasets():
    try:
        from torchvision import datasets as tvd, transforms
        tf = transforms.Compose([transforms.ToTensor(),
                                 transforms.Normalize((0.1307,), (0.3081,))])
        return (tvd.MNIST("./data", train=True, download=False, transform=tf),
                tvd.MNIST("./data", train=False, download=False,
                          transform=tf))
    except Exception:  # noqa: BLE001
        g = torch.Generator().manual_seed(0)
        mk = lambda n: torch.utils.data.TensorDataset(  # noqa: E731
            torch.randn(n, 1, 28, 28, generator=g),
            torch.randint(0, 10, (n,), generator=g))
        return mk(2048), mk(512)


def train(model, device, train_loader, optimizer, epoch):
    """One pass over the training loader (plain supervised step)."""
    model.train()
    for step, (images, labels) in enumerate(train_loader):
        images = images.to(device)
        labels = labels.to(device)
        optimizer.zero_grad()
        nll = F.nll_loss(model(images), labels)
        nll.backward()
        optimizer.step()
        if step % 10 == 0:
            print("epoch {:3d}  step {:4d}  train loss {:.4f}".format(
                epoch, step, nll.item()))

def test(model, device, test_loader, writer, epoch):
    model.eval()
    stats = adl.Accumulator()
    with torch.no_grad():
        for images, labels in test_loader:
            images, labels = images.to(device), labels.to(device)
            scores = model(images)
            stats["loss_sum"] += F.nll_loss(scores, labels,
                                            reduction="sum").item()
            stats["correct"] += scores.argmax(1).eq(labels).sum().item()
            stats["total"] += len(labels)
    with stats.synchronized():
        print("Test set: average loss {:.4f}, accuracy {}/{}".format(
            stats["loss_sum"] / stats["total"], stats["correct"],
            stats["total"]))
        writer.add_scalar("Loss/Test", stats["loss_sum"] / stats["total"], epoch)


def main():
    parser = argparse.ArgumentParser(description="MNIST tutorial")
    parser.add_argument("--batch-size", type=int, default=64)
    parser.add_argument("--epochs", type=int, default=14)
    parser.add_argument("--lr", type=float, default=1.0)
    parser.add_argument("--gamma", type=float, default=0.7)
    args = parser.parse_args()
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda" if use_cuda else "cpu")
    adl.init_process_group("nccl" if use_cuda else "gloo")
    train_set, test_set = datasets()
    train_loader = adl.AdaptiveDataLoader(train_set, batch_size=args.batch_size, shuffle=True, drop_last=True)
    test_loader = adl.AdaptiveDataLoader(test_set, batch_size=1000)
    train_loader.autoscale_batch_size(1028, local_bsz_bounds=(32, 128))
    model = Net().to(device)
    optimizer = optim.Adadelta(model.parameters(), lr=args.lr)
    scheduler = StepLR(optimizer, step_size=1, gamma=args.gamma)
    model = adl.AdaptiveDataParallel(model, optimizer, scheduler)
    from torch.utils.tensorboard import SummaryWriter
    writer = SummaryWriter(os.path.join(os.getenv("ADAPTDL_TENSORBOARD_LOGDIR", "/tmp"), "mnist"))
    for epoch in adl.remaining_epochs_until(args.epochs):
        train(model, device, train_loader, optimizer, epoch)
        test(model, device, test_loader, writer, epoch)
        scheduler.step()
        train_loader.to_tensorboard(writer, epoch, tag_prefix="AdaptDL/Data")
        model.to_tensorboard(writer, epoch, tag_prefix="AdaptDL/Model")


if __name__ == "__main__":
    main()
