"""CIFAR-style ResNets (3x3 stem, 32x32 inputs) -- the model family of the
reference's ``examples/pytorch-cifar`` workload (ResNet18 =
BasicBlock x [2, 2, 2, 2], 11.17 M parameters, 62 parameter tensors).

Written for channels-last bf16 autocast on B200: convolutions have no bias
(BatchNorm follows), the classifier is a single Linear. Every
``bn -> (+ shortcut) -> relu`` group is ONE fused op
(:class:`adaptdl_b200.ops.BatchNormAct2d`, ``csrc/adl_bn.cu``); parameter
names and shapes are those of the plain ``nn.BatchNorm2d`` model, so state
dicts are interchangeable.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F

from adaptdl_b200.ops.bn_act import BatchNormAct2d

__all__ = ["ResNet", "resnet18", "resnet34", "resnet50", "resnet101",
           "resnet152"]


class ResidualBlock(nn.Module):
    """One residual unit, plain (two 3x3 convolutions) or bottleneck (1x1,
    3x3, 1x1 with a 4x wider output). Sub-module names follow the usual
    CIFAR ResNet checkpoints: ``conv{i}`` / ``bn{i}`` for the main path and
    ``shortcut.0`` / ``shortcut.1`` for the projection (empty when the
    input can be added as is).

    Every ``bn -> (+ shortcut) -> relu`` group is one fused op
    (:class:`BatchNormAct2d`): the last normalisation of the main path takes
    the shortcut as its residual input."""

    def __init__(self, in_planes, planes, stride, bottleneck):
        super().__init__()
        self.expansion = 4 if bottleneck else 1
        out_planes = planes * self.expansion
        # (kernel, in, out, stride) of the main path
        if bottleneck:
            plan = [(1, in_planes, planes, 1), (3, planes, planes, stride),
                    (1, planes, out_planes, 1)]
        else:
            plan = [(3, in_planes, planes, stride), (3, planes, planes, 1)]
        self.depth = len(plan)
        for i, (k, cin, cout, s) in enumerate(plan, start=1):
            setattr(self, "conv{}".format(i),
                    nn.Conv2d(cin, cout, k, s, k // 2, bias=False))
            setattr(self, "bn{}".format(i), BatchNormAct2d(cout))
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != out_planes:
            self.shortcut = nn.Sequential(
                nn.Conv2d(in_planes, out_planes, 1, stride, bias=False),
                BatchNormAct2d(out_planes))

    def forward(self, x):
        skip = x
        if len(self.shortcut):                   # 1x1 projection, no ReLU
            skip = self.shortcut[1](self.shortcut[0](x), relu=False)
        out = x
        for i in range(1, self.depth + 1):
            conv = getattr(self, "conv{}".format(i))
            norm = getattr(self, "bn{}".format(i))
            last = i == self.depth
            out = norm(conv(out), residual=skip if last else None)
        return out


def BasicBlock(in_planes, planes, stride=1):
    return ResidualBlock(in_planes, planes, stride, bottleneck=False)


def Bottleneck(in_planes, planes, stride=1):
    return ResidualBlock(in_planes, planes, stride, bottleneck=True)


BasicBlock.expansion = 1
Bottleneck.expansion = 4




class ResNet(nn.Module):
    """Stem (3x3, 64 channels) -> four stages of residual units at widths
    64 / 128 / 256 / 512 (the first unit of stages 2-4 halves the
    resolution) -> global average pool -> linear classifier."""

    WIDTHS = (64, 128, 256, 512)

    def __init__(self, block, num_blocks, num_classes=10):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
        self.bn1 = BatchNormAct2d(64)
        channels = 64
        for stage, (width, count) in enumerate(zip(self.WIDTHS, num_blocks),
                                               start=1):
            units = []
            for unit in range(count):
                stride = 2 if (unit == 0 and stage > 1) else 1
                units.append(block(channels, width, stride))
                channels = width * block.expansion
            setattr(self, "layer{}".format(stage), nn.Sequential(*units))
        self.linear = nn.Linear(channels, num_classes)

    def forward(self, x):
        out = self.bn1(self.conv1(x))
        for stage in (self.layer1, self.layer2, self.layer3, self.layer4):
            out = stage(out)
        out = F.adaptive_avg_pool2d(out, 1)
        return self.linear(torch.flatten(out, 1))


def resnet18(num_classes=10):
    return ResNet(BasicBlock, [2, 2, 2, 2], num_classes)


def resnet34(num_classes=10):
    return ResNet(BasicBlock, [3, 4, 6, 3], num_classes)


def resnet50(num_classes=10):
    return ResNet(Bottleneck, [3, 4, 6, 3], num_classes)


def resnet101(num_classes=10):
    return ResNet(Bottleneck, [3, 4, 23, 3], num_classes)


def resnet152(num_classes=10):
    return ResNet(Bottleneck, [3, 8, 36, 3], num_classes)
