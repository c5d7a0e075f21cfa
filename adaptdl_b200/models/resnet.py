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

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from adaptdl_b200.ops.bn_act import BatchNormAct2d

__all__ = ["ResNet", "resnet18", "resnet34", "resnet50", "resnet101",
           "resnet152"]


def _shortcut(seq, x):
    """identity, or conv1x1 -> bn (no activation)"""
    if len(seq) == 0:
        return x
    return seq[1](seq[0](x), relu=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 3, stride, 1, bias=False)
        self.bn1 = BatchNormAct2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = BatchNormAct2d(planes)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = nn.Sequential(
                nn.Conv2d(in_planes, self.expansion * planes, 1, stride,
                          bias=False),
                BatchNormAct2d(self.expansion * planes))

    def forward(self, x):
        out = self.bn1(self.conv1(x))
        return self.bn2(self.conv2(out), residual=_shortcut(self.shortcut, x))


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 1, bias=False)
        self.bn1 = BatchNormAct2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = BatchNormAct2d(planes)
        self.conv3 = nn.Conv2d(planes, self.expansion * planes, 1,
                               bias=False)
        self.bn3 = BatchNormAct2d(self.expansion * planes)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = nn.Sequential(
                nn.Conv2d(in_planes, self.expansion * planes, 1, stride,
                          bias=False),
                BatchNormAct2d(self.expansion * planes))

    def forward(self, x):
        out = self.bn1(self.conv1(x))
        out = self.bn2(self.conv2(out))
        return self.bn3(self.conv3(out), residual=_shortcut(self.shortcut, x))


def padded_channels_conv2d(x, conv, multiple=8):
    """``conv(x)`` computed on input / filter zero-padded to a multiple of
    ``multiple`` input channels (identical result, tensor-core eligible)."""
    pad = (-x.shape[1]) % multiple
    w = conv.weight
    if pad:
        if x.is_contiguous(memory_format=torch.channels_last):
            # channels are the innermost dimension: pad there, keep the format
            x = F.pad(x.permute(0, 2, 3, 1), (0, pad)).permute(0, 3, 1, 2)
            w = F.pad(w.permute(0, 2, 3, 1), (0, pad)).permute(0, 3, 1, 2)
        else:
            x = F.pad(x, (0, 0, 0, 0, 0, pad))
            w = F.pad(w, (0, 0, 0, 0, 0, pad))
    return F.conv2d(x, w, conv.bias, conv.stride, conv.padding,
                    conv.dilation, conv.groups)


class ResNet(nn.Module):
    def __init__(self, block, num_blocks, num_classes=10):
        super().__init__()
        self.in_planes = 64
        self.conv1 = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
        self.bn1 = BatchNormAct2d(64)
        self.layer1 = self._make_layer(block, 64, num_blocks[0], 1)
        self.layer2 = self._make_layer(block, 128, num_blocks[1], 2)
        self.layer3 = self._make_layer(block, 256, num_blocks[2], 2)
        self.layer4 = self._make_layer(block, 512, num_blocks[3], 2)
        self.linear = nn.Linear(512 * block.expansion, num_classes)

    def _make_layer(self, block, planes, count, stride):
        layers = []
        for s in [stride] + [1] * (count - 1):
            layers.append(block(self.in_planes, planes, s))
            self.in_planes = planes * block.expansion
        return nn.Sequential(*layers)

    def _stem(self, x):
        # Experimental (ADAPTDL_B200_PAD_STEM=1, not timed on hardware yet):
        # cuDNN has no tensor-core kernels for 3 input channels; zero-padding
        # image and filter to 8 channels is the same convolution on the fast
        # path. The parameter keeps its [64, 3, 3, 3] shape.
        if x.is_cuda and os.environ.get("ADAPTDL_B200_PAD_STEM", "0") == "1":
            return padded_channels_conv2d(x, self.conv1)
        return self.conv1(x)

    def forward(self, x):
        out = self.bn1(self._stem(x))
        out = self.layer4(self.layer3(self.layer2(self.layer1(out))))
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
