"""CIFAR-10 model zoo for the pytorch-cifar workload.

The reference example (``examples/pytorch-cifar/main.py:83-100`` with
``examples/pytorch-cifar/models/``) lets ``--model`` pick one of 15 CNN
families for 3x32x32 inputs. This module provides the same families
(VGG, ResNet, PreActResNet, GoogLeNet, DenseNet, ResNeXt, MobileNet,
MobileNetV2, DPN, ShuffleNet, ShuffleNetV2, SENet, PNASNet, LeNet) written
table-driven on a few shared building blocks; every model is channels-last /
bf16-autocast friendly (no in-place tricks that break CUDA-graph capture) and
is reachable by name through :func:`get_model`.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F

from adaptdl_b200.models import resnet as _resnet
from adaptdl_b200.ops.bn_act import BatchNormAct2d


class ConvBN(nn.Module):
    """conv (no bias) -> BatchNorm (-> ReLU); the normalisation and the
    activation are one fused kernel on B200 (``ops.BatchNormAct2d``)."""

    def __init__(self, cin, cout, k=3, stride=1, groups=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, padding=k // 2,
                              groups=groups, bias=False)
        self.bn = BatchNormAct2d(cout)
        self.act = act

    def forward(self, x, residual=None):
        return self.bn(self.conv(x), residual, self.act)


def conv_bn(cin, cout, k=3, stride=1, groups=1, act=True):
    return ConvBN(cin, cout, k, stride, groups, act)


class _Head(nn.Module):
    """global average pool -> flatten -> linear"""

    def __init__(self, cin, num_classes):
        super().__init__()
        self.fc = nn.Linear(cin, num_classes)

    def forward(self, x):
        return self.fc(torch.flatten(F.adaptive_avg_pool2d(x, 1), 1))


# ---------------------------------------------------------------------------
# LeNet / VGG
# ---------------------------------------------------------------------------
class LeNet(nn.Module):
    def __init__(self, num_classes=10):
        super().__init__()
        self.features = nn.Sequential(
            nn.Conv2d(3, 6, 5), nn.ReLU(), nn.MaxPool2d(2),
            nn.Conv2d(6, 16, 5), nn.ReLU(), nn.MaxPool2d(2))
        self.classifier = nn.Sequential(
            nn.Linear(16 * 5 * 5, 120), nn.ReLU(), nn.Linear(120, 84),
            nn.ReLU(), nn.Linear(84, num_classes))

    def forward(self, x):
        return self.classifier(torch.flatten(self.features(x), 1))


_VGG = {
    11: [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
    13: [64, 64, "M", 128, 128, "M", 256, 256, "M", 512, 512, "M",
         512, 512, "M"],
    16: [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M",
         512, 512, 512, "M"],
    19: [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M",
         512, 512, 512, 512, "M", 512, 512, 512, 512, "M"],
}


class VGG(nn.Module):
    def __init__(self, depth=19, num_classes=10):
        super().__init__()
        layers, cin = [], 3
        for item in _VGG[depth]:
            if item == "M":
                layers.append(nn.MaxPool2d(2))
            else:
                layers.append(conv_bn(cin, item))
                cin = item
        self.features = nn.Sequential(*layers)
        self.classifier = nn.Linear(512, num_classes)

    def forward(self, x):
        return self.classifier(torch.flatten(self.features(x), 1))


# ---------------------------------------------------------------------------
# pre-activation ResNet
# ---------------------------------------------------------------------------
class _PreActBlock(nn.Module):
    def __init__(self, cin, planes, stride, bottleneck):
        super().__init__()
        self.expansion = 4 if bottleneck else 1
        cout = planes * self.expansion
        self.bn1 = nn.BatchNorm2d(cin)
        if bottleneck:
            self.body = nn.Sequential(
                nn.Conv2d(cin, planes, 1, bias=False),
                nn.BatchNorm2d(planes), nn.ReLU(inplace=True),
                nn.Conv2d(planes, planes, 3, stride, 1, bias=False),
                nn.BatchNorm2d(planes), nn.ReLU(inplace=True),
                nn.Conv2d(planes, cout, 1, bias=False))
        else:
            self.body = nn.Sequential(
                nn.Conv2d(cin, planes, 3, stride, 1, bias=False),
                nn.BatchNorm2d(planes), nn.ReLU(inplace=True),
                nn.Conv2d(planes, cout, 3, 1, 1, bias=False))
        self.shortcut = None
        if stride != 1 or cin != cout:
            self.shortcut = nn.Conv2d(cin, cout, 1, stride, bias=False)

    def forward(self, x):
        pre = F.relu(self.bn1(x))
        skip = self.shortcut(pre) if self.shortcut is not None else x
        return self.body(pre) + skip


class PreActResNet(nn.Module):
    def __init__(self, blocks, bottleneck=False, num_classes=10):
        super().__init__()
        self.stem = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
        layers, cin = [], 64
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), blocks)):
            for j in range(n):
                blk = _PreActBlock(cin, planes, 2 if (j == 0 and i > 0) else 1,
                                   bottleneck)
                layers.append(blk)
                cin = planes * blk.expansion
        self.layers = nn.Sequential(*layers)
        self.head = _Head(cin, num_classes)

    def forward(self, x):
        return self.head(self.layers(self.stem(x)))


# ---------------------------------------------------------------------------
# ResNeXt-29
# ---------------------------------------------------------------------------
class _ResNeXtBlock(nn.Module):
    def __init__(self, cin, cardinality, width, stride):
        super().__init__()
        mid = cardinality * width
        cout = 2 * mid
        self.body = nn.Sequential(
            conv_bn(cin, mid, 1), conv_bn(mid, mid, 3, stride, cardinality),
            conv_bn(mid, cout, 1, act=False))
        self.shortcut = None
        if stride != 1 or cin != cout:
            self.shortcut = conv_bn(cin, cout, 1, stride, act=False)
        self.cout = cout

    def forward(self, x):
        skip = x if self.shortcut is None else self.shortcut(x)
        return F.relu(self.body(x) + skip)


class ResNeXt(nn.Module):
    def __init__(self, blocks=(3, 3, 3), cardinality=2, width=64,
                 num_classes=10):
        super().__init__()
        self.stem = conv_bn(3, 64, 1)
        layers, cin = [], 64
        for i, n in enumerate(blocks):
            for j in range(n):
                blk = _ResNeXtBlock(cin, cardinality, width,
                                    2 if (j == 0 and i > 0) else 1)
                layers.append(blk)
                cin = blk.cout
            width *= 2
        self.layers = nn.Sequential(*layers)
        self.head = _Head(cin, num_classes)

    def forward(self, x):
        return self.head(self.layers(self.stem(x)))


# ---------------------------------------------------------------------------
# DenseNet
# ---------------------------------------------------------------------------
class _DenseLayer(nn.Module):
    def __init__(self, cin, growth):
        super().__init__()
        self.fn = nn.Sequential(
            nn.BatchNorm2d(cin), nn.ReLU(inplace=True),
            nn.Conv2d(cin, 4 * growth, 1, bias=False),
            nn.BatchNorm2d(4 * growth), nn.ReLU(inplace=True),
            nn.Conv2d(4 * growth, growth, 3, padding=1, bias=False))

    def forward(self, x):
        return torch.cat([self.fn(x), x], 1)


class DenseNet(nn.Module):
    def __init__(self, blocks, growth=12, reduction=0.5, num_classes=10):
        super().__init__()
        c = 2 * growth
        layers = [nn.Conv2d(3, c, 3, padding=1, bias=False)]
        for i, n in enumerate(blocks):
            for _ in range(n):
                layers.append(_DenseLayer(c, growth))
                c += growth
            if i != len(blocks) - 1:
                cout = int(c * reduction)
                layers += [nn.BatchNorm2d(c), nn.ReLU(inplace=True),
                           nn.Conv2d(c, cout, 1, bias=False), nn.AvgPool2d(2)]
                c = cout
        layers += [nn.BatchNorm2d(c), nn.ReLU(inplace=True)]
        self.features = nn.Sequential(*layers)
        self.head = _Head(c, num_classes)

    def forward(self, x):
        return self.head(self.features(x))


# ---------------------------------------------------------------------------
# GoogLeNet
# ---------------------------------------------------------------------------
class _Inception(nn.Module):
    def __init__(self, cin, n1, n3r, n3, n5r, n5, pool):
        super().__init__()
        self.b1 = conv_bn(cin, n1, 1)
        self.b2 = nn.Sequential(conv_bn(cin, n3r, 1), conv_bn(n3r, n3, 3))
        self.b3 = nn.Sequential(conv_bn(cin, n5r, 1), conv_bn(n5r, n5, 3),
                                conv_bn(n5, n5, 3))
        self.b4 = nn.Sequential(nn.MaxPool2d(3, 1, 1), conv_bn(cin, pool, 1))

    def forward(self, x):
        return torch.cat([self.b1(x), self.b2(x), self.b3(x), self.b4(x)], 1)


class GoogLeNet(nn.Module):
    _CFG = [(192, 64, 96, 128, 16, 32, 32), (256, 128, 128, 192, 32, 96, 64),
            "M",
            (480, 192, 96, 208, 16, 48, 64), (512, 160, 112, 224, 24, 64, 64),
            (512, 128, 128, 256, 24, 64, 64), (512, 112, 144, 288, 32, 64, 64),
            (528, 256, 160, 320, 32, 128, 128), "M",
            (832, 256, 160, 320, 32, 128, 128),
            (832, 384, 192, 384, 48, 128, 128)]

    def __init__(self, num_classes=10):
        super().__init__()
        layers = [conv_bn(3, 192, 3)]
        for cfg in self._CFG:
            layers.append(nn.MaxPool2d(3, 2, 1) if cfg == "M"
                          else _Inception(*cfg))
        self.features = nn.Sequential(*layers)
        self.head = _Head(1024, num_classes)

    def forward(self, x):
        return self.head(self.features(x))


# ---------------------------------------------------------------------------
# MobileNet v1 / v2
# ---------------------------------------------------------------------------
class MobileNet(nn.Module):
    _CFG = [64, (128, 2), 128, (256, 2), 256, (512, 2), 512, 512, 512, 512,
            512, (1024, 2), 1024]

    def __init__(self, num_classes=10):
        super().__init__()
        layers, cin = [conv_bn(3, 32, 3)], 32
        for item in self._CFG:
            cout, stride = item if isinstance(item, tuple) else (item, 1)
            layers += [conv_bn(cin, cin, 3, stride, groups=cin),
                       conv_bn(cin, cout, 1)]
            cin = cout
        self.features = nn.Sequential(*layers)
        self.head = _Head(cin, num_classes)

    def forward(self, x):
        return self.head(self.features(x))


class _InvertedResidual(nn.Module):
    def __init__(self, cin, cout, expansion, stride):
        super().__init__()
        mid = cin * expansion
        self.body = nn.Sequential(
            conv_bn(cin, mid, 1), conv_bn(mid, mid, 3, stride, groups=mid),
            conv_bn(mid, cout, 1, act=False))
        self.residual = stride == 1
        self.shortcut = None
        if stride == 1 and cin != cout:
            self.shortcut = conv_bn(cin, cout, 1, act=False)

    def forward(self, x):
        out = self.body(x)
        if self.residual:
            out = out + (x if self.shortcut is None else self.shortcut(x))
        return out


class MobileNetV2(nn.Module):
    # (expansion, out, blocks, stride) -- strides adapted to 32x32 inputs
    _CFG = [(1, 16, 1, 1), (6, 24, 2, 1), (6, 32, 3, 2), (6, 64, 4, 2),
            (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]

    def __init__(self, num_classes=10):
        super().__init__()
        layers, cin = [conv_bn(3, 32, 3)], 32
        for expansion, cout, n, stride in self._CFG:
            for j in range(n):
                layers.append(_InvertedResidual(cin, cout, expansion,
                                                stride if j == 0 else 1))
                cin = cout
        layers.append(conv_bn(cin, 1280, 1))
        self.features = nn.Sequential(*layers)
        self.head = _Head(1280, num_classes)

    def forward(self, x):
        return self.head(self.features(x))


# ---------------------------------------------------------------------------
# SENet-18
# ---------------------------------------------------------------------------
class _SEBlock(nn.Module):
    def __init__(self, cin, planes, stride):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(cin)
        self.conv1 = nn.Conv2d(cin, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.shortcut = None
        if stride != 1 or cin != planes:
            self.shortcut = nn.Conv2d(cin, planes, 1, stride, bias=False)
        self.squeeze = nn.Conv2d(planes, planes // 16, 1)
        self.excite = nn.Conv2d(planes // 16, planes, 1)

    def forward(self, x):
        pre = F.relu(self.bn1(x))
        skip = x if self.shortcut is None else self.shortcut(pre)
        out = self.conv2(F.relu(self.bn2(self.conv1(pre))))
        gate = torch.sigmoid(self.excite(F.relu(self.squeeze(
            F.adaptive_avg_pool2d(out, 1)))))
        return out * gate + skip


class SENet(nn.Module):
    def __init__(self, blocks=(2, 2, 2, 2), num_classes=10):
        super().__init__()
        self.stem = conv_bn(3, 64, 3)
        layers, cin = [], 64
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), blocks)):
            for j in range(n):
                layers.append(_SEBlock(cin, planes,
                                       2 if (j == 0 and i > 0) else 1))
                cin = planes
        self.layers = nn.Sequential(*layers)
        self.head = _Head(cin, num_classes)

    def forward(self, x):
        return self.head(self.layers(self.stem(x)))


# ---------------------------------------------------------------------------
# ShuffleNet v1 / v2
# ---------------------------------------------------------------------------
def channel_shuffle(x, groups):
    n, c, h, w = x.shape
    return x.view(n, groups, c // groups, h, w).transpose(1, 2) \
        .reshape(n, c, h, w)


class _ShuffleUnit(nn.Module):
    def __init__(self, cin, cout, stride, groups):
        super().__init__()
        self.stride, self.groups = stride, groups
        mid = cout // 4
        g1 = 1 if cin == 24 else groups
        self.g1 = g1
        branch_out = cout - cin if stride == 2 else cout
        self.conv1 = conv_bn(cin, mid, 1, groups=g1)
        self.conv2 = conv_bn(mid, mid, 3, stride, groups=mid, act=False)
        self.conv3 = conv_bn(mid, branch_out, 1, groups=groups, act=False)

    def forward(self, x):
        out = channel_shuffle(self.conv1(x), self.g1)
        out = self.conv3(self.conv2(out))
        if self.stride == 2:
            return F.relu(torch.cat([out, F.avg_pool2d(x, 3, 2, 1)], 1))
        return F.relu(out + x)


class ShuffleNet(nn.Module):
    _PLANES = {2: (200, 400, 800), 3: (240, 480, 960)}

    def __init__(self, groups=2, blocks=(4, 8, 4), num_classes=10):
        super().__init__()
        layers, cin = [conv_bn(3, 24, 1)], 24
        for planes, n in zip(self._PLANES[groups], blocks):
            for j in range(n):
                layers.append(_ShuffleUnit(cin, planes, 2 if j == 0 else 1,
                                           groups))
                cin = planes
        self.features = nn.Sequential(*layers)
        self.head = _Head(cin, num_classes)

    def forward(self, x):
        return self.head(self.features(x))


class _ShuffleV2Unit(nn.Module):
    def __init__(self, cin, cout, down):
        super().__init__()
        self.down = down
        if down:
            mid = cout // 2
            self.left = nn.Sequential(
                conv_bn(cin, cin, 3, 2, groups=cin, act=False),
                conv_bn(cin, mid, 1))
            self.right = nn.Sequential(
                conv_bn(cin, mid, 1),
                conv_bn(mid, mid, 3, 2, groups=mid, act=False),
                conv_bn(mid, mid, 1))
        else:
            mid = cin // 2
            self.right = nn.Sequential(
                conv_bn(mid, mid, 1),
                conv_bn(mid, mid, 3, 1, groups=mid, act=False),
                conv_bn(mid, mid, 1))

    def forward(self, x):
        if self.down:
            out = torch.cat([self.left(x), self.right(x)], 1)
        else:
            keep, work = x.chunk(2, dim=1)
            out = torch.cat([keep, self.right(work)], 1)
        return channel_shuffle(out, 2)


class ShuffleNetV2(nn.Module):
    _PLANES = {0.5: (48, 96, 192, 1024), 1: (116, 232, 464, 1024),
               1.5: (176, 352, 704, 1024), 2: (224, 488, 976, 2048)}

    def __init__(self, net_size=1, blocks=(3, 7, 3), num_classes=10):
        super().__init__()
        planes = self._PLANES[net_size]
        layers, cin = [conv_bn(3, 24, 3)], 24
        for cout, n in zip(planes[:3], blocks):
            layers.append(_ShuffleV2Unit(cin, cout, True))
            layers += [_ShuffleV2Unit(cout, cout, False) for _ in range(n)]
            cin = cout
        layers.append(conv_bn(cin, planes[3], 1))
        self.features = nn.Sequential(*layers)
        self.head = _Head(planes[3], num_classes)

    def forward(self, x):
        return self.head(self.features(x))


# ---------------------------------------------------------------------------
# Dual path networks
# ---------------------------------------------------------------------------
class _DualPathBlock(nn.Module):
    def __init__(self, cin, mid, res, dense, stride, first):
        super().__init__()
        self.res = res
        self.body = nn.Sequential(
            conv_bn(cin, mid, 1), conv_bn(mid, mid, 3, stride, groups=32),
            conv_bn(mid, res + dense, 1, act=False))
        self.shortcut = conv_bn(cin, res + dense, 1, stride, act=False) \
            if first else None

    def forward(self, x):
        out = self.body(x)
        skip = x if self.shortcut is None else self.shortcut(x)
        r = self.res
        return F.relu(torch.cat([skip[:, :r] + out[:, :r], skip[:, r:],
                                 out[:, r:]], 1))


class DPN(nn.Module):
    def __init__(self, mids, res_planes, blocks, dense, num_classes=10):
        super().__init__()
        layers, cin = [conv_bn(3, 64, 3)], 64
        for i, (mid, res, n, d) in enumerate(zip(mids, res_planes, blocks,
                                                 dense)):
            for j in range(n):
                layers.append(_DualPathBlock(
                    cin, mid, res, d, 2 if (j == 0 and i > 0) else 1, j == 0))
                cin = res + (j + 2) * d
        self.features = nn.Sequential(*layers)
        self.head = _Head(cin, num_classes)

    def forward(self, x):
        return self.head(self.features(x))


# ---------------------------------------------------------------------------
# PNASNet
# ---------------------------------------------------------------------------
class _SepConv(nn.Module):
    def __init__(self, cin, cout, k, stride):
        super().__init__()
        self.fn = conv_bn(cin, cout, k, stride, groups=cin, act=False)

    def forward(self, x):
        return self.fn(x)


class _PNASCellA(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.stride = stride
        self.sep = _SepConv(cin, cout, 7, stride)
        self.proj = conv_bn(cin, cout, 1, act=False) if stride == 2 else None

    def forward(self, x):
        pooled = F.max_pool2d(x, 3, self.stride, 1)
        if self.proj is not None:
            pooled = self.proj(pooled)
        return F.relu(self.sep(x) + pooled)


class _PNASCellB(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.stride = stride
        self.sep7 = _SepConv(cin, cout, 7, stride)
        self.sep3 = _SepConv(cin, cout, 3, stride)
        self.sep5 = _SepConv(cin, cout, 5, stride)
        self.proj = conv_bn(cin, cout, 1, act=False) if stride == 2 else None
        self.merge = conv_bn(2 * cout, cout, 1)

    def forward(self, x):
        pooled = F.max_pool2d(x, 3, self.stride, 1)
        if self.proj is not None:
            pooled = self.proj(pooled)
        left = F.relu(self.sep7(x) + self.sep3(x))
        right = F.relu(pooled + self.sep5(x))
        return self.merge(torch.cat([left, right], 1))


class PNASNet(nn.Module):
    def __init__(self, cell, num_cells=6, planes=44, num_classes=10):
        super().__init__()
        layers, cin = [conv_bn(3, planes, 3)], planes
        for i in range(3):
            if i > 0:
                layers.append(cell(cin, cin * 2, 2))
                cin *= 2
            layers += [cell(cin, cin, 1) for _ in range(num_cells)]
        self.features = nn.Sequential(*layers)
        self.head = _Head(cin, num_classes)

    def forward(self, x):
        return self.head(self.features(x))


# ---------------------------------------------------------------------------
# registry
# ---------------------------------------------------------------------------
MODELS = {
    "LeNet": LeNet,
    "VGG11": lambda: VGG(11), "VGG13": lambda: VGG(13),
    "VGG16": lambda: VGG(16), "VGG19": lambda: VGG(19),
    "ResNet18": _resnet.resnet18, "ResNet34": _resnet.resnet34,
    "ResNet50": _resnet.resnet50, "ResNet101": _resnet.resnet101,
    "ResNet152": _resnet.resnet152,
    "PreActResNet18": lambda: PreActResNet((2, 2, 2, 2)),
    "PreActResNet34": lambda: PreActResNet((3, 4, 6, 3)),
    "PreActResNet50": lambda: PreActResNet((3, 4, 6, 3), True),
    "PreActResNet101": lambda: PreActResNet((3, 4, 23, 3), True),
    "PreActResNet152": lambda: PreActResNet((3, 8, 36, 3), True),
    "GoogLeNet": GoogLeNet,
    "DenseNet121": lambda: DenseNet((6, 12, 24, 16), 32),
    "DenseNet169": lambda: DenseNet((6, 12, 32, 32), 32),
    "DenseNet201": lambda: DenseNet((6, 12, 48, 32), 32),
    "DenseNet161": lambda: DenseNet((6, 12, 36, 24), 48),
    "DenseNetCifar": lambda: DenseNet((6, 12, 24, 16), 12),
    "ResNeXt29_2x64d": lambda: ResNeXt((3, 3, 3), 2, 64),
    "ResNeXt29_4x64d": lambda: ResNeXt((3, 3, 3), 4, 64),
    "ResNeXt29_8x64d": lambda: ResNeXt((3, 3, 3), 8, 64),
    "ResNeXt29_32x4d": lambda: ResNeXt((3, 3, 3), 32, 4),
    "MobileNet": MobileNet,
    "MobileNetV2": MobileNetV2,
    "DPN26": lambda: DPN((96, 192, 384, 768), (256, 512, 1024, 2048),
                         (2, 2, 2, 2), (16, 32, 24, 128)),
    "DPN92": lambda: DPN((96, 192, 384, 768), (256, 512, 1024, 2048),
                         (3, 4, 20, 3), (16, 32, 24, 128)),
    "ShuffleNetG2": lambda: ShuffleNet(2),
    "ShuffleNetG3": lambda: ShuffleNet(3),
    "ShuffleNetV2": lambda: ShuffleNetV2(1),
    "ShuffleNetV2_0.5": lambda: ShuffleNetV2(0.5),
    "ShuffleNetV2_1.5": lambda: ShuffleNetV2(1.5),
    "ShuffleNetV2_2": lambda: ShuffleNetV2(2),
    "SENet18": SENet,
    "PNASNetA": lambda: PNASNet(_PNASCellA, 6, 44),
    "PNASNetB": lambda: PNASNet(_PNASCellB, 6, 32),
}


def get_model(name):
    """Instantiate a CIFAR-10 model by name (see :data:`MODELS`)."""
    try:
        return MODELS[name]()
    except KeyError:
        raise ValueError("unknown model {!r}; choose from {}".format(
            name, ", ".join(sorted(MODELS)))) from None
