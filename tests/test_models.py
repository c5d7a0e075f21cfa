"""Model zoo smoke tests (CPU): every CIFAR family builds, maps 3x32x32 to 10
logits and back-propagates; heavier variants are only instantiated."""

import pytest
import torch

from adaptdl_b200.models import cifar_zoo

LIGHT = ["LeNet", "VGG11", "ResNet18", "PreActResNet18", "GoogLeNet",
         "DenseNetCifar", "ResNeXt29_2x64d", "MobileNet", "MobileNetV2",
         "DPN26", "ShuffleNetG2", "ShuffleNetG3", "ShuffleNetV2_0.5",
         "SENet18", "PNASNetA", "PNASNetB"]


@pytest.mark.parametrize("name", LIGHT)
def test_forward_backward(name):
    torch.manual_seed(0)
    net = cifar_zoo.get_model(name)
    x = torch.randn(2, 3, 32, 32)
    out = net(x)
    assert out.shape == (2, 10)
    out.sum().backward()
    assert all(p.grad is not None for p in net.parameters())


def test_registry_covers_reference_families():
    families = ["VGG", "ResNet", "PreActResNet", "GoogLeNet", "DenseNet",
                "ResNeXt", "MobileNet", "MobileNetV2", "DPN", "ShuffleNetG",
                "ShuffleNetV2", "SENet", "PNASNet", "LeNet"]
    for fam in families:
        assert any(k.startswith(fam) for k in cifar_zoo.MODELS), fam
    with pytest.raises(ValueError):
        cifar_zoo.get_model("nope")


@pytest.mark.parametrize("name,params_m", [
    ("VGG19", 20.0), ("DenseNet121", 6.9), ("MobileNetV2", 2.3),
    ("ResNeXt29_32x4d", 4.7), ("PreActResNet50", 23.5)])
def test_parameter_counts(name, params_m):
    net = cifar_zoo.get_model(name)
    n = sum(p.numel() for p in net.parameters()) / 1e6
    assert abs(n - params_m) / params_m < 0.08, n


