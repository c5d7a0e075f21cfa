"""Test harness only (put on PYTHONPATH by tests/test_api_surface.py): there
is no network here, so ``torchvision.datasets.CIFAR10`` / ``MNIST`` become
small synthetic datasets of the same shapes (PIL images, int labels, the
caller's transforms applied the way torchvision applies them). This lets the
reference's example scripts run byte for byte; only the download is faked."""
import os

_SIZE = os.environ.get("ADL_TEST_FAKE_DATASETS")
if _SIZE:
    import torchvision

    def _fake(shape):
        class Fake(torchvision.datasets.FakeData):
            def __init__(self, root=None, train=True, download=False,
                         transform=None, **kwargs):
                super().__init__(size=int(_SIZE), image_size=shape,
                                 num_classes=10, transform=transform,
                                 random_offset=0 if train else 10 ** 6)

            def __getitem__(self, index):
                image, target = super().__getitem__(index)
                if shape[0] == 1:
                    pass            # FakeData already returns mode "L"
                return image, int(target)
        return Fake
    torchvision.datasets.CIFAR10 = _fake((3, 32, 32))
    torchvision.datasets.MNIST = _fake((1, 28, 28))

if os.environ.get("ADL_TEST_CUDA_IS_CPU"):
    # examples/NCF calls .cuda() unconditionally; this container has no GPU
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
