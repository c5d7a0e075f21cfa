"""Model zoo: the workloads of the reference's examples, as plain
``torch.nn`` modules (forward/backward stay on PyTorch/cuDNN/cuBLAS; the
framework's own kernels are on the gradient path)."""

from .resnet import (resnet18, resnet34, resnet50, resnet101,  # noqa: F401
                     resnet152, ResNet)
