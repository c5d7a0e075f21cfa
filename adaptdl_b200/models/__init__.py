"""Model zoo: the workloads of the reference's examples, as plain
``torch.nn`` modules (forward/backward stay on PyTorch/cuDNN/cuBLAS; the
framework's own kernels are on the gradient path)."""

from .resnet import (resnet18, resnet34, resnet50, resnet101,  # noqa: F401
                     resnet152, ResNet)
from .bert import (BertModel, MLMTask, NextSentenceTask,  # noqa: F401
                   QuestionAnswerTask, bert_base_mlm)
from .transformer_lm import TransformerModel  # noqa: F401
from .ncf import NCF  # noqa: F401
from .dcgan import Generator, Discriminator  # noqa: F401
from .simple import LinearRegression, MnistNet  # noqa: F401
from . import cifar_zoo  # noqa: F401,E402
from .cifar_zoo import get_model  # noqa: F401,E402
