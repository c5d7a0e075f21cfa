"""Shared helpers for the example scripts: repo on sys.path, device choice,
tensorboard writer that degrades to a no-op."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def device():
    if torch.cuda.is_available():
        from adaptdl_b200 import env
        dev = torch.device("cuda", env.local_rank()
                           % torch.cuda.device_count())
        torch.cuda.set_device(dev)
        return dev
    return torch.device("cpu")


def backend():
    return "nccl" if torch.cuda.is_available() else "gloo"


class _NullWriter(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def add_scalar(self, *args, **kwargs):
        pass


def summary_writer(subdir):
    """TensorBoard writer under $ADAPTDL_TENSORBOARD_LOGDIR on rank 0 (the
    CLI mounts it), no-op elsewhere or without tensorboard."""
    from adaptdl_b200 import env
    logdir = os.getenv("ADAPTDL_TENSORBOARD_LOGDIR")
    if env.replica_rank() != 0 or not logdir:
        return _NullWriter()
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(os.path.join(logdir, subdir))
    except Exception:  # noqa: BLE001
        return _NullWriter()
