"""Checkpoint transport through the Ray object store: a checkpoint
directory <-> ``{relative path: bytes}`` (reference: ``aws/utils.py``)."""

import enum
import os


class Status(enum.Enum):
    FAILED = 0
    SUCCEEDED = 1
    RUNNING = 2


def serialize_checkpoint(checkpoint_dir):
    data = {}
    for base, _, files in os.walk(checkpoint_dir):
        for name in files:
            path = os.path.join(base, name)
            with open(path, "rb") as f:
                data[os.path.relpath(path, checkpoint_dir)] = f.read()
    return data


def checkpoint_obj_to_dir(checkpoint_dir, checkpoint_obj):
    for rel, blob in checkpoint_obj.items():
        path = os.path.join(checkpoint_dir, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            f.write(blob)
