"""A stand-in for the slice of Ray's API that ``adaptdl_b200.ray`` touches,
for test runs on machines where Ray cannot be installed (this image: no
index access). Signatures and semantics follow Ray 2.x
(``ray.tune.experiment.Trial``, ``ray.tune.schedulers.TrialScheduler``,
``ray.tune.PlacementGroupFactory``, ``ray.nodes()``); only what the
integration calls is present. It lets the Tune scheduler / trial classes be
imported, subclass their Ray bases and be driven through their callbacks."""

__version__ = "2.9.0+fixture"

from . import _actors, autoscaler, util  # noqa: E402,F401
from ._actors import (cancel, get, get_actor, kill,  # noqa: E402,F401
                      put, remote, wait)

_NODES = [
    {"NodeManagerAddress": "10.0.0.1", "Alive": True,
     "Resources": {"CPU": 16.0, "GPU": 4.0, "node:10.0.0.1": 1.0}},
    {"NodeManagerAddress": "10.0.0.2", "Alive": True,
     "Resources": {"CPU": 16.0, "GPU": 4.0, "node:10.0.0.2": 1.0}},
]


def is_initialized():
    return True


def nodes():
    return [dict(n) for n in _NODES]


def init(*args, **kwargs):
    return None


def cluster_resources():
    total = {}
    for node in _NODES:
        for key, val in node["Resources"].items():
            total[key] = total.get(key, 0.0) + val
    return total


def available_resources():
    return cluster_resources()
