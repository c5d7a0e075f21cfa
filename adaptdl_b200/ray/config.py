"""Defaults for Ray deployments (reference: ``adaptdl_ray/adaptdl/
config.py``)."""

_JOB_TIMEOUT = 24 * 3600
_CHECKPOINT_TIMEOUT = 120
_RESCHEDULE_TRIGGER_S = 100


def default_device(refresh=False):
    """``"GPU"`` if the Ray cluster has GPUs, else ``"CPU"``."""
    from adaptdl_b200.ray import require_ray
    ray = require_ray()
    gpus = sum(node["Resources"].get("GPU", 0) for node in ray.nodes()
               if node.get("Alive", True))
    return "GPU" if gpus > 0 else "CPU"


def nodes(resources_key=None):
    """``{node address: resources}`` of the live Ray nodes that offer the
    default device."""
    from adaptdl_b200.ray import require_ray
    ray = require_ray()
    key = resources_key or default_device()
    out = {}
    for node in ray.nodes():
        if node.get("Alive", True) and node["Resources"].get(key, 0) > 0:
            out[node["NodeManagerAddress"]] = dict(node["Resources"])
    return out
