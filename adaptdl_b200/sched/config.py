"""Scheduler configuration.

Every setting is an environment variable rendered into the scheduler pods by
the helm chart's ConfigMap (``deploy/helm/.../templates/config.yaml``); the
``get_*`` accessors below are generated from one table so that the variable
name, its type and its default live in a single place (same accessor names as
the reference's ``sched/adaptdl_sched/config.py``).
"""

import json
import os

# custom resource coordinates
GROUP, VERSION, PLURAL = "adaptdl.petuum.com", "v1", "adaptdljobs"
# label of the cluster expander's placeholder pods
ADAPTDL_PH_LABEL = "adaptdl/placeholder"

_NODEGROUP_TAINT = ("petuum.com/nodegroup", "adaptdl")
_NAMESPACE_FILE = "/var/run/secrets/kubernetes.io/serviceaccount/namespace"
_REQUIRED = object()

# accessor -> (environment variable, parser, default)
_SETTINGS = {
    "get_image": ("ADAPTDL_IMAGE", str, _REQUIRED),
    "get_adaptdl_deployment": ("ADAPTDL_SCHED_DEPLOYMENT", str, _REQUIRED),
    "get_storage_subpath": ("ADAPTDL_STORAGE_SUBPATH", str, _REQUIRED),
    "get_supervisor_url": ("ADAPTDL_SUPERVISOR_URL", str, ""),
    "get_supervisor_port": ("ADAPTDL_SUPERVISOR_SERVICE_PORT", int, 8080),
    "get_adaptdl_version": ("ADAPTDL_SCHED_VERSION", str, "0.0.0"),
    "get_job_default_resources": ("ADAPTDL_JOB_DEFAULT_RESOURCES",
                                  json.loads, None),
    "get_job_patch_pods": ("ADAPTDL_JOB_PATCH_PODS", json.loads, None),
    "get_job_patch_containers": ("ADAPTDL_JOB_PATCH_CONTAINERS", json.loads,
                                 None),
}


def _accessor(name, variable, parse, default):
    def get():
        raw = os.environ.get(variable)
        if raw is None or raw == "":
            if default is _REQUIRED:
                raise KeyError(variable)
            return default
        return parse(raw)
    get.__name__ = name
    get.__doc__ = "``{}``{}".format(
        variable, "" if default is _REQUIRED
        else " (default: {!r})".format(default))
    return get


for _name, (_variable, _parse, _default) in _SETTINGS.items():
    globals()[_name] = _accessor(_name, _variable, _parse, _default)


def get_namespace():
    """Namespace of the pod this code runs in (``default`` outside a
    cluster)."""
    try:
        with open(_NAMESPACE_FILE) as f:
            return f.read().strip()
    except OSError:
        return "default"


def allowed_taints(taints):
    """A node is usable when it carries no taint, or exactly the AdaptDL
    node-group taint (dedicated nodes). ``taints`` holds dicts or client
    objects with ``key`` / ``value``."""
    if not taints:
        return True
    if len(taints) > 1:
        return False
    taint = taints[0]
    pair = (taint["key"], taint["value"]) if isinstance(taint, dict) \
        else (taint.key, taint.value)
    return pair == _NODEGROUP_TAINT
