"""Scheduler configuration, from environment variables rendered by the helm
chart's ConfigMap (parity: reference ``sched/adaptdl_sched/config.py``)."""

import json
import os

ADAPTDL_PH_LABEL = "adaptdl/placeholder"
GROUP, VERSION, PLURAL = "adaptdl.petuum.com", "v1", "adaptdljobs"
_NS_FILE = "/var/run/secrets/kubernetes.io/serviceaccount/namespace"


def allowed_taints(taints):
    """Nodes are usable if untainted, or tainted only with the adaptdl
    node-group taint."""
    if not taints:
        return True
    if len(taints) != 1:
        return False
    taint = taints[0]
    key = taint["key"] if isinstance(taint, dict) else taint.key
    value = taint["value"] if isinstance(taint, dict) else taint.value
    return key == "petuum.com/nodegroup" and value == "adaptdl"


def get_namespace():
    if not os.path.exists(_NS_FILE):
        return "default"
    with open(_NS_FILE) as f:
        return f.read().strip()


def get_image():
    return os.environ["ADAPTDL_IMAGE"]


def get_adaptdl_deployment():
    return os.environ["ADAPTDL_SCHED_DEPLOYMENT"]


def get_supervisor_url():
    return os.environ.get("ADAPTDL_SUPERVISOR_URL", "")


def get_supervisor_port():
    return int(os.getenv("ADAPTDL_SUPERVISOR_SERVICE_PORT", 8080))


def get_storage_subpath():
    return os.environ["ADAPTDL_STORAGE_SUBPATH"]


def get_adaptdl_version():
    return os.environ.get("ADAPTDL_SCHED_VERSION", "0.0.0")


def _json_env(name):
    val = os.getenv(name)
    return json.loads(val) if val else None


def get_job_default_resources():
    return _json_env("ADAPTDL_JOB_DEFAULT_RESOURCES")


def get_job_patch_pods():
    return _json_env("ADAPTDL_JOB_PATCH_PODS")


def get_job_patch_containers():
    return _json_env("ADAPTDL_JOB_PATCH_CONTAINERS")
