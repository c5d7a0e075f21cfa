"""Kubernetes resource arithmetic: quantity parsing, pod requests, what a
node still has to offer (parity: reference ``sched/adaptdl_sched/
resources.py:24-140``). Works on plain dicts or on kubernetes client objects
(anything with ``to_dict()``)."""

import copy
import math

from adaptdl_b200.sched import config

_DECIMAL = ["k", "M", "G", "T", "P", "E"]
_BINARY = ["Ki", "Mi", "Gi", "Ti", "Pi", "Ei"]
_OVERCOMMITABLE = ("cpu", "memory", "ephemeral-storage")


def discretize_resource(name, value):
    """Kubernetes quantity -> integer count of the smallest unit (CPU in
    milli-cores, everything else in base units)."""
    factor = 1000.0 if name == "cpu" else 1.0
    if isinstance(value, str):
        text = value.strip()
        if text.endswith("m"):
            factor /= 1000.0
            text = text[:-1]
        else:
            for power, unit in enumerate(_BINARY, start=1):
                if text.endswith(unit):
                    factor *= 1024.0 ** power
                    text = text[:-2]
                    break
            else:
                for power, unit in enumerate(_DECIMAL, start=1):
                    if text.endswith(unit):
                        factor *= 1000.0 ** power
                        text = text[:-1]
                        break
        value = text
    return int(math.ceil(float(value) * factor - 1e-9))


_discretize_resource = discretize_resource     # reference name


def scale_quantity(name, value, factor):
    """``factor`` times a Kubernetes quantity, as a quantity string (CPU in
    milli-cores, everything else in base units) -- the resources of a pod
    that hosts several replicas."""
    total = discretize_resource(name, value) * int(factor)
    return "{}m".format(total) if name == "cpu" else str(total)


def scale_container_resources(container, factor):
    """Multiply every request and limit of ``container`` (in place)."""
    if factor == 1:
        return container
    resources = container.get("resources") or {}
    for kind in ("requests", "limits"):
        for key, val in list((resources.get(kind) or {}).items()):
            if val is not None:
                resources[kind][key] = scale_quantity(key, val, factor)
    return container


def _as_dict(obj):
    return obj if isinstance(obj, dict) else obj.to_dict()


def get_pod_requests(pod_spec):
    """Aggregate requests of all containers of a pod. Over-committable
    resources count their *requests*; everything else (GPUs and other
    extended resources) counts *limits*; every pod also takes one ``pods``
    slot."""
    spec = _as_dict(pod_spec)
    total = {"pods": 1}
    for container in spec.get("containers") or []:
        res = container.get("resources") or {}
        requests = res.get("requests") or {}
        for key in _OVERCOMMITABLE:
            if requests.get(key) is not None:
                total[key] = total.get(key, 0) \
                    + discretize_resource(key, requests[key])
        for key, val in (res.get("limits") or {}).items():
            if key not in _OVERCOMMITABLE and val is not None:
                total[key] = total.get(key, 0) + discretize_resource(key, val)
    return {k: v for k, v in total.items() if v > 0}


def get_node_unrequested(node, pods):
    """Allocatable resources of ``node`` minus the requests of the
    (non-terminated) ``pods`` scheduled on it. Only positive remainders are
    reported."""
    node = _as_dict(node)
    name = node["metadata"]["name"]
    left = {k: discretize_resource(k, v)
            for k, v in (node["status"].get("allocatable") or {}).items()}
    for pod in pods:
        pod = _as_dict(pod)
        spec = pod.get("spec") or {}
        node_name = spec.get("node_name", spec.get("nodeName"))
        phase = (pod.get("status") or {}).get("phase")
        if node_name != name or phase in ("Succeeded", "Failed"):
            continue
        for key, val in get_pod_requests(spec).items():
            if key in left:
                left[key] -= val
    return {k: v for k, v in left.items() if v > 0}


def set_default_resources(pod_spec):
    """Copy of ``pod_spec`` whose first container has the configured default
    requests/limits filled in where absent."""
    pod_spec = copy.deepcopy(pod_spec)
    defaults = config.get_job_default_resources()
    if defaults:
        container = pod_spec["containers"][0]
        resources = container.setdefault("resources", {})
        for kind in ("requests", "limits"):
            if defaults.get(kind) is not None:
                target = resources.setdefault(kind, {})
                for key, val in defaults[kind].items():
                    target.setdefault(key, val)
    return pod_spec
