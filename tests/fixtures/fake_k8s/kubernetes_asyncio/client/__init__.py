import copy

from . import rest  # noqa: F401
from .rest import ApiException

# the "API server": shared by all Api objects of a test
STATE = {"pods": {}, "jobs": {}, "nodes": {}, "events": []}


def reset():
    for value in STATE.values():
        value.clear()


def _dump():
    """``FAKE_K8S_DUMP``: keep a JSON copy of the pods on disk so that a test
    can watch what a scheduler process it started has created / patched."""
    import json
    import os
    path = os.environ.get("FAKE_K8S_DUMP")
    if path:
        with open(path + ".tmp", "w") as f:
            json.dump({"pods": list(STATE["pods"].values()),
                       "jobs": list(STATE["jobs"].values())}, f)
        os.replace(path + ".tmp", path)


class _Model(object):
    """What the generated client returns for core objects."""

    def __init__(self, body):
        self._body = copy.deepcopy(body)

    def to_dict(self):
        return copy.deepcopy(self._body)


class _List(object):
    def __init__(self, items):
        self.items = items


class ApiClient(object):
    def sanitize_for_serialization(self, obj):
        if isinstance(obj, _Model):
            return obj.to_dict()
        if isinstance(obj, list):
            return [self.sanitize_for_serialization(o) for o in obj]
        return copy.deepcopy(obj)


def _match(pod, selector):
    if not selector:
        return True
    labels = pod.get("metadata", {}).get("labels", {})
    for clause in selector.split(","):
        if "=" in clause:
            key, val = clause.split("=", 1)
            if labels.get(key) != val:
                return False
        elif clause not in labels:
            return False
    return True


class CoreV1Api(object):
    async def list_namespaced_pod(self, namespace, label_selector=None,
                                  **kwargs):
        return _List([_Model(p) for (ns, _), p in STATE["pods"].items()
                      if (not namespace or ns == namespace)
                      and _match(p, label_selector)])

    async def list_pod_for_all_namespaces(self, label_selector=None,
                                          **kwargs):
        return await self.list_namespaced_pod("", label_selector)

    async def create_namespaced_pod(self, namespace, body, dry_run=None):
        meta = body.setdefault("metadata", {})
        if not meta.get("name"):
            meta["name"] = meta.get("generateName", "pod-") + "x%04d" % len(
                STATE["pods"])
        key = (namespace, meta["name"])
        if not body.get("spec", {}).get("containers"):
            raise ApiException(422, "no containers")
        if key in STATE["pods"]:
            raise ApiException(409, "exists")
        pod = copy.deepcopy(body)
        pod["metadata"]["namespace"] = namespace
        pod.setdefault("status", {"phase": "Pending"})
        if dry_run != "All":
            STATE["pods"][key] = pod
            STATE["events"].append(("pod", "ADDED", pod))
            _dump()
        return _Model(pod)

    async def create_namespaced_pod_template(self, namespace, body,
                                             dry_run=None):
        spec = (body.get("template") or {}).get("spec") or {}
        if not spec.get("containers"):
            exc = ApiException(422, "Unprocessable Entity")
            exc.body = ('{"kind": "Status", "message": "PodTemplate is '
                        'invalid: template.spec.containers: Required value"}')
            raise exc
        return _Model(copy.deepcopy(body))

    async def delete_namespaced_pod(self, name, namespace):
        if (namespace, name) not in STATE["pods"]:
            raise ApiException(404, "not found")
        pod = STATE["pods"].pop((namespace, name))
        STATE["events"].append(("pod", "DELETED", pod))

    async def list_node(self):
        return _List([_Model(n) for n in STATE["nodes"].values()])

    async def read_node(self, name):
        if name not in STATE["nodes"]:
            raise ApiException(404, "not found")
        return _Model(STATE["nodes"][name])


class CustomObjectsApi(object):
    async def list_namespaced_custom_object(self, group, version, namespace,
                                            plural, **kwargs):
        return {"items": [copy.deepcopy(j) for (ns, _), j in
                          STATE["jobs"].items()
                          if not namespace or ns == namespace]}

    async def get_namespaced_custom_object(self, group, version, namespace,
                                           plural, name):
        if (namespace, name) not in STATE["jobs"]:
            raise ApiException(404, "not found")
        return copy.deepcopy(STATE["jobs"][(namespace, name)])

    async def patch_namespaced_custom_object_status(self, group, version,
                                                    namespace, plural, name,
                                                    body):
        if (namespace, name) not in STATE["jobs"]:
            raise ApiException(404, "not found")
        job = STATE["jobs"][(namespace, name)]
        job.setdefault("status", {}).update(body.get("status", {}))
        STATE["events"].append(("job", "MODIFIED", job))
        _dump()
        return copy.deepcopy(job)
