"""Cluster backend used by the controller / allocator / supervisor.

Everything above this module speaks plain JSON-style dicts (the Kubernetes
REST shapes, camelCase). Two backends:

* :class:`InMemoryCluster` -- a tiny fake API server (jobs, pods, nodes,
  watches) used by the unit tests and by the single-box local scheduler.
* :class:`KubernetesCluster` -- adapter over ``kubernetes_asyncio`` (imported
  lazily; the package is optional).

Keeping the reconciliation logic independent of the client library is what
makes the controller and allocator testable here (the reference tests
neither).
"""

import asyncio
import copy
import itertools

from adaptdl_b200.sched import config


class NotFound(Exception):
    pass


class ApiError(Exception):
    def __init__(self, status, message=""):
        super().__init__("{} {}".format(status, message))
        self.status = status
        self.message = message


def job_key(job):
    return job["metadata"]["namespace"], job["metadata"]["name"]


class InMemoryCluster(object):
    """A fake cluster. Pods never run by themselves; tests (or the local
    launcher) drive ``set_pod_status``."""

    def __init__(self):
        self.jobs = {}       # (ns, name) -> job dict
        self.pods = {}       # (ns, name) -> pod dict
        self.nodes = {}      # name -> node dict
        self._watchers = []
        self._uid = itertools.count(1)
        self.fail_pod_creation = False

    # -- test helpers ------------------------------------------------------

    def add_node(self, name, allocatable, labels=None, taints=None):
        labels = dict(labels or {})
        labels.setdefault("kubernetes.io/hostname", name)
        self.nodes[name] = {
            "metadata": {"name": name, "labels": labels},
            "spec": {"taints": taints or []},
            "status": {"allocatable": dict(allocatable)},
        }

    def add_job(self, namespace, name, spec, status=None):
        job = {"apiVersion": "adaptdl.petuum.com/v1", "kind": "AdaptDLJob",
               "metadata": {"namespace": namespace, "name": name,
                            "uid": "uid-{}".format(next(self._uid)),
                            "creationTimestamp": next(self._uid)},
               "spec": copy.deepcopy(spec),
               "status": copy.deepcopy(status or {})}
        self.jobs[(namespace, name)] = job
        self._notify("job", job)
        return job

    def set_pod_status(self, namespace, name, **status):
        pod = self.pods[(namespace, name)]
        pod.setdefault("status", {}).update(status)
        self._notify("pod", pod)

    def _notify(self, kind, obj):
        for queue in self._watchers:
            queue.put_nowait((kind, copy.deepcopy(obj)))

    # -- API ---------------------------------------------------------------

    async def watch(self):
        """Async iterator of ``(kind, object)`` change events."""
        queue = asyncio.Queue()
        self._watchers.append(queue)
        try:
            while True:
                yield await queue.get()
        finally:
            self._watchers.remove(queue)

    async def list_jobs(self):
        return [copy.deepcopy(j) for j in self.jobs.values()]

    async def get_job(self, namespace, name):
        try:
            return copy.deepcopy(self.jobs[(namespace, name)])
        except KeyError:
            raise NotFound(name)

    async def patch_job_status(self, namespace, name, patch):
        job = self.jobs.get((namespace, name))
        if job is None:
            return None
        status = job.setdefault("status", {})
        for key, val in patch.get("status", {}).items():
            if val is None:
                status.pop(key, None)
            else:
                status[key] = copy.deepcopy(val)
        self._notify("job", job)
        return copy.deepcopy(job)

    async def delete_job(self, namespace, name):
        job = self.jobs.pop((namespace, name), None)
        if job is not None:
            self._notify("job", job)

    async def list_pods(self, namespace=None, label_selector=None):
        out = []
        for (ns, _), pod in self.pods.items():
            if namespace and ns != namespace:
                continue
            labels = pod["metadata"].get("labels", {})
            if label_selector and not _matches(labels, label_selector):
                continue
            out.append(copy.deepcopy(pod))
        return out

    async def create_pod(self, namespace, pod, dry_run=False):
        if self.fail_pod_creation:
            raise ApiError(422, "pod rejected")
        if not pod.get("spec", {}).get("containers"):
            raise ApiError(422, "pod has no containers")
        if dry_run:
            return copy.deepcopy(pod)
        pod = copy.deepcopy(pod)
        pod["metadata"]["namespace"] = namespace
        if not pod["metadata"].get("name"):
            # what the API server does with metadata.generateName
            import uuid
            pod["metadata"]["name"] = \
                pod["metadata"].get("generateName", "pod-") \
                + uuid.uuid4().hex[:5]
        elif (namespace, pod["metadata"]["name"]) in self.pods:
            raise ApiError(409, "pod already exists")
        pod.setdefault("status", {"phase": "Pending"})
        self.pods[(namespace, pod["metadata"]["name"])] = pod
        self._notify("pod", pod)
        return copy.deepcopy(pod)

    async def delete_pod(self, namespace, name):
        pod = self.pods.pop((namespace, name), None)
        if pod is not None:
            self._notify("pod", pod)

    async def list_nodes(self):
        return [copy.deepcopy(n) for n in self.nodes.values()]

    async def read_node(self, name):
        try:
            return copy.deepcopy(self.nodes[name])
        except KeyError:
            raise NotFound(name)


def _matches(labels, selector):
    for clause in selector.split(","):
        clause = clause.strip()
        if not clause:
            continue
        if "!=" in clause:
            key, val = clause.split("!=", 1)
            if labels.get(key) == val:
                return False
        elif "=" in clause:
            key, val = clause.split("=", 1)
            if labels.get(key.rstrip("=")) != val:
                return False
        elif clause.startswith("!"):
            if clause[1:] in labels:
                return False
        elif clause not in labels:
            return False
    return True


class KubernetesCluster(object):
    """Adapter over ``kubernetes_asyncio`` returning plain dicts."""

    def __init__(self):
        import kubernetes_asyncio as k8s      # optional dependency
        self._k8s = k8s
        self._core = k8s.client.CoreV1Api()
        self._objs = k8s.client.CustomObjectsApi()
        self._ser = k8s.client.ApiClient()

    def _plain(self, obj):
        return self._ser.sanitize_for_serialization(obj)

    async def watch(self):
        queue = asyncio.Queue()

        async def pump(kind, fn, *args, **kwargs):
            async with self._k8s.watch.Watch() as watch:
                while True:
                    async for event in watch.stream(fn, *args,
                                                    timeout_seconds=60,
                                                    **kwargs):
                        await queue.put((kind, self._plain(event["object"])))
        tasks = [
            asyncio.ensure_future(pump(
                "job", self._objs.list_namespaced_custom_object,
                config.GROUP, config.VERSION, "", config.PLURAL)),
            asyncio.ensure_future(pump(
                "pod", self._core.list_namespaced_pod, "",
                label_selector="adaptdl/job")),
        ]
        try:
            while True:
                yield await queue.get()
        finally:
            for task in tasks:
                task.cancel()

    async def list_jobs(self):
        out = await self._objs.list_namespaced_custom_object(
            config.GROUP, config.VERSION, "", config.PLURAL)
        return out["items"]

    async def get_job(self, namespace, name):
        try:
            return await self._objs.get_namespaced_custom_object(
                config.GROUP, config.VERSION, namespace, config.PLURAL, name)
        except self._k8s.client.rest.ApiException as exc:
            if exc.status == 404:
                raise NotFound(name)
            raise

    async def patch_job_status(self, namespace, name, patch):
        try:
            return await self._objs.patch_namespaced_custom_object_status(
                config.GROUP, config.VERSION, namespace, config.PLURAL, name,
                patch)
        except self._k8s.client.rest.ApiException as exc:
            if exc.status == 404:
                return None
            raise

    async def list_pods(self, namespace=None, label_selector=None):
        kwargs = {"label_selector": label_selector} if label_selector else {}
        if namespace:
            pods = await self._core.list_namespaced_pod(namespace, **kwargs)
        else:
            pods = await self._core.list_pod_for_all_namespaces(**kwargs)
        return [self._plain(p) for p in pods.items]

    async def create_pod(self, namespace, pod, dry_run=False):
        kwargs = {"dry_run": "All"} if dry_run else {}
        try:
            return self._plain(await self._core.create_namespaced_pod(
                namespace, pod, **kwargs))
        except self._k8s.client.rest.ApiException as exc:
            raise ApiError(exc.status, str(exc))

    async def delete_pod(self, namespace, name):
        try:
            await self._core.delete_namespaced_pod(name, namespace)
        except self._k8s.client.rest.ApiException as exc:
            if exc.status != 404:
                raise

    async def list_nodes(self):
        nodes = await self._core.list_node()
        return [self._plain(n) for n in nodes.items]

    async def read_node(self, name):
        return self._plain(await self._core.read_node(name))
