"""AdaptDLJob lifecycle controller.

Job phases::

    Pending -> Starting -> Running -> Stopping -> Pending ... | Succeeded | Failed

* *Pending*: waits for the allocator to write ``status.allocation``.
* *Starting*: creates one pod per allocation entry (restart generation
  ``status.group`` += 1) and waits until all are ready.
* *Running*: until the allocation changes (-> Stopping) or pods finish.
* *Stopping*: deletes the pods (SIGTERM -> the trainer checkpoints and exits
  with code 143), then back to Pending.

Behaviour follows the reference's ``sched/adaptdl_sched/controller.py``
(pod naming/annotations/env contract, exit-code-143 and admission-error
tolerance for preemptible jobs, validation of pod groups). The design
differs: :func:`reconcile` is a *pure* function from ``(job, pods, now)`` to
a list of actions, so the whole state machine is unit-tested without a
cluster; :class:`AdaptDLController` only feeds it from watch events and
executes the actions against a :mod:`~adaptdl_b200.sched.kube` backend.
"""

import asyncio
import copy
import logging
from datetime import datetime, timezone

from adaptdl_b200.sched import config, k8s_templates as templates
from adaptdl_b200.sched.kube import ApiError, NotFound
from adaptdl_b200.sched.resources import (scale_container_resources,
                                          set_default_resources)

LOG = logging.getLogger(__name__)

try:                                           # metrics are optional
    from prometheus_client import Counter, Summary
    _KW = dict(namespace="adaptdl", subsystem="sched")
    JOB_SUBMISSION_COUNT = Counter("job_submission_count",
                                   "Number of submitted jobs", **_KW)
    JOB_COMPLETION_TIME = Summary("job_completion_time",
                                  "Duration of completed jobs",
                                  labelnames=["status"], **_KW)
except Exception:  # noqa: BLE001
    JOB_SUBMISSION_COUNT = JOB_COMPLETION_TIME = None

MASTER_PORT_BASE = 47000
EXIT_PREEMPTED = 143


# ---------------------------------------------------------------------------
# pod inspection helpers (plain dict pods)
# ---------------------------------------------------------------------------

def _ann(pod, key):
    return pod["metadata"].get("annotations", {})["adaptdl/" + key]


def _phase(pod):
    return (pod.get("status") or {}).get("phase")


def _parse_time(stamp):
    import datetime
    if isinstance(stamp, datetime.datetime):
        return stamp
    try:
        return datetime.datetime.fromisoformat(
            str(stamp).replace("Z", "+00:00"))
    except ValueError:
        return None


def _job_duration_seconds(job, patch):
    """completionTimestamp - metadata.creationTimestamp (0 if unknown)."""
    import datetime
    created = _parse_time((job.get("metadata") or {})
                          .get("creationTimestamp"))
    done = _parse_time(patch.get("completionTimestamp")) or \
        datetime.datetime.now(datetime.timezone.utc)
    if created is None:
        return 0.0
    if created.tzinfo is None:
        created = created.replace(tzinfo=datetime.timezone.utc)
    if done.tzinfo is None:
        done = done.replace(tzinfo=datetime.timezone.utc)
    return max((done - created).total_seconds(), 0.0)


def _deleting(pod):
    return pod["metadata"].get("deletionTimestamp") is not None


def count_ready_pods(pods):
    n = 0
    for pod in pods:
        statuses = (pod.get("status") or {}).get("containerStatuses")
        if statuses and all(s.get("ready") for s in statuses):
            n += 1
    return n


def count_scheduled_pods(pods):
    n = 0
    for pod in pods:
        conditions = (pod.get("status") or {}).get("conditions") or []
        if any(c.get("type") == "PodScheduled" and c.get("status") == "True"
               for c in conditions):
            n += 1
    return n


def _exited_143(pod):
    for status in (pod.get("status") or {}).get("containerStatuses") or []:
        terminated = (status.get("state") or {}).get("terminated")
        if terminated and terminated.get("exitCode") == EXIT_PREEMPTED:
            return True
    return False


def _local_replicas(pod):
    """Replicas hosted by ``pod`` (1 unless the job uses one pod per node)."""
    return int(pod["metadata"].get("annotations", {}).get(
        "adaptdl/local-replicas", 1))


def count_replicas(pods):
    return sum(_local_replicas(pod) for pod in pods)


def canonical_allocation(allocation, pod_per_node=False):
    """The allocation in rank order. With one pod per node the replicas of
    a node take consecutive ranks (nodes in order of first appearance), so
    that each pod hosts one contiguous rank range."""
    if not pod_per_node:
        return list(allocation)
    order = {}
    for node in allocation:
        order.setdefault(node, len(order))
    return sorted(allocation, key=order.__getitem__)


def plan_pods(allocation, pod_per_node=False):
    """``[(first_rank, node, replicas)]``, one entry per pod to create."""
    ranked = canonical_allocation(allocation, pod_per_node)
    if not pod_per_node:
        return [(rank, node, 1) for rank, node in enumerate(ranked)]
    plan = []
    for rank, node in enumerate(ranked):
        if plan and plan[-1][1] == node:
            plan[-1] = (plan[-1][0], node, plan[-1][2] + 1)
        else:
            plan.append((rank, node, 1))
    return plan


def validate_pods(pods):
    """``None`` if the job's pods form one consistent group, else the
    failure message."""
    groups, replicas, ranks = set(), set(), []
    for pod in pods:
        name = pod["metadata"]["name"]
        try:
            groups.add(int(_ann(pod, "group")))
            replicas.add(int(_ann(pod, "replicas")))
            ranks.append(int(_ann(pod, "rank")) + _local_replicas(pod) - 1)
            node = _ann(pod, "node")
        except (KeyError, ValueError):
            return "invalid annotations for pod {}".format(name)
        bound = (pod.get("spec") or {}).get("nodeName")
        if bound and bound != node:
            return "incorrect node for pod {}".format(name)
    if len(groups) > 1 or len(replicas) > 1 or \
            (replicas and any(r >= next(iter(replicas)) for r in ranks)):
        return "inconsistent pods in group"
    return None


def detect_completion(pods, preemptible):
    """Status patch if the job finished (``Succeeded`` / ``Failed``), else
    ``{}``. Preempted pods (deleted, exit code 143), admission errors and
    ``OutOf*`` evictions are *not* failures."""
    if not pods:
        return {}
    want = {int(_ann(p, "replicas")) for p in pods}
    if all(_phase(p) == "Succeeded" for p in pods) and \
            want == {count_replicas(pods)}:
        return {"phase": "Succeeded"}
    for pod in pods:
        if _phase(pod) == "Unknown":
            LOG.warning("Unknown status for pod %s", pod["metadata"]["name"])
        elif _phase(pod) != "Failed":
            continue
        reason = str((pod.get("status") or {}).get("reason"))
        if reason == "UnexpectedAdmissionError" or reason.startswith("Outof"):
            LOG.warning("pod %s: %s", pod["metadata"]["name"], reason)
        elif preemptible and (_deleting(pod) or _exited_143(pod)):
            LOG.warning("pod %s terminated", pod["metadata"]["name"])
        else:
            return {"phase": "Failed", "reason": "PodFailure",
                    "message": "{} {}".format(pod["metadata"]["name"],
                                              _phase(pod))}
    return {}


def detect_restart(pods, allocation, pod_per_node=False):
    """True if the running pods no longer match ``status.allocation``."""
    ranked = canonical_allocation(allocation, pod_per_node)
    for pod in pods:
        replicas, rank = int(_ann(pod, "replicas")), int(_ann(pod, "rank"))
        hosted = ranked[rank:rank + _local_replicas(pod)] \
            if replicas == len(ranked) else []
        if len(hosted) != _local_replicas(pod) or \
                any(node != _ann(pod, "node") for node in hosted):
            return True
        if pod_per_node and ranked.count(_ann(pod, "node")) != len(hosted):
            return True                  # the node's share changed
    return False


# ---------------------------------------------------------------------------
# the pod a replica runs in
# ---------------------------------------------------------------------------

def pod_name(job_metadata, group, rank):
    return "{}-{}-{}-{}".format(job_metadata["name"], job_metadata["uid"],
                                group, rank)


def _apply_json_patch(doc, patch):
    """Minimal RFC-6902 (add / replace / remove), enough for the helm
    values' pod/container patches (the reference uses ``jsonpatch``)."""
    doc = copy.deepcopy(doc)
    for op in patch:
        parts = [p.replace("~1", "/").replace("~0", "~")
                 for p in op["path"].lstrip("/").split("/")]
        parent = doc
        for part in parts[:-1]:
            parent = parent[int(part)] if isinstance(parent, list) \
                else parent.setdefault(part, {})
        last = parts[-1]
        if isinstance(parent, list):
            idx = len(parent) if last == "-" else int(last)
            if op["op"] == "add":
                parent.insert(idx, op["value"])
            elif op["op"] == "replace":
                parent[idx] = op["value"]
            elif op["op"] == "remove":
                parent.pop(idx)
        elif op["op"] in ("add", "replace"):
            parent[last] = op["value"]
        elif op["op"] == "remove":
            parent.pop(last, None)
    return doc


def build_pod(job_metadata, pod_template, allocation, group, rank,
              node_hostname, local_replicas=1):
    """Manifest of the pod hosting replicas ``rank .. rank+local_replicas-1``
    of restart generation ``group``: pinned to its node, memory-backed
    ``/dev/shm``, and the ``ADAPTDL_*`` environment the trainer reads
    (``adaptdl_b200.env``). With ``local_replicas`` > 1 (``spec.podPerNode``)
    every container's requests and limits are multiplied and
    ``ADAPTDL_LOCAL_REPLICAS`` tells ``python -m adaptdl_b200.launch`` how
    many rank processes to start -- all the job's GPUs on the node are then
    inside ONE container, which is what the peer-memory gradient reducer
    needs (one-GPU pods fall back to NCCL)."""
    pod = copy.deepcopy(pod_template)
    pod["apiVersion"], pod["kind"] = "v1", "Pod"
    meta = pod.setdefault("metadata", {})
    meta["name"] = pod_name(job_metadata, group, rank)
    meta["ownerReferences"] = templates.owner_reference_template(
        job_metadata["namespace"], job_metadata["name"], job_metadata["uid"])
    labels = meta.setdefault("labels", {})
    labels.update({"adaptdl": "true", "adaptdl/job": job_metadata["name"],
                   "petuum.com/nodegroup": "all"})
    meta.setdefault("annotations", {}).update({
        "adaptdl/replicas": str(len(allocation)),
        "adaptdl/group": str(group), "adaptdl/rank": str(rank),
        "adaptdl/node": allocation[rank]})
    if local_replicas != 1:
        meta["annotations"]["adaptdl/local-replicas"] = str(local_replicas)
    spec = pod["spec"]
    spec["hostname"] = "{}-{}-{}".format(job_metadata["name"], group, rank)
    spec.setdefault("nodeSelector", {})["kubernetes.io/hostname"] = \
        node_hostname
    spec["restartPolicy"] = "Never"
    spec.setdefault("volumes", []).append(
        {"name": "adaptdl-shm", "emptyDir": {"medium": "Memory"}})
    pod["spec"] = spec = set_default_resources(spec)
    env = [
        ("ADAPTDL_JOB_ID", "{}/{}".format(job_metadata["namespace"],
                                          job_metadata["name"])),
        ("ADAPTDL_MASTER_PORT", str(MASTER_PORT_BASE + group)),
        ("ADAPTDL_NUM_NODES", str(len(set(allocation)))),
        ("ADAPTDL_NUM_RESTARTS", str(group)),
        ("ADAPTDL_NUM_REPLICAS", str(len(allocation))),
        ("ADAPTDL_REPLICA_RANK", str(rank)),
        ("ADAPTDL_SUPERVISOR_URL", config.get_supervisor_url()),
        ("ADAPTDL_SCHED_VERSION", config.get_adaptdl_version()),
    ]
    if local_replicas != 1:
        env.append(("ADAPTDL_LOCAL_REPLICAS", str(local_replicas)))
    for container in spec["containers"]:
        scale_container_resources(container, local_replicas)
        container.setdefault("volumeMounts", []).append(
            {"name": "adaptdl-shm", "mountPath": "/dev/shm"})
        cenv = container.setdefault("env", [])
        cenv.extend({"name": k, "value": v} for k, v in env)
        limits = (container.get("resources") or {}).get("limits") or {}
        if not limits.get("nvidia.com/gpu"):
            cenv.append({"name": "NVIDIA_VISIBLE_DEVICES", "value": "none"})
    if config.get_job_patch_pods():
        pod = _apply_json_patch(pod, config.get_job_patch_pods())
    if config.get_job_patch_containers():
        patch = config.get_job_patch_containers()
        pod["spec"]["containers"] = [_apply_json_patch(c, patch)
                                     for c in pod["spec"]["containers"]]
    return pod


# ---------------------------------------------------------------------------
# the state machine
# ---------------------------------------------------------------------------

def reconcile(job, pods, now=None):
    """Decide what to do for one job.

    Returns ``(status_patch, actions)`` where actions are
    ``("delete_pods", [pods])`` / ``("create_pods", group, allocation)``.
    ``job`` is ``None`` when the job object no longer exists.
    """
    now = now or datetime.now(timezone.utc).isoformat()
    if job is None:
        return {}, [("delete_pods", list(pods))]
    invalid = validate_pods(pods)
    if invalid:
        return {"phase": "Failed", "reason": "Invalid",
                "message": invalid}, []
    old = job.get("status") or {}
    new = dict(old)
    actions = []
    allocation = new.get("allocation") or []
    phase = new.setdefault("phase", "Pending")
    replicas = new.get("replicas") or 0
    preemptible = job["spec"].get("preemptible", True)
    per_node = bool(job["spec"].get("podPerNode", False))
    want_pods = len(plan_pods(allocation, per_node))
    completion = detect_completion(pods, preemptible) \
        if phase not in ("Succeeded", "Failed") else {}
    if completion:
        new.update(completion)
        new.setdefault("completionTimestamp", now)
        new["allocation"] = allocation = []
        # failed pods are kept for debugging
        actions.append(("delete_pods",
                        [p for p in pods if _phase(p) != "Failed"]))
    elif phase == "Pending":
        if allocation and not pods:
            new["phase"] = "Starting"
    elif phase == "Starting":
        if not allocation or (count_scheduled_pods(pods) != want_pods
                              and detect_restart(pods, allocation,
                                                 per_node)):
            new["phase"] = "Stopping"
        elif not pods:
            new["group"] = new.get("group", -1) + 1
            actions.append(("create_pods", new["group"], list(allocation)))
        elif len(pods) != want_pods:
            new["phase"] = "Stopping"
        elif count_ready_pods(pods) == want_pods:
            new["phase"] = "Running"
    elif phase == "Running":
        if not pods or detect_restart(pods, allocation, per_node):
            new["phase"] = "Stopping"
    elif phase == "Stopping":
        if pods:
            actions.append(("delete_pods", list(pods)))
        else:
            new["phase"] = "Pending"
    if allocation:
        new["replicas"] = len(allocation)
        new["readyReplicas"] = count_replicas(
            [p for p in pods if count_ready_pods([p])])
    else:
        new["allocation"] = new["replicas"] = new["readyReplicas"] = None
    patch = {k: v for k, v in new.items() if old.get(k) != v}
    return patch, actions


class AdaptDLController(object):
    """Feeds :func:`reconcile` from watch events; one job is never processed
    concurrently with itself (single worker draining a queue)."""

    def __init__(self, cluster):
        self._cluster = cluster
        self._queue = asyncio.Queue()

    async def run(self):
        await asyncio.gather(self._watch(), self._sync_worker())

    async def _watch(self):
        async for kind, obj in self._cluster.watch():
            meta = obj["metadata"]
            if kind == "job":
                await self._queue.put((meta["namespace"], meta["name"]))
            else:
                name = meta.get("labels", {}).get("adaptdl/job")
                if name:
                    await self._queue.put((meta["namespace"], name))

    async def _sync_worker(self):
        while True:
            namespace, name = await self._queue.get()
            try:
                await self.sync_job(namespace, name)
            except Exception:  # noqa: BLE001 - keep the controller alive
                LOG.exception("sync of %s/%s failed", namespace, name)
            self._queue.task_done()

    async def sync_job(self, namespace, name):
        cluster = self._cluster
        pods = await cluster.list_pods(
            namespace, label_selector="adaptdl/job={}".format(name))
        try:
            job = await cluster.get_job(namespace, name)
        except NotFound:
            job = None
        patch, actions = reconcile(job, pods)
        for action in actions:
            if action[0] == "delete_pods":
                await self._delete_pods(action[1])
            elif action[0] == "create_pods":
                failure = await self._create_pods(job, action[1], action[2])
                if failure:
                    patch.update(failure)
        if patch and job is not None:
            LOG.info("Patch AdaptDLJob %s: %s", name, patch)
            if JOB_COMPLETION_TIME is not None and \
                    patch.get("phase") in ("Succeeded", "Failed"):
                JOB_COMPLETION_TIME.labels(patch["phase"]).observe(
                    _job_duration_seconds(job, patch))
            if JOB_SUBMISSION_COUNT is not None and \
                    patch.get("phase") == "Pending" and \
                    not (job.get("status") or {}).get("phase"):
                JOB_SUBMISSION_COUNT.inc()       # first time we see the job
            await cluster.patch_job_status(namespace, name, {"status": patch})
        return patch

    async def _delete_pods(self, pods):
        doomed = [p for p in pods if not _deleting(p)]
        if doomed:
            LOG.info("Deleting %s", [p["metadata"]["name"] for p in doomed])
            await asyncio.gather(
                *[self._cluster.delete_pod(p["metadata"]["namespace"],
                                           p["metadata"]["name"])
                  for p in doomed], return_exceptions=True)

    async def _create_pods(self, job, group, allocation):
        meta = job["metadata"]
        created = []
        try:
            per_node = bool(job["spec"].get("podPerNode", False))
            ranked = canonical_allocation(allocation, per_node)
            for rank, node_name, count in plan_pods(allocation, per_node):
                node = await self._cluster.read_node(node_name)
                hostname = node["metadata"].get("labels", {}).get(
                    "kubernetes.io/hostname", node["metadata"]["name"])
                pod = build_pod(meta, job["spec"]["template"], ranked,
                                group, rank, hostname, count)
                created.append(await self._cluster.create_pod(
                    meta["namespace"], pod))
        except (ApiError, NotFound) as exc:
            LOG.warning("Failed to create pod for %s: %s", meta["name"], exc)
            await self._delete_pods(created)
            return {"phase": "Failed", "reason": "PodCreationError",
                    "message": str(exc)}
        return None


if __name__ == "__main__":      # ``python -m adaptdl_sched.controller``, as in
    from adaptdl_b200.sched.__main__ import main  # the reference's chart
    main(["controller"])
