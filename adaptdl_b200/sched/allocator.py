"""Allocator: turns scheduling hints into replica placements.

Two activities under one lock (parity: reference ``sched/adaptdl_sched/
allocator.py:37-293``):

* **fast path** -- a newly submitted preemptible job gets a first-fit
  placement immediately (``PolluxPolicy.allocate_job``) instead of waiting
  for the next cycle;
* **optimisation cycle** (every ``period`` seconds) -- every active job's
  hints (``status.train``: fitted perf params, gradient statistics, batch
  size limits) become a goodput-based speedup function, and
  ``PolluxPolicy.optimize`` re-allocates the whole cluster; the result is
  written to each job's ``status.allocation`` (the controller reacts) and the
  desired cluster size drives the :class:`ClusterExpander`.

Talks to the cluster only through a :mod:`~adaptdl_b200.sched.kube` backend.
"""

import asyncio
import logging
import time

from adaptdl_b200.goodput import GoodputFunction, GradParams, PerfParams
from adaptdl_b200.sched import config, metrics
from adaptdl_b200.sched.kube import NotFound
from adaptdl_b200.sched.policy import (JobInfo, NodeInfo, PolluxPolicy,
                                       SpeedupFunction)
from adaptdl_b200.sched.resources import (get_node_unrequested,
                                          get_pod_requests,
                                          set_default_resources)
from adaptdl_b200.sched_hints import PERF_PARAMS

LOG = logging.getLogger(__name__)
ACTIVE_PHASES = ("Pending", "Starting", "Running", "Stopping")


def _linear_speedup(nodes, replicas):
    return replicas


def job_info_from(job):
    """Build the policy's :class:`JobInfo` from an AdaptDLJob object."""
    spec = job["spec"]
    pod_spec = set_default_resources(spec["template"]["spec"])
    resources = get_pod_requests(pod_spec)
    hints = (job.get("status") or {}).get("train") or {}
    # exploration is gradual: at most twice what has been profiled
    max_replicas = max(2 * hints.get("maxProfiledReplicas", 0), 1)
    if spec.get("maxReplicas"):
        max_replicas = min(max_replicas, spec["maxReplicas"])
    min_replicas = spec.get("minReplicas", 0)
    max_replicas = max(max_replicas, min_replicas)
    preemptible = spec.get("preemptible", True)
    speedup_fn = _linear_speedup
    if hints.get("perfParams") and hints.get("initBatchSize") and preemptible:
        max_batch_size = hints.get("maxBatchSize") or hints["initBatchSize"]
        bounds = hints.get("localBszBounds")
        if bounds:
            min_local = bounds[0] or 1
            if max_batch_size < min_local * max_replicas:
                max_replicas = max(int(max_batch_size / min_local), 1)
        perf = PerfParams(*[hints["perfParams"][k] for k in PERF_PARAMS])
        grad = hints.get("gradParams")
        grad = GradParams(grad["norm"], grad["var"]) if grad \
            else GradParams(0.0, 1.0)
        speedup_fn = SpeedupFunction(
            GoodputFunction(perf, grad, hints["initBatchSize"]),
            hints.get("maxBatchSize"), tuple(bounds) if bounds else None,
            hints.get("gradientAccumulation", False))
    return JobInfo(resources, speedup_fn,
                   job["metadata"].get("creationTimestamp"),
                   min_replicas, max(max_replicas, min_replicas, 1),
                   preemptible)


class AdaptDLAllocator(object):

    def __init__(self, cluster, expander=None, policy=None, period=60.0):
        self._cluster = cluster
        self._expander = expander
        self._policy = policy or PolluxPolicy()
        self._period = period
        self._lock = asyncio.Lock()
        self._desired_nodes = None

    async def run(self):
        await asyncio.gather(self._allocate_one_loop(),
                             self._optimize_all_loop())

    # -- fast path -----------------------------------------------------------

    async def _allocate_one_loop(self):
        seen = set()
        async for kind, obj in self._cluster.watch():
            if kind != "job":
                continue
            key = (obj["metadata"]["namespace"], obj["metadata"]["name"])
            if key in seen or not obj["spec"].get("preemptible", True):
                continue
            seen.add(key)
            async with self._lock:
                await self.allocate_one(*key)

    async def allocate_one(self, namespace, name):
        try:
            job = await self._cluster.get_job(namespace, name)
        except NotFound:
            return None
        status = job.get("status") or {}
        if status.get("allocation") is not None or \
                status.get("group") is not None:
            return None                      # already known to the scheduler
        nodes, _ = await self.find_nodes()
        allocation = self._policy.allocate_job(job_info_from(job), nodes)
        LOG.info("new job %s/%s -> %s", namespace, name, allocation)
        await self._cluster.patch_job_status(
            namespace, name, {"status": {"allocation": allocation}})
        return allocation

    # -- full cycle -------------------------------------------------------------

    async def _optimize_all_loop(self):
        while True:
            async with self._lock:
                try:
                    await self.optimize_all()
                except Exception:  # noqa: BLE001
                    LOG.exception("allocation cycle failed")
            await asyncio.sleep(self._period)

    async def optimize_all(self):
        nodes, template = await self.find_nodes(
            pod_label_selector="!adaptdl/job")
        jobs, previous = await self.find_jobs_and_allocations()
        known = list(jobs)
        start = time.time()
        # the genetic search takes seconds on a big cluster: keep it off the
        # event loop so that watches, the fast path's API calls and the
        # cluster expander keep being served meanwhile
        allocations = await asyncio.get_running_loop().run_in_executor(
            None, self.allocate, jobs, nodes, previous, template)
        elapsed = time.time() - start
        LOG.info("allocations (%.3f s): %s", elapsed, allocations)
        await self.update_allocations(allocations)
        metrics.observe_cycle(allocations, nodes, seconds=elapsed,
                              desired_nodes=self._desired_nodes,
                              known_jobs=known)
        return allocations

    async def find_nodes(self, pod_label_selector=None):
        """Usable nodes with what they still offer (net of the selected
        pods), plus the autoscaling node template (per-resource maximum)."""
        pods = await self._cluster.list_pods(
            label_selector=pod_label_selector)
        infos = {}
        for node in await self._cluster.list_nodes():
            if not config.allowed_taints((node.get("spec") or {})
                                         .get("taints")):
                continue
            resources = get_node_unrequested(node, pods)
            if not resources.get("pods"):
                LOG.warning("node %s has no free pod slots",
                            node["metadata"]["name"])
            infos[node["metadata"]["name"]] = NodeInfo(resources, False)
        biggest = {}
        for info in infos.values():
            for key, val in info.resources.items():
                biggest[key] = max(biggest.get(key, 0), val)
        return infos, NodeInfo(biggest, True)

    async def find_jobs_and_allocations(self):
        infos, allocations = {}, {}
        for job in await self._cluster.list_jobs():
            status = job.get("status") or {}
            if status.get("phase", "Pending") not in ACTIVE_PHASES:
                continue
            key = (job["metadata"]["namespace"], job["metadata"]["name"])
            if status.get("allocation") is not None:
                allocations[key] = list(status["allocation"])
            infos[key] = job_info_from(job)
        return infos, allocations

    def allocate(self, jobs, nodes, previous, template):
        for key in list(jobs):               # drop jobs no node can host
            need = jobs[key].resources
            if not any(all(val <= node.resources.get(res, 0)
                           for res, val in need.items())
                       for node in nodes.values()):
                LOG.warning("job %s cannot be scheduled on any node", key)
                jobs.pop(key)
        allocations, active = {}, []
        self._desired_nodes = None
        if jobs and nodes:
            allocations, desired = self._policy.optimize(
                jobs, nodes, previous, template)
            self._desired_nodes = desired
            if desired < len(nodes):
                used = [set(a) for a in allocations.values() if a]
                active = sorted(set().union(*used)) if used else []
            else:
                active = list(nodes)
                active += ["~{}".format(i + 1)
                           for i in range(desired - len(nodes))]
        elif jobs:
            active = ["~1"]                  # no node at all: ask for one
        if self._expander is not None:
            self._expander.fit(active)
        return allocations

    async def update_allocations(self, allocations):
        for job in await self._cluster.list_jobs():
            key = (job["metadata"]["namespace"], job["metadata"]["name"])
            current = (job.get("status") or {}).get("allocation") or []
            wanted = list(allocations.get(key, []))
            if list(current) != wanted:
                LOG.info("Patch AdaptDLJob %s/%s allocation: %s", *key,
                         wanted)
                await self._cluster.patch_job_status(
                    key[0], key[1], {"status": {"allocation": wanted}})


if __name__ == "__main__":      # ``python -m adaptdl_sched.allocator``, as in
    from adaptdl_b200.sched.__main__ import main  # the reference's chart
    main(["allocator"])
