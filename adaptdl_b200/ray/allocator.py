"""Feeds Ray's view of the cluster and the trials' ``JobInfo`` into the
Pollux policy (reference: ``adaptdl_ray/adaptdl/adaptdl_allocator.py``)."""

from collections import Counter

from adaptdl_b200.sched.policy import NodeInfo, PolluxPolicy


class AdaptDLAllocator(object):

    def __init__(self, nodes=None, policy=None):
        """``nodes``: ``{address: resources}`` (defaults to the live Ray
        nodes when Ray is available)."""
        if nodes is None:
            from adaptdl_b200.ray import config
            nodes = config.nodes()
        self._nodes = {addr: NodeInfo(dict(res), preemptible=False)
                       for addr, res in nodes.items()}
        biggest = {}
        for info in self._nodes.values():
            for key, val in info.resources.items():
                biggest[key] = max(biggest.get(key, 0), val)
        self._template = NodeInfo(biggest, preemptible=False)
        self._policy = policy or PolluxPolicy()

    def default_allocation(self, num_devices=1):
        """Place ``num_devices`` replicas on the first node(s)."""
        return [list(self._nodes)[0]] * num_devices

    def allocate(self, jobs, nodes=None):
        """``jobs``: objects with ``job_id``, ``job_info`` and
        ``_allocation_in_use()``. Returns ``(allocations, desired_nodes)``
        with allocations restricted to jobs whose placement changed."""
        infos = {job.job_id: job.job_info for job in jobs}
        base = {job.job_id: job._allocation_in_use() for job in jobs}
        allocations, desired = self._policy.optimize(
            infos, self._nodes if nodes is None else nodes, base,
            self._template)
        changed = {key: alloc for key, alloc in allocations.items()
                   if Counter(alloc) != Counter(base.get(key, []))}
        return changed, desired
