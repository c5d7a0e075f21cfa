"""Plain descriptions of jobs and nodes handed to the policy (parity:
reference ``sched/adaptdl_sched/policy/utils.py:16-47``)."""


class JobInfo(object):
    """Arguments:
        resources (dict): resources requested by ONE replica (e.g.
            ``{"nvidia.com/gpu": 1, "pods": 1}``).
        speedup_fn: callable ``(num_nodes, num_replicas) -> speedup``
            (vectorised), usually a :class:`SpeedupFunction`.
        creation_timestamp: sortable; earlier jobs win ties (FIFO).
        min_replicas (int): guaranteed replicas when the job runs at all.
        max_replicas (int): upper bound on replicas (> 0, >= min).
        preemptible (bool): may the job be restarted to re-allocate it?
    """

    def __init__(self, resources, speedup_fn, creation_timestamp,
                 min_replicas, max_replicas, preemptible=True):
        assert max_replicas > 0
        assert max_replicas >= min_replicas
        self.resources = resources
        self.speedup_fn = speedup_fn
        self.creation_timestamp = creation_timestamp
        self.min_replicas = min_replicas
        self.max_replicas = max_replicas
        self.preemptible = preemptible


class NodeInfo(object):
    """Arguments:
        resources (dict): resources available on the node.
        preemptible (bool): spot/preemptible node (allocated last).
    """

    def __init__(self, resources, preemptible):
        self.resources = resources
        self.preemptible = preemptible
