"""Cluster expander: makes the Kubernetes cluster-autoscaler provision the
number of nodes the allocator wants.

It keeps exactly ``len(active_nodes)`` tiny *placeholder* pods alive; they
repel each other (required pod anti-affinity on the hostname topology), so
each needs its own node and every Pending placeholder makes the autoscaler
add one. When fewer nodes are wanted, placeholders are removed -- those on
nodes that host no allocation go first -- and the autoscaler scales down
(parity: reference ``sched/adaptdl_sched/cluster_expander.py:28-188``).
"""

import asyncio
import logging

from adaptdl_b200.sched import config, k8s_templates as templates

LOG = logging.getLogger(__name__)
PERIOD_S = 30.0


def placeholder_pod(owner_reference=None):
    selector = {"matchExpressions": [{"key": config.ADAPTDL_PH_LABEL,
                                      "operator": "In", "values": ["true"]}]}
    return {
        "apiVersion": "v1", "kind": "Pod",
        "metadata": {
            "generateName": "adaptdl-placeholder-",
            "labels": {config.ADAPTDL_PH_LABEL: "true",
                       "petuum.com/nodegroup": "adaptdl"},
            "ownerReferences": owner_reference or [],
        },
        "spec": {
            "affinity": {"podAntiAffinity": {
                "requiredDuringSchedulingIgnoredDuringExecution": [{
                    "labelSelector": selector,
                    "topologyKey": "kubernetes.io/hostname"}]}},
            "containers": [{
                "name": "placeholder", "image": "busybox",
                "command": ["/bin/sh", "-ec",
                            "while :; do echo '.'; sleep 5 ; done"],
                "resources": {"requests": {"memory": "5Mi", "cpu": "1m"}},
            }],
            "restartPolicy": "Never",
        },
    }


class ClusterExpander(object):

    def __init__(self, cluster, namespace=None, owner_reference=None):
        self._cluster = cluster
        self._namespace = namespace or config.get_namespace()
        self._owner_reference = owner_reference
        self._active_nodes = set()
        self._allocations = set()

    def fit(self, active_nodes):
        """``active_nodes``: names of the nodes to keep plus one ``"~k"``
        entry per additional node wanted."""
        self._active_nodes = set(active_nodes)
        self._allocations = {n for n in self._active_nodes
                             if not str(n).startswith("~")}

    @property
    def expected(self):
        return len(self._active_nodes)

    async def reconcile(self):
        selector = "{}=true".format(config.ADAPTDL_PH_LABEL)
        pods = await self._cluster.list_pods(self._namespace,
                                             label_selector=selector)

        def keep_priority(pod):
            running = (pod.get("status") or {}).get("phase") == "Running"
            on_allocated = (pod.get("spec") or {}).get("nodeName") \
                in self._allocations
            return 2 if (running and on_allocated) else 1 if running else 0
        pods.sort(key=keep_priority)
        live = [p for p in pods
                if (p.get("status") or {}).get("phase") in ("Running",
                                                             "Pending")]
        expected = self.expected
        LOG.info("placeholders: want %d, have %d", expected, len(live))
        if expected > len(live):
            for _ in range(expected - len(live)):
                # no explicit name: the API server derives a unique one
                # from the template's generateName, so placeholders that
                # survived a scheduler restart can never collide (409) with
                # new ones
                pod = placeholder_pod(self._owner_reference)
                pod["metadata"].pop("name", None)
                await self._cluster.create_pod(self._namespace, pod)
        elif expected < len(live):
            for pod in live[:len(live) - expected]:
                await self._cluster.delete_pod(self._namespace,
                                               pod["metadata"]["name"])

    async def run(self):
        while True:
            try:
                await self.reconcile()
            except Exception:  # noqa: BLE001
                LOG.exception("placeholder reconciliation failed")
            await asyncio.sleep(PERIOD_S)


_ = templates
