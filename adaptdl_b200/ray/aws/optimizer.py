"""Replica-count optimiser for ONE job on an autoscaling Ray cluster
(reference: ``aws/optimizer.py:19-94``): fill nodes in order, keep growing
while the speedup per node stays above half of a single replica's, and only
move if the gain is worth a restart (5 % better and > 15 % different)."""

import numpy as np


def greedy_allocation(existing_nodes, worker_resources, max_cluster_size):
    """``existing_nodes``: ``[(address, resources)]``. Returns one node name
    per possible worker, existing nodes first, then virtual ones."""
    names = [addr for addr, _ in existing_nodes]
    names += ["adaptdl_virtual_node_{}".format(i)
              for i in range(max(max_cluster_size - len(existing_nodes), 0))]
    resources = dict(existing_nodes)
    allocation = []
    for name in names:
        have = resources.get(name, worker_resources)
        fits = int(min(have.get(key, 0.0) / val
                       for key, val in worker_resources.items()))
        allocation += [name] * max(fits, 0)
    return allocation


def optimize(hints, speedup_fn, existing_nodes, worker_resources,
             max_cluster_size, current_replicas):
    if not hints:
        return ["adaptdl_virtual_node_0"]
    allocation = greedy_allocation(existing_nodes, worker_resources,
                                   max_cluster_size)
    if not allocation:
        return []
    workers = np.arange(1, len(allocation) + 1)
    node_counts = np.array([len(set(allocation[:k])) for k in workers])
    speedups = np.asarray(speedup_fn(node_counts, workers), dtype=float)
    base = float(speedup_fn(1, 1))
    best_replicas, best_speedup = 0, 0.0
    for k, speedup in zip(workers, speedups):
        if speedup / node_counts[k - 1] >= 0.5 * base:
            best_replicas, best_speedup = int(k), float(speedup)
    current = min(max(current_replicas, 1), len(allocation))
    current_speedup = float(speedups[current - 1])
    if best_speedup < 1.05 * current_speedup or \
            abs(best_replicas + 1 - current_replicas) < \
            0.15 * current_replicas:
        best_replicas = current_replicas
    return allocation[:best_replicas]
