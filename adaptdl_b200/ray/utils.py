"""Allocation <-> Ray placement-group conversions (reference:
``adaptdl_ray/adaptdl/utils.py``). Pure functions take bundles as plain
lists of dicts; the Ray-typed wrappers import Ray lazily."""

import collections


def allocation_to_bundles(allocation, resources_per_replica):
    """``["nodeA", "nodeA", "nodeB"]`` -> one bundle per replica pinned to
    its node with Ray's ``node:<ip>`` custom resource."""
    bundles = []
    for node in allocation:
        bundle = dict(resources_per_replica)
        if not str(node).startswith("virtual"):
            bundle["node:{}".format(node)] = 0.01
        bundles.append(bundle)
    return bundles


def bundles_to_allocation(bundles):
    """Inverse of :func:`allocation_to_bundles`."""
    allocation = []
    for i, bundle in enumerate(bundles):
        node = next((k[len("node:"):] for k in bundle
                     if k.startswith("node:")), None)
        allocation.append(node if node is not None
                          else "virtual-{}".format(i))
    return allocation


def unique_nodes(bundles):
    """Number of distinct nodes a set of bundles spans."""
    return max(len(set(bundles_to_allocation(bundles))), 1)


def allocation_to_pgf(allocation, resources_per_replica=None):
    from adaptdl_b200.ray import require_ray
    require_ray()
    from ray.tune import PlacementGroupFactory
    from adaptdl_b200.ray.config import default_device
    resources = resources_per_replica or {"CPU": 1, default_device(): 1}
    return PlacementGroupFactory(
        [{"CPU": 0.001}] + allocation_to_bundles(allocation, resources))


def pgf_to_allocation(pgf):
    return bundles_to_allocation(list(pgf.bundles)[1:])


def pgf_to_num_replicas(pgf):
    return len(pgf.bundles) - 1


def unique_nodes_pg():
    """Distinct nodes of the current placement group (called from inside a
    Tune worker by ``init_process_group``)."""
    from adaptdl_b200.ray import require_ray
    ray = require_ray()
    pg = ray.util.get_current_placement_group()
    if pg is None:
        return 1
    table = ray.util.placement_group_table(pg)
    nodes = collections.Counter(table.get("bundles_to_node_id", {}).values())
    return max(len(nodes), 1)
