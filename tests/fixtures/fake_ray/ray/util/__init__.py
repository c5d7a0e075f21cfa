from .._actors import get_node_ip_address  # noqa: F401


def get_current_placement_group():
    """Not inside a placement group (the driver process of a test)."""
    return None


def placement_group_table(group):
    return {"bundles": {}, "bundles_to_node_id": {}}
