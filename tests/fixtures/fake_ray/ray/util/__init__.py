from .._actors import get_node_ip_address  # noqa: F401
