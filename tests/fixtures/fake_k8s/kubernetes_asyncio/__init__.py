"""Stand-in for the slice of ``kubernetes_asyncio`` that
``adaptdl_b200.sched.kube.KubernetesCluster`` calls (the real package cannot
be installed in this image): same call signatures, an in-memory API server
behind them. Objects are returned the way the real client returns them --
model objects with ``to_dict`` for the core API, plain dicts for custom
objects -- so the adapter's ``sanitize_for_serialization`` path is
exercised."""
from . import client, config, watch  # noqa: F401
