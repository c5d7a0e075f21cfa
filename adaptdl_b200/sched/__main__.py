"""``python -m adaptdl_b200.sched [controller|allocator|supervisor|validator]``

The three long-running scheduler containers of the helm chart (controller,
allocator, supervisor) plus the admission webhook. Needs the optional
``kubernetes_asyncio`` package and in-cluster credentials."""

import asyncio
import logging
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    role = argv.pop(0) if argv else "controller"
    logging.basicConfig(level=logging.INFO)
    if role == "validator":
        from adaptdl_b200.sched import validator
        return validator.main(argv)
    import kubernetes_asyncio as kubernetes
    kubernetes.config.load_incluster_config()
    from adaptdl_b200.sched.kube import KubernetesCluster
    cluster = KubernetesCluster()
    if role == "controller":
        try:
            from prometheus_client import start_http_server
            start_http_server(9091)
        except Exception:  # noqa: BLE001
            pass
        from adaptdl_b200.sched.controller import AdaptDLController
        asyncio.run(AdaptDLController(cluster).run())
    elif role == "allocator":
        from adaptdl_b200.sched.allocator import AdaptDLAllocator
        from adaptdl_b200.sched.cluster_expander import ClusterExpander
        from adaptdl_b200.sched import metrics
        metrics.serve(9092)              # per-cycle and per-job series
        expander = ClusterExpander(cluster)
        allocator = AdaptDLAllocator(cluster, expander)

        async def both():
            await asyncio.gather(expander.run(), allocator.run())
        asyncio.run(both())
    elif role == "supervisor":
        from adaptdl_b200.sched.supervisor import Supervisor
        Supervisor(cluster).run()
    else:
        raise SystemExit("unknown role: " + role)


if __name__ == "__main__":
    main()
