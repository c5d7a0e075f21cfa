"""Supervisor: the REST endpoint replicas talk to.

* ``GET  /discover/{namespace}/{name}/{group}`` -- long-poll rendezvous:
  replies with the list of pod IPs of restart generation ``group`` (index =
  rank) once every replica has one, else ``408`` after ``?timeout=`` seconds
  (trainer side: ``adaptdl_b200.torch.init_process_group``).
* ``PUT  /hints/{namespace}/{name}`` -- scheduling hints from rank 0, stored
  in the job's ``status.train`` (only recognised keys are kept).
* ``GET  /healthz``
* ``GET  /metrics`` -- Prometheus series derived from the hints (per-job
  batch sizes, gradient statistics, fitted parameters, predicted speedups:
  ``sched/metrics.py``)

Parity: reference ``sched/adaptdl_sched/supervisor.py:27-99``; here over the
backend abstraction, polling the pod list instead of holding a K8s watch.
"""

import asyncio
import logging
import time

from aiohttp import web

from adaptdl_b200.sched import config, metrics
from adaptdl_b200.sched_hints import SCHED_HINTS

LOG = logging.getLogger(__name__)


class Supervisor(object):

    def __init__(self, cluster, port=None, host="0.0.0.0", poll=0.25):
        self._cluster = cluster
        self._host = host
        self._port = port if port is not None else \
            config.get_supervisor_port()
        self._poll = poll
        self.app = web.Application()
        self.app.add_routes([
            web.get("/healthz", self._handle_healthz),
            web.get("/discover/{namespace}/{name}/{group}",
                    self._handle_discover),
            web.put("/hints/{namespace}/{name}", self._handle_report),
            web.get("/metrics", self._handle_metrics),
        ])

    async def _handle_healthz(self, request):
        return web.Response()

    async def _pod_ips(self, namespace, name, group):
        pods = await self._cluster.list_pods(
            namespace, label_selector="adaptdl/job={}".format(name))
        ips = None
        for pod in pods:
            ann = pod["metadata"].get("annotations", {})
            if ann.get("adaptdl/group") != group:
                continue
            replicas, rank = int(ann["adaptdl/replicas"]), \
                int(ann["adaptdl/rank"])
            if ips is None:
                ips = [None] * replicas
            # a node pod (spec.podPerNode) hosts a range of ranks
            hosted = int(ann.get("adaptdl/local-replicas", 1))
            for local in range(hosted):
                ips[rank + local] = (pod.get("status") or {}).get("podIP")
        return ips

    async def _handle_discover(self, request):
        info = request.match_info
        timeout = float(request.query.get("timeout", "30"))
        deadline = time.monotonic() + timeout
        while True:
            ips = await self._pod_ips(info["namespace"], info["name"],
                                      info["group"])
            if ips and all(ip for ip in ips):
                return web.json_response(ips)
            if time.monotonic() >= deadline:
                return web.json_response(None, status=408)
            await asyncio.sleep(self._poll)

    async def _handle_report(self, request):
        info = request.match_info
        hints = await request.json()
        hints = {k: hints[k] for k in SCHED_HINTS if k in hints}
        LOG.info("hints for %s/%s: %s", info["namespace"], info["name"],
                 hints)
        patched = await self._cluster.patch_job_status(
            info["namespace"], info["name"], {"status": {"train": hints}})
        if patched is None:
            metrics.forget_job(info["namespace"], info["name"])
            return web.Response(status=404)
        metrics.observe_hints(info["namespace"], info["name"], hints)
        return web.Response()

    async def _handle_metrics(self, request):
        """Prometheus exposition of the per-job series (sched/metrics.py)."""
        body, content_type = metrics.render()
        return web.Response(body=body,
                            headers={"Content-Type": content_type})

    def run(self):
        web.run_app(self.app, host=self._host, port=self._port)


if __name__ == "__main__":      # ``python -m adaptdl_sched.supervisor``, as in
    from adaptdl_b200.sched.__main__ import main  # the reference's chart
    main(["supervisor"])
