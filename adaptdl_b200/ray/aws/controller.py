"""Ray-on-AWS controller: supervisor + scheduler for ONE elastic job
(reference: ``aws/controller.py:52-455``).

The controller is a Ray actor that
* serves the supervisor REST endpoints (``/discover``, ``/hints``) so the
  unmodified trainer API works (``ADAPTDL_SUPERVISOR_URL``),
* keeps the job's worker tasks running on the current allocation,
* re-optimises the replica count whenever new hints arrive (at most every
  5 minutes) or immediately when a node receives a spot-termination notice,
* asks the Ray autoscaler for nodes and waits for them (bounded), and
* carries the checkpoint between generations through the object store.

``cluster_ready`` / ``trim_allocation`` are pure and unit-tested; the actor
itself needs Ray.
"""

import asyncio
import copy
import logging
import os
import time
import uuid

from adaptdl_b200.goodput import GoodputFunction, GradParams, PerfParams
from adaptdl_b200.ray.aws import optimizer
from adaptdl_b200.ray.aws.utils import Status
from adaptdl_b200.sched.policy import SpeedupFunction
from adaptdl_b200.sched_hints import PERF_PARAMS, SCHED_HINTS

LOG = logging.getLogger(__name__)
MIN_RESCHEDULE_PERIOD_S = 300
FULL_RESCALE_TIMEOUT_S = 1200
WORKER_FAILURE_BACKOFF_S = 60


def cluster_ready(allocation, nodes, worker_resources):
    """Can ``allocation`` (node addresses, virtual names for nodes yet to
    come) be placed on ``nodes`` (``{address: resources}``)? Returns
    ``(ready, workers_placeable)``."""
    free = {addr: dict(res) for addr, res in nodes.items()}
    placed, floating = 0, 0
    for node in allocation:
        if "adaptdl_virtual" in node or node not in free:
            floating += 1
            continue
        for key, amount in worker_resources.items():
            free[node][key] = free[node].get(key, 0.0) - amount
        placed += 1
    for _ in range(floating):
        for res in free.values():
            if all(res.get(key, 0.0) >= amount
                   for key, amount in worker_resources.items()):
                for key, amount in worker_resources.items():
                    res[key] -= amount
                placed += 1
                break
        else:
            return False, placed
    return placed >= len(allocation), placed


def trim_allocation(allocation, placeable):
    """Keep real nodes first when only ``placeable`` workers fit."""
    real = [n for n in allocation if "adaptdl_virtual" not in n]
    virtual = [n for n in allocation if "adaptdl_virtual" in n]
    return (real + virtual)[:placeable]


def speedup_from_hints(hints):
    perf = PerfParams(*[hints["perfParams"][k] for k in PERF_PARAMS])
    grad = hints.get("gradParams")
    grad = GradParams(grad["norm"], grad["var"]) if grad \
        else GradParams(1.0, 1.0)
    bounds = hints.get("localBszBounds")
    return SpeedupFunction(
        GoodputFunction(perf, grad, hints["initBatchSize"]),
        hints.get("maxBatchSize"), tuple(bounds) if bounds else None,
        hints.get("gradientAccumulation", False))


class JobState(object):
    """Bookkeeping of the single job (no Ray calls)."""

    def __init__(self, worker_resources, checkpoint_timeout, job_params):
        self.uid = uuid.uuid4().hex[:8]
        self.worker_resources = dict(worker_resources)
        self.checkpoint_timeout = checkpoint_timeout
        self.params = dict(job_params)
        self.workers = {}            # rank -> node address
        self.tasks = {}
        self.hints = None
        self.checkpoint = None
        self.iteration = 0
        self.status = Status.RUNNING
        self.running = False

    def register_hints(self, hints):
        self.hints = {k: copy.deepcopy(hints[k]) for k in SCHED_HINTS
                      if k in hints}

    def register_status(self, status):
        if self.status is not Status.SUCCEEDED:
            self.status = Status(status)


def make_controller_actor():
    """Build the ``Controller`` Ray actor class (needs Ray + aiohttp)."""
    from aiohttp import web
    from adaptdl_b200.ray import require_ray
    from adaptdl_b200.ray.aws.worker import remote_functions
    ray = require_ray()
    from ray.autoscaler import sdk
    listen_for_spot_termination, run_adaptdl = remote_functions()

    @ray.remote(num_cpus=1)
    class Controller(object):
        def __init__(self, cluster_size, rescale_timeout=120):
            self._cluster_size = cluster_size
            self._rescale_timeout = rescale_timeout
            # the reference's fixed port; overridable where 8080 is taken
            self._port = int(os.environ.get(
                "ADAPTDL_B200_RAY_CONTROLLER_PORT", "8080"))
            self._url = "http://{}:{}".format(
                os.environ.get("ADAPTDL_B200_RAY_CONTROLLER_HOST")
                or ray.util.get_node_ip_address(), self._port)
            self._job = None
            self._terminating = set()
            self._force = asyncio.Event()
            self._ready = asyncio.Event()
            self._completed = asyncio.Event()
            self._ckpt_received = asyncio.Event()
            self._queue = asyncio.Queue(maxsize=1)
            self._last = 0.0
            self._spot_tasks = {}

        def get_url(self):
            return self._url

        # -- REST (supervisor role) ----------------------------------------
        async def _run_app(self):
            app = web.Application()
            app.add_routes([
                web.get("/discover/{ns}/{name}/{group}", self._discover),
                web.put("/hints/{ns}/{name}", self._hints)])
            self._runner = web.AppRunner(app)
            await self._runner.setup()
            await web.TCPSite(self._runner, "0.0.0.0", self._port).start()
            self._ready.set()

        async def _discover(self, request):
            # a worker registers itself with a fire-and-forget call just
            # before its script starts, so the first replicas can ask before
            # the last one is known: hold the answer until the generation is
            # complete (408 = "ask again", what the trainer's client expects)
            job = self._job
            deadline = time.monotonic() + float(
                request.query.get("timeout", "30"))
            while not job.tasks or len(job.workers) < len(job.tasks):
                if time.monotonic() >= deadline:
                    return web.json_response(None, status=408)
                await asyncio.sleep(0.1)
            ips = [ip for _, ip in sorted(job.workers.items())]
            return web.json_response(ips)

        async def _hints(self, request):
            self._job.register_hints(await request.json())
            await self._enqueue(False)
            return web.Response(text="ok")

        # -- lifecycle -------------------------------------------------------------
        async def run_controller(self):
            asyncio.ensure_future(self._run_app())
            asyncio.ensure_future(self._listener())
            await self._completed.wait()
            await self._runner.cleanup()

        async def create_job(self, worker_resources, worker_port_offset=0,
                             checkpoint_timeout=120, **job_params):
            await self._ready.wait()
            self._job = JobState(worker_resources, checkpoint_timeout,
                                 dict(job_params, offset=worker_port_offset))
            await self._enqueue(True)
            while self._job.status is Status.RUNNING:
                await asyncio.sleep(1.0)
            self._completed.set()
            return self._job.status.value

        async def _enqueue(self, immediate):
            if immediate:
                while not self._queue.empty():
                    self._queue.get_nowait()
                await self._queue.put(True)
            else:
                try:
                    self._queue.put_nowait(False)
                except asyncio.QueueFull:
                    pass

        async def _listener(self):
            while True:
                immediate = await self._queue.get()
                wait = MIN_RESCHEDULE_PERIOD_S - (time.time() - self._last)
                if not immediate and wait > 0:
                    await asyncio.sleep(wait)
                await self._reschedule()
                self._last = time.time()

        def _nodes(self):
            me = ray.util.get_node_ip_address()
            return {n["NodeManagerAddress"]: dict(n["Resources"])
                    for n in ray.nodes()
                    if n.get("alive", n.get("Alive")) and "Resources" in n
                    and n["NodeManagerAddress"] not in self._terminating
                    and n["NodeManagerAddress"] != me}

        async def _reschedule(self):
            job = self._job
            nodes = self._nodes()
            fn = speedup_from_hints(job.hints) if job.hints and \
                job.hints.get("perfParams") else None
            allocation = optimizer.optimize(
                job.hints if fn else None, fn, list(nodes.items()),
                job.worker_resources, self._cluster_size,
                max(len(job.tasks), 1))
            allocation = await self._expand(allocation)
            await self._update_workers(allocation)
            self._force.clear()

        async def _expand(self, allocation):
            job = self._job
            alive = set(self._nodes())
            lost = [r for r, ip in job.workers.items() if ip not in alive]
            timeout = FULL_RESCALE_TIMEOUT_S \
                if len(lost) == len(job.workers) else self._rescale_timeout
            bundles = [dict(job.worker_resources,
                            CPU=job.worker_resources.get("CPU", 1) + 0.1)
                       for _ in range(len(allocation) + len(lost))]
            sdk.request_resources(bundles=bundles)
            waited = 0.0
            while waited < timeout and not self._force.is_set() and not \
                    cluster_ready(allocation, self._nodes(),
                                  job.worker_resources)[0]:
                await asyncio.sleep(1.0)
                waited += 1.0
            ready, count = cluster_ready(allocation, self._nodes(),
                                         job.worker_resources)
            return allocation if ready else trim_allocation(allocation,
                                                            count)

        async def _update_workers(self, allocation):
            job = self._job
            if set(job.workers.values()) == set(allocation) and job.tasks:
                return
            if job.running:               # ask for a checkpoint, then stop
                for task in job.tasks.values():
                    ray.cancel(task, force=False)
                job.running = False
                try:
                    await asyncio.wait_for(self._ckpt_received.wait(),
                                           job.checkpoint_timeout)
                except asyncio.TimeoutError:
                    LOG.warning("no checkpoint in time; reusing the "
                                "previous one")
            for task in job.tasks.values():
                try:
                    ray.cancel(task, force=True)
                except Exception:  # noqa: BLE001
                    pass
            job.tasks, job.workers = {}, {}
            ckpt_ref = ray.put(job.checkpoint) if job.checkpoint else None
            for rank, node in enumerate(allocation):
                options = dict(
                    num_cpus=job.worker_resources.get("CPU", 1),
                    num_gpus=job.worker_resources.get("GPU", 0))
                if "adaptdl_virtual" not in node:
                    options["resources"] = {"node:{}".format(node): 0.01}
                job.tasks[rank] = run_adaptdl.options(**options).remote(
                    "default/job", job.uid, rank, len(allocation),
                    job.iteration, ckpt_ref, **job.params)
            self._ckpt_received.clear()
            job.running = True
            job.iteration += 1
            asyncio.ensure_future(self._watch(list(job.tasks.values())))

        async def _watch(self, tasks):
            try:
                await asyncio.gather(*tasks)
            except Exception as exc:  # noqa: BLE001
                if self._job.workers:
                    LOG.error("worker failure: %s", exc)
                    await asyncio.sleep(WORKER_FAILURE_BACKOFF_S)
                    await self._enqueue(True)

        # -- callbacks from workers -----------------------------------------------
        async def register_worker(self, rank, ip):
            self._job.workers[rank] = ip
            if ip not in self._spot_tasks:
                self._spot_tasks[ip] = listen_for_spot_termination.options(
                    num_cpus=0.1,
                    resources={"node:{}".format(ip): 0.01}).remote()
                asyncio.ensure_future(self._on_spot(self._spot_tasks[ip]))

        async def _on_spot(self, task):
            try:
                ip = await task
            except Exception:  # noqa: BLE001
                return
            if ip:
                self._terminating.add(ip)
                self._force.set()
                await self._enqueue(True)

        async def register_checkpoint(self, checkpoint):
            self._job.checkpoint = checkpoint
            self._ckpt_received.set()
            return True

        async def register_status(self, status):
            self._job.register_status(status)

    return Controller
