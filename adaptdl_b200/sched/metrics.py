"""Per-job and per-cycle Prometheus series of the scheduler.

The reference ships a Grafana dashboard over ``job_replicas``,
``job_perf_predict{replicas=..}`` ... that nothing in its code exports any more
(SURVEY.md S12). Here the series exist: the supervisor derives them from the
scheduling hints every trainer reports (``PUT /hints``), the allocator from the
outcome of each optimisation cycle, and ``deploy/grafana/dashboard.json`` plots
them.

=============================================  ===============================
``adaptdl_job_replicas{namespace,job}``        replicas in the job's allocation
``adaptdl_job_nodes{namespace,job}``           distinct nodes it spans
``adaptdl_job_batch_size{..,kind}``            ``init`` / ``max`` batch size
``adaptdl_job_max_profiled_replicas{..}``      largest replica count profiled
``adaptdl_job_grad_sqr`` / ``_grad_var``       gradient statistics (GNS = var/sqr)
``adaptdl_job_perf_param{..,param}``           the seven fitted parameters
``adaptdl_job_speedup_predict{..,replicas}``   modelled speedup on 1 node
``adaptdl_sched_cycle_seconds``                duration of an allocation cycle
``adaptdl_sched_cluster_gpus{state}``          ``total`` / ``allocated`` GPUs
``adaptdl_sched_desired_nodes``                nodes the policy asked for
=============================================  ===============================

``prometheus_client`` is optional: without it every function is a no-op and
:func:`render` returns an empty exposition.
"""

import logging

import numpy as np

from adaptdl_b200.goodput import GoodputFunction, GradParams, PerfParams
from adaptdl_b200.sched_hints import PERF_PARAMS

LOG = logging.getLogger(__name__)

PREDICT_REPLICAS = (1, 2, 4, 8, 16, 32)
GPU_RESOURCE = "nvidia.com/gpu"

try:
    import prometheus_client as _prom
except Exception:  # noqa: BLE001
    _prom = None

_JOB = ["namespace", "job"]
_GAUGES = {}


def _gauge(name, doc, labels=()):
    if _prom is None:
        return None
    if name not in _GAUGES:
        _GAUGES[name] = _prom.Gauge(name, doc, labelnames=list(labels))
    return _GAUGES[name]


def _set(name, doc, value, **labels):
    gauge = _gauge(name, doc, labels)
    if gauge is None or value is None:
        return
    (gauge.labels(**labels) if labels else gauge).set(float(value))


def speedup_predictions(hints, replicas=PREDICT_REPLICAS):
    """``{replicas: speedup}`` on a single node from a hints dict, or ``{}``
    while the job has not reported both a performance fit and gradient
    statistics."""
    perf, grad = hints.get("perfParams"), hints.get("gradParams")
    init_bsz = hints.get("initBatchSize")
    if not (perf and grad and init_bsz):
        return {}
    try:
        fn = GoodputFunction(
            PerfParams(*[perf[k] for k in PERF_PARAMS]),
            GradParams(grad["norm"], grad["var"]), init_bsz)
        search = dict(
            max_batch_size=hints.get("maxBatchSize") or init_bsz,
            atomic_bsz_range=tuple(hints["localBszBounds"])
            if hints.get("localBszBounds") else None,
            accumulation=bool(hints.get("gradientAccumulation")))
        counts = np.asarray(replicas)
        best, _, _ = fn.optimize(np.ones_like(counts), counts, **search)
        base = float(np.atleast_1d(best)[0])
        if not base > 0:
            return {}
        return {int(r): float(g) / base for r, g in zip(counts, best)}
    except Exception:  # noqa: BLE001
        LOG.debug("speedup prediction failed", exc_info=True)
        return {}


def observe_hints(namespace, name, hints):
    """Supervisor side: publish what a trainer just reported."""
    job = dict(namespace=namespace, job=name)
    for kind, key in (("init", "initBatchSize"), ("max", "maxBatchSize")):
        _set("adaptdl_job_batch_size", "Batch size configuration",
             hints.get(key), kind=kind, **job)
    _set("adaptdl_job_max_profiled_replicas",
         "Largest replica count the job has profiled",
         hints.get("maxProfiledReplicas"), **job)
    grad = hints.get("gradParams") or {}
    _set("adaptdl_job_grad_sqr", "Squared norm of the true gradient",
         grad.get("norm"), **job)
    _set("adaptdl_job_grad_var", "Trace of the gradient covariance",
         grad.get("var"), **job)
    for param, value in (hints.get("perfParams") or {}).items():
        _set("adaptdl_job_perf_param", "Fitted performance model parameter",
             value, param=param, **job)
    for replicas, speedup in speedup_predictions(hints).items():
        _set("adaptdl_job_speedup_predict",
             "Modelled speedup over one replica (single node)", speedup,
             replicas=str(replicas), **job)


def observe_cycle(allocations, nodes, seconds=None, desired_nodes=None,
                  known_jobs=()):
    """Allocator side: publish the outcome of an optimisation cycle.
    ``allocations``: ``{(namespace, name): [node, ...]}``; ``nodes``:
    ``{name: NodeInfo}`` (free resources net of non-AdaptDL pods)."""
    keys = set(known_jobs) | set(allocations)
    for key in keys:
        alloc = allocations.get(key) or []
        job = dict(namespace=key[0], job=key[1])
        _set("adaptdl_job_replicas", "Replicas allocated to the job",
             len(alloc), **job)
        _set("adaptdl_job_nodes", "Nodes the job's allocation spans",
             len(set(alloc)), **job)
    total = sum(info.resources.get(GPU_RESOURCE, 0) for info in nodes.values())
    used = sum(len(alloc or []) for alloc in allocations.values())
    _set("adaptdl_sched_cluster_gpus", "GPUs seen by the allocator", total,
         state="total")
    _set("adaptdl_sched_cluster_gpus", "GPUs seen by the allocator", used,
         state="allocated")
    _set("adaptdl_sched_cycle_seconds", "Duration of the last allocation "
         "cycle", seconds)
    _set("adaptdl_sched_desired_nodes", "Nodes requested from the cluster "
         "autoscaler", desired_nodes)


def forget_job(namespace, name):
    """Drop every series of a finished job (keeps cardinality bounded)."""
    for gauge in _GAUGES.values():
        names = list(gauge._labelnames)
        if names[:2] != _JOB:
            continue
        for labels in list(gauge._metrics):
            if labels[:2] == (namespace, name):
                gauge.remove(*labels)


def render():
    """``(body, content_type)`` of the Prometheus exposition."""
    if _prom is None:
        return b"", "text/plain"
    return _prom.generate_latest(), _prom.CONTENT_TYPE_LATEST


def serve(port):
    """Serve the exposition from a background thread (allocator container)."""
    if _prom is None:
        return False
    try:
        _prom.start_http_server(port)
    except OSError as exc:
        # observability must not take the allocator down with it
        LOG.warning("metrics endpoint not started on port %s: %s", port,
                    exc)
        return False
    return True
