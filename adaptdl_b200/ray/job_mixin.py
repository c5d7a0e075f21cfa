"""Mixin giving an elastic Ray job (a Tune trial or the AWS controller's
job) the ``job_info`` the Pollux policy needs, built from the scheduling
hints its workers report (reference: ``adaptdl_ray/adaptdl/
adaptdl_job_mixin.py``)."""

import time

from adaptdl_b200.goodput import GoodputFunction, GradParams, PerfParams
from adaptdl_b200.sched.policy import JobInfo, SpeedupFunction
from adaptdl_b200.sched_hints import PERF_PARAMS


class AdaptDLJobMixin(object):

    def __init__(self, *args, job_id=None, **kwargs):
        self._job_id = job_id
        self.creation_timestamp = time.time()
        self._hints = None
        super().__init__(*args, **kwargs)

    @property
    def job_id(self):
        return self._job_id

    @property
    def hints(self):
        return self._hints

    def update_hints(self, hints):
        self._hints = hints

    def _fetch_metrics(self):
        return self._hints

    def _allocation_in_use(self):
        raise NotImplementedError

    @property
    def job_info(self, resources_per_replica=None, max_replicas=64):
        hints = self._fetch_metrics()
        resources = resources_per_replica or \
            getattr(self, "rescale_resources", {"CPU": 1})
        if hints and hints.get("perfParams"):
            perf = PerfParams(*[hints["perfParams"][k]
                                for k in PERF_PARAMS])
            grad = hints.get("gradParams")
            grad = GradParams(grad["norm"], grad["var"]) if grad \
                else GradParams(0.0, 1.0)
            bounds = hints.get("localBszBounds")
            speedup_fn = SpeedupFunction(
                GoodputFunction(perf, grad, hints["initBatchSize"]),
                hints.get("maxBatchSize"),
                tuple(bounds) if bounds else None,
                hints.get("gradientAccumulation", False))
            max_replicas = min(
                max(2 * hints.get("maxProfiledReplicas", 0), 1),
                max_replicas)
        else:
            def speedup_fn(nodes, replicas):
                return replicas
            max_replicas = 1
        return JobInfo(resources, speedup_fn, self.creation_timestamp, 0,
                       max_replicas)
