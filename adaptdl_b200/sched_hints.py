"""Scheduling-hints schema and client (job -> scheduler telemetry).

Rank 0 periodically PUTs a JSON document to ``{supervisor}/hints/{job}``;
the scheduler stores it in the job's ``status.train`` and derives the job's
speedup function from it (parity: reference ``sched_hints.py:30-59``). The
wire key for the squared gradient norm is ``norm`` (``sqr`` in GradParams).
"""

import json
import logging
from collections import OrderedDict
from types import MappingProxyType

from adaptdl_b200 import env
from adaptdl_b200.goodput import PerfParams

LOG = logging.getLogger(__name__)

PERF_PARAMS = MappingProxyType(
    OrderedDict((k, 0.0) for k in PerfParams._fields))

SCHED_HINTS = MappingProxyType({
    "initBatchSize": 0,
    "localBszBounds": None,        # [min, max]
    "globalBatchSize": None,
    "maxBatchSize": 0,
    "maxProfiledReplicas": 0,
    "gradientAccumulation": False,
    "gradParams": None,            # {"norm": sqr, "var": var}
    "perfParams": None,            # {alpha_c, ..., gamma}
})


def validate_hints(hints):
    unknown = [k for k in hints if k not in SCHED_HINTS]
    if unknown:
        raise ValueError("unknown sched hint keys: {}".format(unknown))


def post_sched_hints(sched_hints, job_key):
    """PUT the hints to the supervisor; silently a no-op without one."""
    import os
    # the single-box launcher (sched/local.py) has no /discover but does
    # accept hints
    url = env.supervisor_url() or os.environ.get("ADAPTDL_HINTS_URL")
    if not url:
        return None
    try:
        validate_hints(sched_hints)
        import requests
        response = requests.put(
            url="{}/hints/{}".format(url, job_key),
            data=json.dumps(sched_hints),
            headers={"Content-Type": "application/json"}, timeout=10)
        if response.status_code != 200:
            LOG.warning("sched hints rejected: HTTP %s", response.status_code)
        return response.status_code
    except Exception as exc:  # noqa: BLE001 - telemetry must never kill a job
        LOG.warning("could not post sched hints: %s", exc)
        return None
