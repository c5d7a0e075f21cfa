"""Step profiler, performance-model fitting and scheduler hints.

The profile is a table keyed by ``(num_nodes, num_replicas, atomic_bsz)`` with
counters ``accum_step_time / accum_count / optim_step_time / optim_sync_time /
optim_count`` (parity: reference ``torch/_metrics.py:29-199``). It survives
restarts (so measurements at several replica counts accumulate) and feeds
:func:`adaptdl_b200.goodput.fit_perf_params`.

B200 difference: step and sync durations may be supplied by the on-device
timer (``%globaltimer`` stamps written by the fused reducer; see
``parallel/reducer_cuda.py``) through :func:`profile_step_commit`'s
``step_time`` argument instead of host wall-clock, which removes the
blocking ``event.synchronize()`` the reference needs every step.
"""

import collections
import pickle
import os
import time

import numpy as np

from adaptdl_b200 import checkpoint, env
from adaptdl_b200.goodput import GoodputFunction, fit_perf_params
from adaptdl_b200.sched_hints import SCHED_HINTS, PERF_PARAMS, \
    post_sched_hints

# seconds between performance fits / scheduling-hint reports (reference: 30 s,
# torch/_metrics.py:63); overridable for tests and short jobs
REPORT_PERIOD_S = float(os.environ.get("ADAPTDL_REPORT_PERIOD", "30"))


class _MetricsState(checkpoint.State):
    def __init__(self):
        super().__init__("adaptdl-metrics")
        self.profile = collections.defaultdict(collections.Counter)
        self.perf_params = None
        self.grad_params = None
        self.init_batch_size = None
        self.max_batch_size = None
        self.local_bsz_bounds = None
        self.gradient_accumulation = False
        self.progress = 0.0      # scale-invariant iterations completed

    _FIELDS = ("profile", "perf_params", "grad_params", "init_batch_size",
               "max_batch_size", "local_bsz_bounds", "gradient_accumulation",
               "progress")

    def save(self, fileobj):     # 8 consecutive pickles (App. B)
        for name in self._FIELDS:
            pickle.dump(getattr(self, name), fileobj)

    def load(self, fileobj):
        for name in self._FIELDS:
            setattr(self, name, pickle.load(fileobj))


_METRICS_STATE = None
_PREV_REPORT = None
_GRAD_PARAM_DICT = {}


def _metrics_state():
    global _METRICS_STATE
    if _METRICS_STATE is None:
        _METRICS_STATE = _MetricsState()
        checkpoint.load_state(_METRICS_STATE)
    return _METRICS_STATE


def profile_step_start(atomic_bsz):
    state = _metrics_state()
    state.atomic_bsz = atomic_bsz
    state.step_start = time.time()
    state.sync_time = 0.0


def profile_sync_time(sync_time):
    _metrics_state().sync_time += sync_time


def profile_step_commit(accumulation_step=False, step_time=None):
    """Commit the measurements of the step opened by
    :func:`profile_step_start`. ``step_time`` overrides the host wall-clock
    (device-timed steps)."""
    global _PREV_REPORT
    state = _metrics_state()
    if step_time is None:
        step_time = time.time() - state.step_start
    key = (env.num_nodes(), env.num_replicas(), state.atomic_bsz)
    row = state.profile[key]
    if accumulation_step:
        row["accum_step_time"] += step_time
        row["accum_count"] += 1
    else:
        row["optim_step_time"] += step_time
        row["optim_sync_time"] += state.sync_time
        row["optim_count"] += 1
    del state.atomic_bsz, state.step_start, state.sync_time
    if not accumulation_step:
        now = time.time()
        if _PREV_REPORT is None:
            _PREV_REPORT = now
        if env.replica_rank() == 0 and now - _PREV_REPORT > REPORT_PERIOD_S:
            _fit_perf_params()
            _report_sched_hints()
            _PREV_REPORT = time.time()


def update_grad_params(edp_key, grad_norm_sqr, grad_variance):
    """Record one data-parallel instance's (sqr, var); instances are summed
    (e.g. GAN generator + discriminator)."""
    _GRAD_PARAM_DICT[edp_key] = np.asarray([grad_norm_sqr, grad_variance],
                                           dtype=float)
    total = sum(_GRAD_PARAM_DICT.values())
    _metrics_state().grad_params = (float(total[0]), float(total[1]))


def update_progress(progress):
    _metrics_state().progress = progress


def get_progress():
    return _metrics_state().progress


def set_batch_size(init_batch_size, max_batch_size, local_bsz_bounds,
                   gradient_accumulation):
    state = _metrics_state()
    state.init_batch_size = init_batch_size
    state.max_batch_size = max_batch_size
    state.local_bsz_bounds = local_bsz_bounds
    state.gradient_accumulation = gradient_accumulation


def get_goodput_fn():
    state = _metrics_state()
    if state.grad_params is None or state.perf_params is None:
        return None
    return GoodputFunction(state.perf_params, state.grad_params,
                           state.init_batch_size)


def _fit_perf_params():
    state = _metrics_state()
    profile = {k: v for k, v in state.profile.items() if v.get("optim_count")}
    if not profile:
        return
    num_nodes, num_replicas, atomic_bsz = (
        np.array(col) for col in zip(*profile.keys()))
    rows = list(profile.values())

    def column(name, dtype=float):
        return np.array([row.get(name, 0) for row in rows], dtype=dtype)

    accum_step_time = column("accum_step_time")
    accum_count = column("accum_count")
    optim_step_time = column("optim_step_time")
    optim_sync_time = column("optim_sync_time")
    optim_count = column("optim_count")
    assert np.all(optim_count > 0)
    # the model requires step time >= sync time (device-timed sync can
    # exceed a wall-clock step by jitter; clamp instead of asserting)
    optim_sync_time = np.minimum(optim_sync_time, optim_step_time)
    # The non-sync part of an optimisation step costs about as much as an
    # accumulation step; pool the two kinds of sample.
    accum_step_time = (accum_step_time + optim_step_time - optim_sync_time) \
        / (accum_count + optim_count)
    optim_step_time = optim_step_time / optim_count
    state.perf_params = fit_perf_params(num_nodes, num_replicas, atomic_bsz,
                                        accum_step_time, optim_step_time)


def _get_sched_hints():
    state = _metrics_state()
    if len(state.profile) == 0:
        return None
    _fit_perf_params()
    return _metrics_state()


def _build_sched_hints():
    state = _metrics_state()
    hints = dict(SCHED_HINTS)
    if state.perf_params is not None:
        hints["perfParams"] = {k: float(v) for k, v in
                               zip(PERF_PARAMS.keys(), state.perf_params)}
    hints["maxBatchSize"] = state.max_batch_size
    hints["localBszBounds"] = state.local_bsz_bounds
    hints["initBatchSize"] = state.init_batch_size
    if state.grad_params:
        hints["gradParams"] = {"norm": float(state.grad_params[0]),
                               "var": float(state.grad_params[1])}
    hints["maxProfiledReplicas"] = max(key[1] for key in state.profile)
    hints["gradientAccumulation"] = state.gradient_accumulation
    return hints


def _report_sched_hints():
    assert env.replica_rank() == 0
    post_sched_hints(_build_sched_hints(), env.job_id())


def _reset_for_tests():
    global _METRICS_STATE, _PREV_REPORT
    _METRICS_STATE = None
    _PREV_REPORT = None
    _GRAD_PARAM_DICT.clear()
