"""Step profiler, performance-model fitting and scheduler hints.

Three small pieces:

``_MetricsState``
    the checkpointed record: a profile table keyed by ``(num_nodes,
    num_replicas, atomic_bsz)`` whose rows count ``accum_step_time /
    accum_count / optim_step_time / optim_sync_time / optim_count``, the
    fitted performance parameters, the latest gradient statistics, the batch
    size configuration and the job's scale-invariant progress. It survives
    restarts, so measurements taken at different replica counts accumulate
    (on-disk layout: Appendix B of SURVEY.md, eight consecutive pickles).
``_OpenStep``
    the measurements of the iteration in flight (not checkpointed).
module functions
    the semi-public API the data loader, the data-parallel wrapper and the
    Ray integration call (same names as the reference's
    ``torch/_metrics.py:29-199``).

B200 difference: step and sync durations may come from the on-device timer
(``%globaltimer`` stamps written by the fused reducer, see
``parallel/reducer_cuda.py``) via ``profile_step_commit(step_time=...)``
instead of the host clock, which removes the blocking event synchronisation
the reference pays every step.
"""

import collections
import logging
import os
import threading
import time

import numpy as np

from adaptdl_b200 import checkpoint, env
from adaptdl_b200.goodput import GoodputFunction, fit_perf_params
from adaptdl_b200.sched_hints import SCHED_HINTS, PERF_PARAMS, \
    post_sched_hints

# seconds between performance fits / scheduling-hint reports (reference: 30 s,
# torch/_metrics.py:63); overridable for tests and short jobs
REPORT_PERIOD_S = float(os.environ.get("ADAPTDL_REPORT_PERIOD", "30"))
# The periodic fit (an L-BFGS solve, ~0.2 s) and the hints PUT (up to its 10 s
# timeout against a slow supervisor) run on a background thread so that
# telemetry never stalls the step loop; ADAPTDL_ASYNC_REPORT=0 restores the
# reference's inline behaviour.
ASYNC_REPORT = os.environ.get("ADAPTDL_ASYNC_REPORT", "1") != "0"

LOG = logging.getLogger(__name__)


class _MetricsState(checkpoint.PickledFields):
    # the order of the eight pickles on disk
    FIELDS = ("profile", "perf_params", "grad_params", "init_batch_size",
              "max_batch_size", "local_bsz_bounds", "gradient_accumulation",
              "progress")
    LAYOUT = "sequence"

    def __init__(self):
        super().__init__("adaptdl-metrics")
        for name in self.FIELDS:
            setattr(self, name, None)
        self.profile = collections.defaultdict(collections.Counter)
        self.gradient_accumulation = False
        self.progress = 0.0


    def sync(self):
        # before a checkpoint: book every step the device has timed
        _book_device_records(wait=True)


class _OpenStep(object):
    """Clock and counters of the iteration between ``profile_step_start``
    and ``profile_step_commit``."""

    __slots__ = ("atomic_bsz", "began", "sync_time")

    def __init__(self, atomic_bsz):
        self.atomic_bsz = atomic_bsz
        self.began = time.time()
        self.sync_time = 0.0


_METRICS_STATE = None        # the singleton record (lazily loaded)
_DEVICE_TIMER = None         # weakref to a parallel.timer.DeviceStepTimer
_OPEN_STEP = None            # iteration in flight
_PREV_REPORT = None          # host time of the last hints report
_REPORT_THREAD = None        # background fit + report in flight
_GRAD_PARAM_DICT = {}        # data-parallel instance -> (sqr, var)


def _metrics_state():
    global _METRICS_STATE
    if _METRICS_STATE is None:
        _METRICS_STATE = _MetricsState()
        checkpoint.load_state(_METRICS_STATE)
    return _METRICS_STATE


# ---------------------------------------------------------------------------
# profiling one iteration
# ---------------------------------------------------------------------------

def profile_step_start(atomic_bsz):
    global _OPEN_STEP
    _metrics_state()                       # make sure the record exists
    _OPEN_STEP = _OpenStep(atomic_bsz)


def profile_sync_time(sync_time):
    if device_timer() is not None:
        return               # booked from the device record of the step
    _OPEN_STEP.sync_time += sync_time


def profile_step_commit(accumulation_step=False, step_time=None):
    """Book the iteration opened by :func:`profile_step_start` into the
    profile. ``step_time`` (seconds) replaces the host wall-clock for
    device-timed steps. Rank 0 re-fits the performance model and reports
    scheduling hints every ``REPORT_PERIOD_S`` seconds."""
    global _OPEN_STEP
    step, _OPEN_STEP = _OPEN_STEP, None
    elapsed = (time.time() - step.began) if step_time is None else step_time
    row = _metrics_state().profile[
        (env.num_nodes(), env.num_replicas(), step.atomic_bsz)]
    kind = "accum" if accumulation_step else "optim"
    row[kind + "_step_time"] += elapsed
    row[kind + "_count"] += 1
    if accumulation_step:
        return
    row["optim_sync_time"] += step.sync_time
    _maybe_report()


# ---------------------------------------------------------------------------
# device-timed steps (parallel/timer.py)
# ---------------------------------------------------------------------------

def set_device_timer(timer):
    """Install the on-device step timer (``None`` removes it). Only the
    first data-parallel instance of a process registers one: with several
    (a GAN's generator and discriminator) its finalize-to-finalize interval
    is the whole iteration."""
    global _DEVICE_TIMER
    import weakref
    _DEVICE_TIMER = weakref.ref(timer) if timer is not None else None


def device_timer():
    """The active device timer, or ``None`` (host clocks are used)."""
    timer = _DEVICE_TIMER() if _DEVICE_TIMER is not None else None
    return timer if (timer is not None and timer.active()) else None


def device_timer_reset():
    timer = device_timer()
    if timer is not None:
        timer.reset()


def _book_device_records(wait=False):
    timer = device_timer()
    if timer is None:
        return 0
    profile = _metrics_state().profile
    records = timer.drain(wait=wait)
    for rec in records:
        row = profile[rec.key]
        row["optim_step_time"] += rec.step_time
        row["optim_sync_time"] += rec.sync_time
        row["optim_count"] += 1
        if rec.accum_count:
            row["accum_step_time"] += rec.accum_time
            row["accum_count"] += rec.accum_count
    return len(records)


def profile_step_commit_device(accumulation_step, commit):
    """:func:`profile_step_commit` when the device times the steps: the
    iteration is only *noted* here; it is booked (with the device's numbers)
    once the GPU has published them, one or two iterations later, so the host
    never waits. Accumulation micro-steps arrive with the optimizer step that
    closes them."""
    global _OPEN_STEP
    step, _OPEN_STEP = _OPEN_STEP, None
    timer = device_timer()
    if not accumulation_step:
        key = (env.num_nodes(), env.num_replicas(), step.atomic_bsz) \
            if (commit and step is not None) else None
        timer.note(key)
    _book_device_records()
    if commit and not accumulation_step:
        _maybe_report()


def _maybe_report():
    global _PREV_REPORT
    now = time.time()
    if _PREV_REPORT is None:
        _PREV_REPORT = now
    due = now - _PREV_REPORT > REPORT_PERIOD_S
    if not (due and env.replica_rank() == 0):
        return
    if ASYNC_REPORT:
        _PREV_REPORT = now
        _report_in_background()
    else:
        _fit_perf_params()
        _report_sched_hints()
        _PREV_REPORT = time.time()


def _report_in_background():
    """Fit and report from a snapshot of the profile on a daemon thread; a
    report still in flight (slow supervisor) is not doubled up."""
    global _REPORT_THREAD
    if _REPORT_THREAD is not None and _REPORT_THREAD.is_alive():
        return
    table = _profile_snapshot()
    job = env.job_id()

    def work():
        try:
            params = _fit_from(table)
            if params is not None:
                _metrics_state().perf_params = params
            post_sched_hints(_build_sched_hints(
                [key for key, _ in table]), job)
        except Exception:  # noqa: BLE001 - telemetry must not kill a job
            LOG.exception("background performance fit / report failed")
    _REPORT_THREAD = threading.Thread(target=work, daemon=True,
                                      name="adaptdl-report")
    _REPORT_THREAD.start()


def wait_for_report(timeout=None):
    """Join the background report, if any (tests, orderly shutdown)."""
    thread = _REPORT_THREAD
    if thread is not None:
        thread.join(timeout)


# ---------------------------------------------------------------------------
# statistics and configuration pushed in by the trainer
# ---------------------------------------------------------------------------

def update_grad_params(edp_key, grad_norm_sqr, grad_variance):
    """Latest ``(|g|^2, variance)`` of one data-parallel instance. A job may
    hold several (GAN generator + discriminator): the job-level statistic is
    their sum."""
    _GRAD_PARAM_DICT[edp_key] = (float(grad_norm_sqr), float(grad_variance))
    sqr = sum(pair[0] for pair in _GRAD_PARAM_DICT.values())
    var = sum(pair[1] for pair in _GRAD_PARAM_DICT.values())
    _metrics_state().grad_params = (sqr, var)


def update_progress(progress):
    _metrics_state().progress = progress


def get_progress():
    return _metrics_state().progress


def set_batch_size(init_batch_size, max_batch_size, local_bsz_bounds,
                   gradient_accumulation):
    record = _metrics_state()
    record.init_batch_size, record.max_batch_size = \
        init_batch_size, max_batch_size
    record.local_bsz_bounds = local_bsz_bounds
    record.gradient_accumulation = gradient_accumulation


def get_goodput_fn():
    """The job's current goodput model, or ``None`` until both a performance
    fit and gradient statistics exist."""
    record = _metrics_state()
    if record.perf_params is None or record.grad_params is None:
        return None
    return GoodputFunction(record.perf_params, record.grad_params,
                           record.init_batch_size)


# ---------------------------------------------------------------------------
# performance fit and scheduler hints
# ---------------------------------------------------------------------------

def _profile_snapshot():
    """``[(key, counters)]`` of every configuration with at least one
    optimisation step, copied so that it can be read off-thread."""
    return [(key, dict(row)) for key, row in
            list(_metrics_state().profile.items()) if row.get("optim_count")]


def _fit_from(table):
    """Turn a profile snapshot into per-configuration mean step times and
    fit the throughput model to them (``None`` for an empty table)."""
    if not table:
        return None
    nodes, replicas, atomic = (np.array(col) for col in
                               zip(*(key for key, _ in table)))

    def col(name):
        return np.array([row.get(name, 0) for _, row in table], dtype=float)
    optim_total, optim_n = col("optim_step_time"), col("optim_count")
    accum_total, accum_n = col("accum_step_time"), col("accum_count")
    # a measured sync can reach a step's duration by jitter: clamp (the
    # model needs step >= sync)
    sync_total = np.minimum(col("optim_sync_time"), optim_total)
    # an optimisation step minus its synchronisation costs about what an
    # accumulation step costs: pool both kinds of sample for the local time
    local_mean = (accum_total + optim_total - sync_total) / (accum_n + optim_n)
    optim_mean = optim_total / optim_n
    # a row whose sync ate the whole step would put log(0) into the fit
    # (the reference asserts here): keep the local time strictly positive
    local_mean = np.maximum(local_mean, 1e-3 * optim_mean)
    local_mean = np.maximum(local_mean, 1e-7)
    params = fit_perf_params(nodes, replicas, atomic, local_mean, optim_mean)
    if params is None or not np.all(np.isfinite(np.array(params,
                                                          dtype=float))):
        LOG.warning("performance fit did not converge to finite parameters; "
                    "keeping the previous ones")
        return None
    return params


def _fit_perf_params():
    params = _fit_from(_profile_snapshot())
    if params is not None:
        _metrics_state().perf_params = params


def _get_sched_hints():
    """(Ray integration) the record with a fresh fit, or ``None`` before the
    first committed step."""
    record = _metrics_state()
    if not record.profile:
        return None
    _fit_perf_params()
    return record


def _build_sched_hints(profile_keys=None):
    record = _metrics_state()
    if profile_keys is None:
        profile_keys = list(record.profile)
    hints = dict(SCHED_HINTS)
    hints.update(
        initBatchSize=record.init_batch_size,
        maxBatchSize=record.max_batch_size,
        localBszBounds=record.local_bsz_bounds,
        gradientAccumulation=record.gradient_accumulation,
        maxProfiledReplicas=max(key[1] for key in profile_keys))
    if record.perf_params is not None:
        hints["perfParams"] = dict(zip(
            PERF_PARAMS.keys(), (float(v) for v in record.perf_params)))
    if record.grad_params:
        sqr, var = record.grad_params
        hints["gradParams"] = {"norm": float(sqr), "var": float(var)}
    return hints


def _report_sched_hints():
    assert env.replica_rank() == 0
    post_sched_hints(_build_sched_hints(), env.job_id())


def _reset_for_tests():
    global _METRICS_STATE, _OPEN_STEP, _PREV_REPORT
    wait_for_report(5.0)
    if _METRICS_STATE is not None:
        _METRICS_STATE.unregister()
    _METRICS_STATE = _OPEN_STEP = _PREV_REPORT = None
    _GRAD_PARAM_DICT.clear()
    set_device_timer(None)
