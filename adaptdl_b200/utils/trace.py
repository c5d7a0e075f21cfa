"""Opt-in tracing of the step path.

Two independent sinks, both off by default (when off, :func:`traced` returns
the function unchanged and :func:`span` a shared no-op object, so the step
path pays nothing):

``ADAPTDL_B200_NVTX=1``
    every span is also an NVTX range ``adl/<name>``. There is no nsys in the
    target image, but Nsight Compute filters on ranges:
    ``ncu --nvtx --nvtx-include "adl/optimizer_step/" ...`` profiles only the
    kernels launched by the fused optimizer step.
``ADAPTDL_B200_TRACE=<path>``
    host-side timeline: every span becomes a complete ("X") event of a
    Chrome trace (open in ``chrome://tracing`` / Perfetto), one process row
    per replica; written at interpreter exit as ``<path>.rank<r>.json`` (or
    on demand with :func:`dump`). Shows where the HOST spends the step (data
    loading, launch overhead, graph replay, mailbox reads) -- the device side
    is ``tools/step_profile.py`` (CUPTI).

The reference has no tracing beyond its step-time counters.
"""

import atexit
import functools
import json
import os
import threading
import time

NVTX = os.environ.get("ADAPTDL_B200_NVTX", "0") == "1"
TRACE_PATH = os.environ.get("ADAPTDL_B200_TRACE") or None
ENABLED = NVTX or TRACE_PATH is not None

_EVENTS = []
_LOCK = threading.Lock()
_MAX_EVENTS = 2_000_000


class _NullSpan(object):
    __slots__ = ()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL = _NullSpan()


class _Span(object):
    __slots__ = ("name", "t0", "pushed")

    def __init__(self, name):
        self.name = name
        self.t0 = 0
        self.pushed = False

    def __enter__(self):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_push("adl/" + self.name)
            self.pushed = True
        self.t0 = time.perf_counter_ns()
        return self

    def __exit__(self, *exc):
        t1 = time.perf_counter_ns()
        if self.pushed:
            import torch
            torch.cuda.nvtx.range_pop()
        if TRACE_PATH is not None and len(_EVENTS) < _MAX_EVENTS:
            _EVENTS.append((self.name, threading.get_ident(), self.t0, t1))
        return False


def span(name):
    """``with span("phase"): ...``"""
    return _Span(name) if ENABLED else _NULL


def traced(name):
    """Decorator form; the identity when tracing is off."""
    def decorate(fn):
        if not ENABLED:
            return fn

        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            with _Span(name):
                return fn(*args, **kwargs)
        return wrapper
    return decorate


def events():
    """Recorded ``(name, thread, start_ns, end_ns)`` tuples."""
    return list(_EVENTS)


def dump(path=None):
    """Write the Chrome trace; returns the file name (``None`` when there
    is nothing to write)."""
    path = path or TRACE_PATH
    if path is None or not _EVENTS:
        return None
    rank = int(os.environ.get("ADAPTDL_REPLICA_RANK",
                              os.environ.get("RANK", "0")))
    with _LOCK:
        records = [{"name": name, "ph": "X", "pid": rank, "tid": tid,
                    "ts": t0 / 1e3, "dur": (t1 - t0) / 1e3, "cat": "adl"}
                   for name, tid, t0, t1 in _EVENTS]
    records.append({"name": "process_name", "ph": "M", "pid": rank,
                    "args": {"name": "replica {}".format(rank)}})
    out = "{}.rank{}.json".format(path, rank)
    with open(out, "w") as f:
        json.dump({"traceEvents": records, "displayTimeUnit": "ms"}, f)
    return out


def summary():
    """``{name: (count, total_ms, mean_us)}`` of the recorded spans."""
    table = {}
    for name, _, t0, t1 in _EVENTS:
        count, total = table.get(name, (0, 0))
        table[name] = (count + 1, total + (t1 - t0))
    return {name: (count, total / 1e6, total / count / 1e3)
            for name, (count, total) in table.items()}


if TRACE_PATH is not None:
    atexit.register(dump)
