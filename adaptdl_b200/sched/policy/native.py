"""The Pollux search on the native core (``csrc/host/adl_pollux.cpp``).

:func:`minimize` has the contract of :func:`nsga2.minimize` for a
:class:`pollux.ClusterProblem`: same operators, objectives and survival rule,
executed candidate by candidate in C++ on a sparse representation (per job
the few (node, replicas) entries it has; every operator costs O(entries),
candidates spread over threads) instead of ~25 numpy passes per generation
over the dense ``[population, jobs, nodes]`` tensor. The one thing
that stays in Python is the goodput model: the core keeps a speedup table per
job, stops when a candidate needs an entry it does not have, and the entries
are computed here through the jobs' ``SpeedupFunction`` objects (vectorised,
a handful of round trips per cycle -- the same few hundred allocations come
back all the time).
"""

import ctypes
import os

import numpy as np

from adaptdl_b200._native import host

_MAX_FIT_CAP = 1 << 20


def available():
    return host.load() is not None


def _ptr(array, ctype):
    return array.ctypes.data_as(ctypes.POINTER(ctype))


class NativeSearch(object):
    """One optimisation cycle's search object (owns the C++ state)."""

    def __init__(self, problem, pop_size, n_gen, seed, threads=0):
        self._lib = host.load()
        if self._lib is None:
            raise RuntimeError("native host library unavailable")
        self._problem = problem
        J, W = problem.base.shape
        self.shape = (J, W)
        job_res = np.ascontiguousarray(problem.job_res, dtype=np.int64)
        node_res = np.ascontiguousarray(problem.node_res, dtype=np.int64)
        R = job_res.shape[1]
        base = np.ascontiguousarray(problem.base, dtype=np.int32)
        pinned = np.zeros(J, dtype=np.uint8)
        pinned[problem.pinned] = 1
        big = np.iinfo(np.int32).max
        min_rep = np.ascontiguousarray(
            np.minimum(problem.min_replicas, big), dtype=np.int32)
        max_rep = np.ascontiguousarray(
            np.minimum(problem.max_replicas[:, 0], big), dtype=np.int32)
        min_fill = np.ascontiguousarray(problem.min_fill, dtype=np.int32)
        max_fit = np.ascontiguousarray(
            np.minimum(problem.max_fit, _MAX_FIT_CAP), dtype=np.int32)
        weight = np.ascontiguousarray(
            problem.dominant_share * len(problem.nodes), dtype=np.float64)
        self._handle = self._lib.adl_pollux_create(
            J, W, R, _ptr(job_res, ctypes.c_int64),
            _ptr(node_res, ctypes.c_int64), _ptr(base, ctypes.c_int32),
            _ptr(pinned, ctypes.c_uint8), _ptr(min_rep, ctypes.c_int32),
            _ptr(max_rep, ctypes.c_int32), _ptr(min_fill, ctypes.c_int32),
            _ptr(max_fit, ctypes.c_int32), _ptr(weight, ctypes.c_double),
            float(problem.restart_penalty), int(pop_size), int(n_gen),
            int(seed) & (2 ** 64 - 1), int(threads))
        if not self._handle:
            raise RuntimeError("adl_pollux_create rejected the problem")
        self.round_trips = 0
        self.entries_filled = 0

    def close(self):
        if self._handle:
            self._lib.adl_pollux_destroy(self._handle)
            self._handle = None

    __del__ = close

    def seed(self, initial):
        initial = np.ascontiguousarray(initial, dtype=np.int32)
        assert initial.shape[1:] == self.shape
        kept = self._lib.adl_pollux_seed(
            self._handle, _ptr(initial, ctypes.c_int32), len(initial))
        if kept <= 0:
            raise RuntimeError("adl_pollux_seed failed")

    def _fill_missing(self, count):
        job = np.empty(count, dtype=np.int32)
        nodes = np.empty(count, dtype=np.int32)
        replicas = np.empty(count, dtype=np.int32)
        got = self._lib.adl_pollux_missing(
            self._handle, _ptr(job, ctypes.c_int32),
            _ptr(nodes, ctypes.c_int32), _ptr(replicas, ctypes.c_int32),
            count)
        assert got == count
        value = np.empty(count, dtype=np.float64)
        order = np.argsort(job, kind="stable")
        bounds = np.flatnonzero(np.diff(job[order])) + 1
        for idx in np.split(order, bounds):
            fn = self._problem.jobs[int(job[idx[0]])].speedup_fn
            fn = getattr(fn, "lookup", fn)
            value[idx] = np.asarray(
                fn(nodes[idx].astype(np.int64),
                   replicas[idx].astype(np.int64)), dtype=np.float64)
        value = np.nan_to_num(value, nan=0.0, posinf=0.0, neginf=0.0)
        rc = self._lib.adl_pollux_fill(
            self._handle, count, _ptr(job, ctypes.c_int32),
            _ptr(nodes, ctypes.c_int32), _ptr(replicas, ctypes.c_int32),
            _ptr(value, ctypes.c_double))
        if rc != 0:
            raise RuntimeError("adl_pollux_fill failed")
        self.round_trips += 1
        self.entries_filled += count

    def run(self):
        while True:
            missing = self._lib.adl_pollux_run(self._handle)
            if missing < 0:
                raise RuntimeError("adl_pollux_run failed")
            if missing == 0:
                break
            self._fill_missing(missing)
        if os.environ.get("ADAPTDL_B200_POLICY_TIMING"):
            t = (ctypes.c_double * 5)()
            self._lib.adl_pollux_timing(self._handle, t)
            print("pollux core seconds: rank %.3f breed %.3f dedupe %.3f "
                  "tables %.3f select %.3f; %d round trips, %d entries"
                  % (tuple(t) + (self.round_trips, self.entries_filled)))
        n = self._lib.adl_pollux_population(self._handle)
        states = np.empty((n,) + self.shape, dtype=np.int32)
        values = np.empty((n, 2), dtype=np.float64)
        self._lib.adl_pollux_result(self._handle,
                                    _ptr(states, ctypes.c_int32),
                                    _ptr(values, ctypes.c_double))
        return states, values

    # single-candidate entry points (tests)

    def repair(self, state, stream=0):
        out = np.ascontiguousarray(state, dtype=np.int32).copy()
        self._lib.adl_pollux_repair(self._handle,
                                    _ptr(out, ctypes.c_int32), stream)
        return out

    def mutate(self, state, stream=0):
        out = np.ascontiguousarray(state, dtype=np.int32).copy()
        self._lib.adl_pollux_mutate(self._handle,
                                    _ptr(out, ctypes.c_int32), stream)
        return out


def minimize(problem, initial, pop_size, n_gen, rng):
    """``(states, values)`` of the final population, like
    :func:`nsga2.minimize`; the search's random streams are seeded from one
    draw of ``rng``."""
    seed = int(rng.integers(0, np.iinfo(np.int64).max))
    threads = int(os.environ.get("ADAPTDL_B200_POLICY_THREADS", "0"))
    search = NativeSearch(problem, pop_size, n_gen, seed, threads)
    try:
        search.seed(initial)
        return search.run()
    finally:
        search.close()
