"""Speedup of a job as a function of its allocation.

``SpeedupFunction(goodput_fn, ...)(num_nodes, num_replicas)`` is the job's
best achievable goodput on that allocation (batch size and gradient
accumulation re-optimised for it) divided by its best goodput on a single
replica, so ``speedup(1, 1) == 1`` and an empty allocation is worth 0.

The genetic search in ``pollux.py`` evaluates the same few hundred
allocations tens of thousands of times per cycle, so results are kept in a
dense ``[nodes, replicas]`` table (allocations beyond the table are computed
on demand and not stored). Capabilities of the reference's
``policy/speedup.py:18-70``.
"""

import numpy as np

_UNKNOWN = -1.0


class SpeedupFunction(object):

    def __init__(self, goodput_fn, max_batch_size=None, atomic_bsz_range=None,
                 accumulation=False, mem_size=32):
        self._goodput_fn = goodput_fn
        self._search = {"max_batch_size": max_batch_size,
                        "atomic_bsz_range": atomic_bsz_range,
                        "accumulation": accumulation}
        self._limit = mem_size
        self._cache = np.full((mem_size, mem_size), _UNKNOWN)
        self._cache[0, 0] = 0.0                  # nothing allocated
        self._unit = self._best_goodput(1, 1)    # the denominator

    def _best_goodput(self, nodes, replicas):
        best, _atomic_bsz, _accum_steps = self._goodput_fn.optimize(
            nodes, replicas, **self._search)
        return best

    def _lookup(self, nodes, replicas):
        """Cached speedups for flat integer arrays (``_UNKNOWN`` where the
        table has no answer yet or cannot hold one)."""
        found = np.full(nodes.shape, _UNKNOWN)
        inside = replicas < self._limit
        found[inside] = self._cache[nodes[inside], replicas[inside]]
        return found

    def _compute(self, nodes, replicas):
        """Speedups of the DISTINCT allocations among the arguments, written
        back to the table where they fit; returns one value per argument."""
        distinct, where = np.unique(np.stack([nodes, replicas]), axis=1,
                                    return_inverse=True)
        n, r = distinct
        speedup = np.asarray(self._best_goodput(n, r), dtype=float) \
            / self._unit
        storable = r < self._limit
        self._cache[n[storable], r[storable]] = speedup[storable]
        return speedup[np.asarray(where).reshape(-1)]

    def lookup(self, nodes, replicas):
        """Speedups for flat integer arrays of VALID allocations (what the
        policy's inner loop has): the table answer, computing what is
        missing, without the checks and broadcasting of ``__call__``."""
        result = self._lookup(nodes, replicas)
        missing = result == _UNKNOWN
        if missing.any():
            result[missing] = self._compute(nodes[missing], replicas[missing])
        return result

    def __call__(self, num_nodes, num_replicas):
        nodes_in, replicas_in = np.asarray(num_nodes), np.asarray(num_replicas)
        if np.any(nodes_in < 0) or np.any(nodes_in > replicas_in) or \
                np.any((nodes_in > 0) != (replicas_in > 0)):
            raise AssertionError(
                "an allocation needs 0 <= nodes <= replicas, and replicas "
                "exactly when it has nodes")
        shape = np.broadcast(nodes_in, replicas_in).shape
        nodes = np.broadcast_to(nodes_in, shape).astype(int).ravel()
        replicas = np.broadcast_to(replicas_in, shape).astype(int).ravel()
        result = self._lookup(nodes, replicas)
        missing = result == _UNKNOWN
        if missing.any():
            result[missing] = self._compute(nodes[missing], replicas[missing])
        if np.any(result < 0):
            raise AssertionError("negative speedup from the goodput model")
        if np.isscalar(num_nodes) and np.isscalar(num_replicas):
            return float(result[0])
        return result.reshape(shape)
