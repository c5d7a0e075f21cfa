"""Speedup of a job as a function of (nodes, replicas), relative to one
replica, each at its own best batch size:

    speedup(n, r) = max_bsz goodput(n, r, bsz) / max_bsz goodput(1, 1, bsz)

Evaluations are memoised in a small dense table because the policy's genetic
search asks the same few hundred questions thousands of times (parity:
reference ``policy/speedup.py:18-70``)."""

import numpy as np


class SpeedupFunction(object):

    def __init__(self, goodput_fn, max_batch_size=None, atomic_bsz_range=None,
                 accumulation=False, mem_size=32):
        self._goodput_fn = goodput_fn
        self._kwargs = dict(max_batch_size=max_batch_size,
                            atomic_bsz_range=atomic_bsz_range,
                            accumulation=accumulation)
        self._mem_size = mem_size
        self._base_goodput, _, _ = goodput_fn.optimize(
            num_nodes=1, num_replicas=1, **self._kwargs)
        self._table = np.full((mem_size, mem_size), -1.0)
        self._table[0, 0] = 0.0          # nothing allocated, no progress

    def __call__(self, num_nodes, num_replicas):
        assert np.all(np.less_equal(0, num_nodes))
        assert np.all(np.less_equal(num_nodes, num_replicas))
        assert np.all((np.asarray(num_nodes) > 0)
                      == (np.asarray(num_replicas) > 0))
        scalar = np.isscalar(num_nodes) and np.isscalar(num_replicas)
        shape = np.broadcast(num_nodes, num_replicas).shape
        nodes = np.broadcast_to(num_nodes, shape).reshape(-1).astype(int)
        repl = np.broadcast_to(num_replicas, shape).reshape(-1).astype(int)
        out = np.full(nodes.shape, -1.0)
        small = repl < self._mem_size
        out[small] = self._table[nodes[small], repl[small]]
        todo = out < 0
        if np.any(todo):
            pairs, inverse = np.unique(
                np.stack([nodes[todo], repl[todo]]), axis=1,
                return_inverse=True)
            inverse = np.asarray(inverse).reshape(-1)
            goodput, _, _ = self._goodput_fn.optimize(
                pairs[0], pairs[1], **self._kwargs)
            speedup = np.asarray(goodput, dtype=float) / self._base_goodput
            keep = pairs[1] < self._mem_size
            self._table[pairs[0][keep], pairs[1][keep]] = speedup[keep]
            out[todo] = speedup[inverse]
        assert np.all(out >= 0)
        out = out.reshape(shape)
        return out.item() if scalar else out
