"""Collectives on arbitrary (small, picklable) Python objects.

General but not fast: use it for control decisions (exit flag, batch-size
broadcast, Accumulator sums, port exchange). Tensors go through the fused
gradient reducer / ``torch.distributed``. Functions here must be invoked in
the same order on all replicas (parity: reference ``collective.py:34-144``).
"""

from adaptdl_b200 import env
from adaptdl_b200.reducer import Reducer, Future, default_reduce_fn

__all__ = ["initialize", "teardown", "allreduce", "allreduce_async",
           "broadcast", "is_initialized", "default_reduce_fn", "Future"]

_REDUCER = None


def initialize(master_addr=None, master_port=None, replica_rank=None,
               num_replicas=None):
    """Connect this replica to the control plane. Blocks until rank 0's
    server accepts the connection.

    Raises:
        RuntimeError: if already initialised.
    """
    global _REDUCER
    if _REDUCER is not None:
        raise RuntimeError("{} is already initialized".format(__name__))
    rank = env.replica_rank() if replica_rank is None else replica_rank
    size = env.num_replicas() if num_replicas is None else num_replicas
    addr = env.master_addr() if master_addr is None else master_addr
    port = env.master_port() if master_port is None else master_port
    _REDUCER = Reducer(rank, size, addr, port)


def is_initialized():
    return _REDUCER is not None


def teardown():
    """Disconnect from the control plane (the reference leaves this
    unimplemented; elastic in-process tests need it)."""
    global _REDUCER
    if _REDUCER is None:
        raise RuntimeError("{} has not been initialized".format(__name__))
    _REDUCER.close()
    _REDUCER = None


def _reducer():
    if _REDUCER is None:
        raise RuntimeError("{} has not been initialized".format(__name__))
    return _REDUCER


def allreduce(value, reduce_fn=default_reduce_fn):
    """Reduce ``value`` across replicas; everyone gets the result."""
    return _reducer().allreduce(value, reduce_fn)


def allreduce_async(value, reduce_fn=default_reduce_fn):
    """Non-blocking :func:`allreduce`; returns a ``Future``."""
    return _reducer().allreduce_async(value, reduce_fn)


def broadcast(value):
    """Broadcast rank 0's ``value`` to all replicas."""
    return _reducer().broadcast(value)
