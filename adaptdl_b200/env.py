"""Job configuration, read from ``ADAPTDL_*`` environment variables.

Every knob of a training replica is an environment variable with a default,
exposed through a getter (parity: reference ``adaptdl/adaptdl/env.py:23-173``).
The scheduler / launcher sets them; standalone runs get single-replica
defaults. Getters re-read ``os.environ`` on every call so that test harnesses
and launchers may mutate the environment between restarts.
"""

import os

__all__ = [
    "checkpoint_path", "share_path", "job_id", "master_addr", "master_port",
    "replica_rank", "num_nodes", "num_replicas", "num_restarts",
    "adaptdl_sched_version", "supervisor_url", "from_ray", "local_rank",
    "force_torch_reducer",
]


def _get(name, cast=str, default=None):
    raw = os.environ.get(name)
    if raw is None or raw == "":
        return default
    return cast(raw)


def checkpoint_path():
    """Directory where checkpoints are written (must be shared by replicas
    and survive restarts). ``None`` disables checkpointing."""
    return _get("ADAPTDL_CHECKPOINT_PATH")


def share_path():
    """Directory shared by all replicas of the job (datasets etc.)."""
    return _get("ADAPTDL_SHARE_PATH")


def job_id():
    """Unique job identifier (``namespace/name`` under Kubernetes)."""
    return _get("ADAPTDL_JOB_ID")


def master_addr():
    """Address of the rank-0 replica. Defaults to ``0.0.0.0`` (local)."""
    return _get("ADAPTDL_MASTER_ADDR", default="0.0.0.0")


def master_port():
    """Control-plane port of the rank-0 replica; 0 = pick one (local mode)."""
    return _get("ADAPTDL_MASTER_PORT", int, 0)


def replica_rank():
    """Rank of this replica, in ``[0, num_replicas())``."""
    return _get("ADAPTDL_REPLICA_RANK", int, 0)


def num_nodes():
    """Number of distinct nodes hosting replicas (defaults to one node per
    replica, like the reference)."""
    return _get("ADAPTDL_NUM_NODES", int, None) or num_replicas()


def num_replicas():
    """Total number of replicas (= data-parallel world size)."""
    return _get("ADAPTDL_NUM_REPLICAS", int, 1)


def num_restarts():
    """How many times this job has been restarted (the restart generation)."""
    return _get("ADAPTDL_NUM_RESTARTS", int, 0)


def adaptdl_sched_version():
    """Version of the scheduler that launched this job (or ``None``)."""
    return _get("ADAPTDL_SCHED_VERSION")


def supervisor_url():
    """URL of the job supervisor (rendezvous + scheduling hints), if any."""
    return _get("ADAPTDL_SUPERVISOR_URL")


def from_ray():
    """True when running under the Ray Tune trial scheduler."""
    return _get("ADAPTDL_TUNE_TRIAL_SCHED", lambda s: s.lower() == "true",
                False)


# ---- additions over the reference (B200 single-box runtime) ---------------

def local_rank():
    """GPU ordinal on this box. ``LOCAL_RANK`` (torchrun) wins, then
    ``ADAPTDL_LOCAL_RANK``, then rank modulo replicas-per-node."""
    for name in ("ADAPTDL_LOCAL_RANK", "LOCAL_RANK"):
        val = _get(name, int)
        if val is not None:
            return val
    per_node = max(num_replicas() // max(num_nodes(), 1), 1)
    return replica_rank() % per_node


def force_torch_reducer():
    """``ADAPTDL_B200_REDUCER=torch`` forces the pure-``torch.distributed``
    gradient reducer (NCCL/gloo) instead of the fused sm_100a kernel."""
    return (_get("ADAPTDL_B200_REDUCER", default="auto") or "auto").lower() \
        == "torch"
