"""Data-parallel machinery: flat gradient layout, gradient reducers (fused
sm_100a kernel over NVLink peer memory, or stock torch collectives), symmetric
memory, topology discovery, CUDA-graph step capture."""


import logging

LOG = logging.getLogger(__name__)

# Default bucket caps. A fused bucket kernel costs ~10-20 us on the comm
# stream (nothing on the host inside a CUDA graph), so buckets smaller than
# DDP's 25 MB pay off: they start reducing while backward is still producing
# gradients and keep the un-overlappable tail short. Measured on ResNet-18 at
# N = 2 (profiles/r2_n2): 25 MB 2.110 ms/step, 8 MB 2.084, 4 MB 2.100,
# 2 MB 2.124. The first and the last bucket have their own small caps
# (reducer_base.FIRST_BUCKET_BYTES / LAST_BUCKET_BYTES).
CUDA_BUCKET_CAP_MB = 8
TORCH_BUCKET_CAP_MB = 25


def choose_backend(backend, device_type, num_nodes, force_torch=False):
    """``"cuda"`` or ``"torch"`` for a ``backend`` request (see
    :func:`make_reducer`). The fused kernels reduce through peer-mapped
    memory, which exists inside one NVLink / NVSwitch domain only: a job whose
    replicas span several nodes (``ADAPTDL_NUM_NODES`` > 1, as the cluster
    scheduler may well allocate) uses NCCL through the torch reducer in
    ``auto`` mode. (A two-level reducer -- fused inside the node, NCCL across
    -- is the planned successor.)"""
    if backend == "torch":
        return "torch"
    if backend == "cuda":
        if num_nodes > 1:
            raise ValueError(
                "reducer='cuda' needs all replicas in one NVLink domain, "
                "but the job spans {} nodes".format(num_nodes))
        return "cuda"
    if backend != "auto":
        raise ValueError("unknown reducer backend {!r}".format(backend))
    if device_type != "cuda" or force_torch:
        return "torch"
    if num_nodes > 1:
        LOG.info("replicas run in %d separate hosts / containers: gradients "
                 "are reduced with NCCL (the fused peer-memory kernels need "
                 "all ranks in one)", num_nodes)
        return "torch"
    return "cuda"


def hosts_spanned(world_size, process_group=None):
    """How many separate memory-sharing domains the replicas run in (1 =
    the fused peer-memory path is possible).

    Peer mappings are set up by passing file descriptors between the rank
    processes over unix sockets, so the ranks must share a host AND a
    container: ``ADAPTDL_NUM_NODES`` > 1 settles it; otherwise the ranks
    compare host names (every Kubernetes pod has its own, even on one node --
    one-GPU pods therefore reduce with NCCL; run one multi-GPU pod per node,
    or the single-box launchers, to get the fused path). The reference's
    default of "one node per replica" when the variable is unset is NOT used:
    it would send every hand-launched single-box job to NCCL."""
    import os
    stated = int(os.environ.get("ADAPTDL_NUM_NODES") or 0)
    if world_size <= 1:
        return 1
    if stated > 1:
        return stated
    import socket
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    names = [None] * dist.get_world_size(process_group)
    dist.all_gather_object(names, socket.gethostname(), group=process_group)
    return len(set(names))


def make_reducer(param_groups, world_size, rank, should_sync,
                 bucket_cap_mb=None, process_group=None, backend="auto",
                 name="reducer"):
    """Pick the gradient reducer for this process.

    ``bucket_cap_mb``: ``None`` = the reducer's own default (``CUDA_BUCKET_CAP_MB`` /
    ``TORCH_BUCKET_CAP_MB``).

    ``backend``: ``"cuda"`` = fused sm_100a kernels (raises if unavailable),
    ``"torch"`` = stock torch collectives, ``"auto"`` = fused kernels when the
    parameters are on a CUDA device and the native extension loads, else
    torch.
    """
    from adaptdl_b200 import env
    device = None
    for group in param_groups:
        for p in group["params"]:
            device = p.device
            break
        if device is not None:
            break
    want_cuda = choose_backend(
        backend, device.type if device is not None else None,
        hosts_spanned(world_size, process_group),
        env.force_torch_reducer()) == "cuda"
    if want_cuda:
        try:
            from adaptdl_b200.parallel.reducer_cuda import CudaGradReducer
            return CudaGradReducer(param_groups, world_size, rank,
                                   should_sync,
                                   bucket_cap_mb or CUDA_BUCKET_CAP_MB,
                                   process_group=process_group, name=name)
        except Exception:
            if backend == "cuda" or (device is not None
                                     and device.type == "cuda"):
                # On a GPU box a missing/broken extension must be loud, not a
                # silent fallback to library collectives.
                raise
    from adaptdl_b200.parallel.reducer_torch import TorchGradReducer
    return TorchGradReducer(param_groups, world_size, rank, should_sync,
                            bucket_cap_mb or TORCH_BUCKET_CAP_MB,
                            process_group=process_group,
                            name=name)
