"""Data-parallel machinery: flat gradient layout, gradient reducers (fused
sm_100a kernel over NVLink peer memory, or stock torch collectives), symmetric
memory, topology discovery, CUDA-graph step capture."""


def make_reducer(param_groups, world_size, rank, should_sync,
                 bucket_cap_mb=25, process_group=None, backend="auto",
                 name="reducer"):
    """Pick the gradient reducer for this process.

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
    want_cuda = backend == "cuda" or (
        backend == "auto" and device is not None and device.type == "cuda"
        and not env.force_torch_reducer())
    if want_cuda:
        try:
            from adaptdl_b200.parallel.reducer_cuda import CudaGradReducer
            return CudaGradReducer(param_groups, world_size, rank,
                                   should_sync, bucket_cap_mb,
                                   process_group=process_group, name=name)
        except Exception:
            if backend == "cuda" or (device is not None
                                     and device.type == "cuda"):
                # On a GPU box a missing/broken extension must be loud, not a
                # silent fallback to library collectives.
                raise
    from adaptdl_b200.parallel.reducer_torch import TorchGradReducer
    return TorchGradReducer(param_groups, world_size, rank, should_sync,
                            bucket_cap_mb, process_group=process_group,
                            name=name)
