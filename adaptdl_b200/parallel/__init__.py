"""Data-parallel machinery: flat gradient layout, gradient reducers (fused
sm_100a kernel over NVLink peer memory, or stock torch collectives), symmetric
memory, topology discovery, CUDA-graph step capture."""


# Default bucket caps. The fused kernels cost ~20 us per bucket on the comm
# stream (nothing on the host inside a CUDA graph), so smaller buckets than
# DDP's are plausible: with 25 MB, ResNet-18's 22 MB of bf16 gradients are one
# bucket that only completes with the first layer, i.e. the whole all-reduce
# (~50 us at N=8) is exposed. Smaller buckets overlap it with backward but put
# flag-spinning CTAs next to the backward kernels; that trade has not been
# measured yet (ROUND2_PLAN.md, `bench.py --bucket-cap-mb 4`), so the default
# stays at the value every published number was taken with.
CUDA_BUCKET_CAP_MB = 25
TORCH_BUCKET_CAP_MB = 25


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
    want_cuda = backend == "cuda" or (
        backend == "auto" and device is not None and device.type == "cuda"
        and not env.force_torch_reducer())
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
