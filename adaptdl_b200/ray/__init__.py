"""Ray integration (optional: needs ``ray[tune]``, which is not a hard
dependency): a Tune trial scheduler that rescales elastic trials with the
Pollux policy (``tune/``), and a single-job elastic controller for Ray
clusters on AWS spot instances (``aws/``). Reference: ``ray/adaptdl_ray``.

Everything that does not strictly need Ray (allocation <-> placement-group
arithmetic, the single-job replica optimiser, checkpoint (de)serialisation,
the spot-termination poller) is importable and tested without it."""

import os


def have_ray():
    try:
        import ray  # noqa: F401
        return True
    except ImportError:
        return False


def require_ray():
    if not have_ray():
        raise ImportError("this feature needs the optional 'ray[tune]' "
                          "package")
    import ray
    return ray


def tune_checkpoint_dir(for_save):
    """Checkpoint directory under Ray Tune (used by
    ``adaptdl_b200.checkpoint`` when ``ADAPTDL_TUNE_TRIAL_SCHED`` is set)."""
    if for_save:
        path = os.path.join("/tmp", "adaptdl-tune-ckpt-{}".format(
            os.getpid()))
        os.makedirs(path, exist_ok=True)
        return path
    try:
        from ray.tune import session
        return session.get_session().get_checkpoint()
    except Exception:  # noqa: BLE001
        return None



def __getattr__(name):
    # ``from adaptdl_ray.adaptdl import AdaptDLAllocator, AdaptDLJobMixin,
    # default_device`` (the reference's ray/adaptdl_ray/adaptdl/__init__.py)
    if name == "AdaptDLAllocator":
        from adaptdl_b200.ray.allocator import AdaptDLAllocator
        return AdaptDLAllocator
    if name == "AdaptDLJobMixin":
        from adaptdl_b200.ray.job_mixin import AdaptDLJobMixin
        return AdaptDLJobMixin
    if name == "default_device":
        from adaptdl_b200.ray.config import default_device
        return default_device
    raise AttributeError(name)
