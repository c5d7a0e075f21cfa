"""Ray Tune integration: ``AdaptDLScheduler`` re-allocates GPUs among the
elastic trials of a Tune experiment with the Pollux policy (reference:
``ray/adaptdl_ray/tune``).

``workers`` (elastic worker groups) and ``trainable.ElasticTrial`` work
without Ray; ``AdaptDLScheduler`` / ``AdaptDLTrial`` /
``AdaptDLTrainableCreator`` need ``ray[tune]``."""

_LAZY = ("AdaptDLScheduler", "AdaptDLTrial", "AdaptDLTrainableCreator")


def __getattr__(name):
    if name in _LAZY:
        from adaptdl_b200.ray.tune import scheduler
        return getattr(scheduler, name)
    if name in ("ElasticTrial", "make_trainable"):
        from adaptdl_b200.ray.tune import trainable
        return getattr(trainable, name)
    if name in ("WorkerGroup", "ProcessSpawner", "RayActorSpawner"):
        from adaptdl_b200.ray.tune import workers
        return getattr(workers, name)
    raise AttributeError(name)
