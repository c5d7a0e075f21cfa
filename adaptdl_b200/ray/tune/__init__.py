"""Ray Tune integration: ``AdaptDLScheduler`` re-allocates GPUs among the
elastic trials of a Tune experiment with the Pollux policy (reference:
``ray/adaptdl_ray/tune``). Needs ``ray[tune]``."""


def __getattr__(name):
    if name in ("AdaptDLScheduler", "AdaptDLTrial", "AdaptDLTrainableCreator"):
        from adaptdl_b200.ray.tune import scheduler
        return getattr(scheduler, name)
    raise AttributeError(name)
