"""``import adaptdl_ray`` -> ``adaptdl_b200.ray`` (see adaptdl_b200/compat.py): scripts
written for petuum/adaptdl run on this framework without edits. The reference's
module paths (``ray/adaptdl_ray/{adaptdl,tune,aws}/*.py``) are kept importable."""
from adaptdl_b200.compat import alias as _alias

_alias("adaptdl_ray", "adaptdl_b200.ray", renames={
    # from adaptdl_ray.adaptdl import AdaptDLAllocator, AdaptDLJobMixin
    "adaptdl": "",
    "adaptdl.adaptdl_allocator": "allocator",
    "adaptdl.adaptdl_job_mixin": "job_mixin",
    # from adaptdl_ray.tune.adaptdl_trial_sched import AdaptDLScheduler
    "tune.adaptdl_trial_sched": "tune.scheduler",
    "tune.adaptdl_trial": "tune.scheduler",
    "tune.adaptdl_trainable": "tune.scheduler",
})
