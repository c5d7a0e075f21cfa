"""Driver of tests/test_tune_workers.py::test_ray_actor_spawner...: one
elastic trial whose replica is a Ray actor (``RayActorSpawner``) on the
in-process stand-in for Ray: results stream through ``ray.util.queue``, the
group is preempted through the actor's ``preempt`` method, its in-memory
checkpoint seeds a second generation that runs to the end. One replica per
generation: emulated actors share this process."""
import json
import os
import sys

import adaptdl_b200.torch  # noqa: F401 - signal handlers: main thread only
import ray
from adaptdl_b200.ray.aws.worker import _forget_previous_generation
from adaptdl_b200.ray.tune import workers

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tune_workload  # noqa: E402

config = {"lr": 0.05, "epochs": 12, "pause": 0.05}
spawner = workers.RayActorSpawner(placement_group="pg-0",
                                  resources_per_replica={"CPU": 1})
group = workers.WorkerGroup(tune_workload.train_fn, config, ["n0"], spawner)
seen = [group.next_result(timeout=120) for _ in range(3)]
snapshot = group.checkpoint(timeout=120)
group.shutdown()
finished_first = group.finished

_forget_previous_generation()        # a real actor would be a new process
group = workers.WorkerGroup(tune_workload.train_fn, config, ["n1"], spawner,
                            checkpoint=snapshot, generation=1)
first = group.next_result(timeout=120)
last = first
while True:
    result = group.next_result(timeout=120)
    if result is None:
        break
    last = result
group.shutdown()

# the same thing one level up: the Tune trainable built by
# AdaptDLTrainableCreator, driven the way Tune drives a Trainable (train /
# save / stop, then restore in a clone and train to the end)
_forget_previous_generation()
from adaptdl_b200.ray.tune import AdaptDLTrainableCreator  # noqa: E402
cls = AdaptDLTrainableCreator(tune_workload.train_fn, num_workers=1,
                              resources_per_replica={"CPU": 1})
trial = cls(config={"lr": 0.05, "epochs": 8, "pause": 0.05},
            trial_id="t1")
steps = [trial.train() for _ in range(2)]
state = trial.save()
trial.stop()
_forget_previous_generation()
clone = cls(config={"lr": 0.05, "epochs": 8, "pause": 0.05}, trial_id="t1")
clone.restore(state)
tail = []
while True:
    result = clone.train()
    if result.get("done"):
        break
    tail.append(result)
clone.stop()
trainable = {
    "name": cls.__name__, "first": [r["epoch"] for r in steps],
    "generation_saved": state["generation"],
    "resumed": [r["epoch"] for r in tail][:1] + [r["epoch"]
                                                 for r in tail][-1:],
    "resumed_generation": tail[0]["generation"],
    "resources": cls.default_resource_request({}).bundles}

print(json.dumps({
    "trainable": trainable,
    "first_epochs": [r["epoch"] for r in seen],
    "finished_first": finished_first,
    "snapshot_files": sorted(snapshot)[:3],
    "resumed_at": first["epoch"], "restarts": first["restarts"],
    "last_epoch": last["epoch"], "finished": group.finished,
    "calls": [[k, n] for k, n, _ in ray._actors.CALLS]}))
os._exit(0)
