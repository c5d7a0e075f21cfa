"""Driver side of an elastic Tune trial.

:class:`ElasticTrial` is the Ray-free state machine: it owns the trial's
:class:`~adaptdl_b200.ray.tune.workers.WorkerGroup`, turns its reports into
Tune results and implements *save* / *restore* so that a trial cloned onto a
different placement group (``AdaptDLTrial.create_from`` in ``scheduler.py``)
continues from an in-memory checkpoint with another replica count.
:func:`make_trainable` wraps it into a ``ray.tune.Trainable`` subclass whose
replicas are Ray actors on the trial's placement group (reference:
``ray/adaptdl_ray/tune/adaptdl_trainable.py:29-81``).
"""

import logging

from adaptdl_b200.ray.tune.workers import (ProcessSpawner, RayActorSpawner,
                                           WorkerGroup)

LOG = logging.getLogger(__name__)

CHECKPOINT_KEY = "adaptdl_b200_checkpoint"


class ElasticTrial(object):
    """One trial = a sequence of worker groups ("generations"), each on the
    allocation current at the time, chained by in-memory checkpoints."""

    def __init__(self, train_fn, config, allocation, spawner=None,
                 job_id="tune/trial", result_timeout=3600.0):
        self._train_fn, self._config = train_fn, config
        self.allocation = list(allocation)
        self._spawner = spawner or ProcessSpawner()
        self._job_id = job_id
        self._timeout = result_timeout
        self._group = None
        self._snapshot = None          # checkpoint the next group starts from
        self._done = False
        self.generation = 0
        self.last_hints = None

    def _running_group(self):
        if self._group is None:
            self._group = WorkerGroup(
                self._train_fn, self._config, self.allocation, self._spawner,
                checkpoint=self._snapshot, generation=self.generation,
                job_id=self._job_id)
        return self._group

    def step(self):
        """One Tune result: rank 0's next report plus bookkeeping keys;
        ``{"done": True}`` when the training function has returned."""
        if self._done:
            return {"done": True}
        result = self._running_group().next_result(self._timeout)
        if result is None:
            self._done = True
            return {"done": True}
        if result.get("sched_hints"):
            self.last_hints = result["sched_hints"]
        result.setdefault("sched_hints", self.last_hints)
        result.update(num_replicas=len(self.allocation),
                      generation=self.generation, done=False)
        return result

    def save(self):
        """Stop the replicas at an iteration boundary; the returned state
        restarts the trial (here or in a clone) where it stopped."""
        if self._group is not None:
            self._snapshot = self._group.checkpoint(self._timeout)
            self._done = self._done or self._group.finished
            self._group.shutdown()
            self._group = None
            self.generation += 1
        return {CHECKPOINT_KEY: self._snapshot, "generation": self.generation,
                "done": self._done, "sched_hints": self.last_hints}

    def restore(self, state):
        self.stop()
        self._snapshot = state[CHECKPOINT_KEY]
        self.generation = state["generation"]
        self._done = state.get("done", False)
        self.last_hints = state.get("sched_hints")

    def stop(self):
        if self._group is not None:
            self._group.shutdown()
            self._group = None


def current_allocation(num_workers):
    """Node of every replica bundle of the placement group this code runs
    in (bundle 0 belongs to the trial driver), or virtual names outside a
    placement group."""
    from adaptdl_b200.ray import require_ray
    ray = require_ray()
    group = ray.util.get_current_placement_group()
    if group is None:
        return ["virtual-{}".format(i) for i in range(num_workers)], None
    table = ray.util.placement_group_table(group)
    node_of = table.get("bundles_to_node_id", {})
    count = len(table.get("bundles", {})) - 1
    return [str(node_of.get(i + 1, "virtual-{}".format(i)))
            for i in range(max(count, 1))], group


def make_trainable(train_fn, num_workers=1, resources_per_replica=None):
    """``ray.tune.Trainable`` running ``train_fn(config, report)`` on the
    trial's placement group with :class:`ElasticTrial`."""
    from adaptdl_b200.ray import require_ray
    require_ray()
    from ray import tune
    from adaptdl_b200.ray import utils as ray_utils

    class AdaptDLTrainable(tune.Trainable):

        def setup(self, config):
            allocation, group = current_allocation(num_workers)
            self._trial = ElasticTrial(
                train_fn, config, allocation,
                RayActorSpawner(group, resources_per_replica),
                job_id="tune/{}".format(self.trial_id))

        def step(self):
            return self._trial.step()

        def save_checkpoint(self, checkpoint_dir):
            return self._trial.save()

        def load_checkpoint(self, state):
            self._trial.restore(state)

        def cleanup(self):
            self._trial.stop()

        def get_sched_hints(self):
            return self._trial.last_hints

        @classmethod
        def default_resource_request(cls, config):
            return ray_utils.allocation_to_pgf(
                ["virtual-{}".format(i) for i in range(num_workers)],
                resources_per_replica)

    AdaptDLTrainable.__name__ = "AdaptDL_{}".format(
        getattr(train_fn, "__name__", "trainable"))
    return AdaptDLTrainable
