"""Tune trial scheduler / trial / trainable for elastic trials.

* ``AdaptDLTrainableCreator(train_fn, num_workers=...)`` wraps a training
  function into a Tune trainable whose replicas are Ray actors forming an
  ordinary adaptdl_b200 job (``workers.py`` / ``trainable.py``); its
  checkpoints are in-memory objects, so a clone on another placement group
  resumes with a different replica count.
* ``AdaptDLTrial`` = a Tune ``Trial`` that can be *cloned onto a new
  placement group* and restored from an in-memory checkpoint (= rescale).
* ``AdaptDLScheduler`` (a ``TrialScheduler``) calls the allocator every
  ``_RESCHEDULE_TRIGGER`` results and rescales / pauses / resumes trials.

All classes are created lazily by :func:`_build` because their base classes
come from Ray.
"""

import logging

from adaptdl_b200.ray.allocator import AdaptDLAllocator
from adaptdl_b200.ray.job_mixin import AdaptDLJobMixin
from adaptdl_b200.ray import utils as ray_utils

LOG = logging.getLogger(__name__)
RESCHEDULE_EVERY_N_RESULTS = 100

_CLASSES = {}


def _build():
    if _CLASSES:
        return _CLASSES
    from adaptdl_b200.ray import require_ray
    require_ray()
    from ray.tune.experiment import Trial
    from ray.tune.schedulers import TrialScheduler

    class AdaptDLTrial(AdaptDLJobMixin, Trial):
        """A trial whose resources can change across its lifetime."""

        def __init__(self, *args, **kwargs):
            self.rescale_count = kwargs.pop("rescale_count", 0)
            self._cached_metrics = None
            super().__init__(*args, job_id=kwargs.pop("job_id", None),
                             **kwargs)
            self._job_id = self._job_id or self.trial_id

        def _allocation_in_use(self):
            return ray_utils.pgf_to_allocation(self.placement_group_factory)

        def _fetch_metrics(self):
            hints = getattr(self, "last_result", {}).get("sched_hints")
            return hints or self._cached_metrics

        @classmethod
        def create_from(cls, trial, trial_runner, new_allocation,
                        copy_state=False):
            """Clone ``trial`` onto ``new_allocation`` and restore it from a
            checkpoint taken now."""
            checkpoint = trial_runner.trial_executor.save(
                trial, storage="memory") if copy_state else None
            clone = cls(trial.trainable_name, config=trial.config,
                        experiment_tag=trial.experiment_tag,
                        evaluated_params=trial.evaluated_params,
                        stopping_criterion=trial.stopping_criterion,
                        trial_id=trial.trial_id,
                        placement_group_factory=ray_utils.allocation_to_pgf(
                            new_allocation),
                        rescale_count=getattr(trial, "rescale_count", 0) + 1)
            clone._cached_metrics = trial._fetch_metrics() \
                if hasattr(trial, "_fetch_metrics") else None
            if checkpoint is not None:
                clone.restore_path = None
                clone.on_checkpoint(checkpoint)
            return clone

    class AdaptDLScheduler(TrialScheduler):
        """Re-allocates devices among trials with the Pollux policy."""

        def __init__(self, allocator=None):
            self._allocator = allocator or AdaptDLAllocator()
            self._results = 0
            self._allocs = {}

        def on_trial_add(self, trial_runner, trial):
            if not isinstance(trial, AdaptDLTrial):
                trials = trial_runner._trials
                idx = trials.index(trial)
                trials[idx] = AdaptDLTrial.create_from(
                    trial, trial_runner,
                    self._allocator.default_allocation())

        def on_trial_error(self, trial_runner, trial):
            pass

        def on_trial_result(self, trial_runner, trial, result):
            self._results += 1
            if self._results % RESCHEDULE_EVERY_N_RESULTS:
                return TrialScheduler.CONTINUE
            trials = [t for t in trial_runner.get_trials()
                      if t.status in (Trial.RUNNING, Trial.PAUSED,
                                      Trial.PENDING)]
            self._allocs, _ = self._allocator.allocate(trials)
            if trial.job_id in self._allocs:
                alloc = self._allocs.pop(trial.job_id)
                if not alloc:
                    return TrialScheduler.PAUSE
                idx = trial_runner._trials.index(trial)
                trial_runner._trials[idx] = AdaptDLTrial.create_from(
                    trial, trial_runner, alloc, copy_state=True)
                return TrialScheduler.STOP
            return TrialScheduler.CONTINUE

        def on_trial_complete(self, trial_runner, trial, result):
            pass

        def on_trial_remove(self, trial_runner, trial):
            pass

        def choose_trial_to_run(self, trial_runner):
            for trial in trial_runner.get_trials():
                if trial.status == Trial.PENDING and \
                        trial_runner.trial_executor.has_resources_for_trial(
                            trial):
                    return trial
            for trial in trial_runner.get_trials():
                if trial.status == Trial.PAUSED and \
                        self._allocs.get(trial.job_id):
                    idx = trial_runner._trials.index(trial)
                    new = AdaptDLTrial.create_from(
                        trial, trial_runner,
                        self._allocs.pop(trial.job_id), copy_state=True)
                    trial_runner._trials[idx] = new
                    return new
            return None

        def debug_string(self):
            return "AdaptDLScheduler (adaptdl_b200)"

    def AdaptDLTrainableCreator(func, num_workers=1,
                                resources_per_replica=None, **_ignored):
        """Wrap ``func(config, report)`` -- an adaptdl_b200 training loop
        that calls ``report(**metrics)`` -- as an elastic Tune trainable
        whose replicas run as Ray actors (``trainable.py``)."""
        from adaptdl_b200.ray.tune.trainable import make_trainable
        return make_trainable(func, num_workers, resources_per_replica)

    _CLASSES.update(AdaptDLTrial=AdaptDLTrial,
                    AdaptDLScheduler=AdaptDLScheduler,
                    AdaptDLTrainableCreator=AdaptDLTrainableCreator)
    return _CLASSES


def __getattr__(name):
    if name in ("AdaptDLTrial", "AdaptDLScheduler",
                "AdaptDLTrainableCreator"):
        return _build()[name]
    raise AttributeError(name)
