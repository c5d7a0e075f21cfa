"""``adaptdl_b200.ray.tune`` (trial + trial scheduler) driven through a
stand-in for Ray's API (``tests/fixtures/fake_ray``): Ray itself cannot be
installed in this image, so the classes are exercised against a fixture
that reproduces the Ray 2.x signatures they subclass and call -- the
scenario of the reference's ``ray/adaptdl_ray/tests/test_trial_sched.py``
(trials are added, report results with scheduling hints, get re-allocated,
paused and resumed) without a live cluster."""

import os
import sys

import pytest

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                       "fixtures", "fake_ray")


@pytest.fixture
def tune(monkeypatch):
    try:
        import ray  # noqa: F401
        real = "fixture" not in getattr(ray, "__version__", "")
    except ImportError:
        real = False
    if real:
        pytest.skip("a real Ray is installed: covered by its own tests")
    monkeypatch.syspath_prepend(FIXTURE)
    for name in [m for m in sys.modules if m == "ray"
                 or m.startswith("ray.")]:
        monkeypatch.delitem(sys.modules, name)
    import adaptdl_b200.ray.tune.scheduler as sched
    monkeypatch.setattr(sched, "_CLASSES", {})
    monkeypatch.setattr(sched, "RESCHEDULE_EVERY_N_RESULTS", 2)
    yield sched
    for name in [m for m in sys.modules if m == "ray"
                 or m.startswith("ray.")]:
        sys.modules.pop(name, None)


def _hints(var=1.0, max_profiled=4):
    return {
        "perfParams": {"alpha_c": 0.1, "beta_c": 0.01, "alpha_n": 0.0,
                       "beta_n": 0.0, "alpha_r": 0.0, "beta_r": 0.0,
                       "gamma": 1.0},
        "gradParams": {"norm": 0.01, "var": var},
        "initBatchSize": 128, "maxBatchSize": 4096,
        "localBszBounds": [32, 256], "gradientAccumulation": False,
        "maxProfiledReplicas": max_profiled,
    }


class _Executor(object):
    def __init__(self):
        self.saved = []

    def save(self, trial, storage="memory"):
        self.saved.append((trial.trial_id, storage))
        return {"trial": trial.trial_id, "n": len(self.saved)}

    def has_resources_for_trial(self, trial):
        return True


class _Runner(object):
    def __init__(self):
        self._trials = []
        self.trial_executor = _Executor()

    def get_trials(self):
        return list(self._trials)


def test_trial_scheduler_rescales_pauses_and_resumes(tune):
    from ray.tune import PlacementGroupFactory
    from ray.tune.experiment import Trial
    from ray.tune.schedulers import TrialScheduler
    from adaptdl_b200.ray import utils as ray_utils
    AdaptDLTrial, AdaptDLScheduler = tune.AdaptDLTrial, tune.AdaptDLScheduler
    assert issubclass(AdaptDLTrial, Trial)
    assert issubclass(AdaptDLScheduler, TrialScheduler)

    scheduler = AdaptDLScheduler()          # nodes from (fixture) ray.nodes()
    runner = _Runner()
    # a plain Tune trial is replaced by an elastic one on the default spot
    plain = Trial("train_fn", config={"lr": 0.1}, trial_id="t0",
                  placement_group_factory=PlacementGroupFactory(
                      [{"CPU": 1}]))
    runner._trials.append(plain)
    scheduler.on_trial_add(runner, plain)
    trial = runner._trials[0]
    assert isinstance(trial, AdaptDLTrial) and trial.trial_id == "t0"
    assert trial.job_id == "t0" and trial.rescale_count == 1
    assert ray_utils.pgf_to_num_replicas(trial.placement_group_factory) == 1
    assert trial._allocation_in_use() == ["10.0.0.1"]

    # results without hints: speedup = replicas capped at 1 -> stays put
    trial.status = Trial.RUNNING
    assert scheduler.on_trial_result(runner, trial, {}) == \
        TrialScheduler.CONTINUE
    assert scheduler.on_trial_result(runner, trial, {}) in (
        TrialScheduler.CONTINUE, TrialScheduler.STOP)

    # with noisy-gradient hints the policy wants more replicas: the trial is
    # cloned onto the new placement with an in-memory checkpoint and stopped
    trial = runner._trials[0]
    trial.status = Trial.RUNNING
    trial.last_result = {"sched_hints": _hints()}
    decisions = [scheduler.on_trial_result(runner, trial,
                                           trial.last_result)
                 for _ in range(2)]
    assert TrialScheduler.STOP in decisions
    clone = runner._trials[0]
    assert clone is not trial and isinstance(clone, AdaptDLTrial)
    assert clone.rescale_count == trial.rescale_count + 1
    replicas = ray_utils.pgf_to_num_replicas(clone.placement_group_factory)
    assert 1 < replicas <= 8
    assert clone.checkpoint == {"trial": "t0", "n": 1}
    assert runner.trial_executor.saved == [("t0", "memory")]
    assert clone._fetch_metrics() == _hints()      # hints survive the clone
    assert clone.job_info.max_replicas == 8        # 2 x maxProfiledReplicas

    # a paused trial with a pending allocation is resumed as a new clone
    clone.status = Trial.PAUSED
    scheduler._allocs[clone.job_id] = ["10.0.0.2"] * 2
    chosen = scheduler.choose_trial_to_run(runner)
    assert chosen is runner._trials[0] and chosen is not clone
    assert chosen._allocation_in_use() == ["10.0.0.2", "10.0.0.2"]
    assert scheduler.choose_trial_to_run(_Runner()) is None
    assert "AdaptDLScheduler" in scheduler.debug_string()


def test_allocation_round_trip_through_placement_groups(tune):
    from adaptdl_b200.ray import utils as ray_utils
    alloc = ["10.0.0.1", "10.0.0.1", "10.0.0.2"]
    pgf = ray_utils.allocation_to_pgf(alloc, {"CPU": 1, "GPU": 1})
    assert pgf.head_bundle_is_empty
    assert ray_utils.pgf_to_allocation(pgf) == alloc
    assert ray_utils.pgf_to_num_replicas(pgf) == 3
    assert ray_utils.unique_nodes(pgf.bundles[1:]) == 2


def test_reference_module_paths_of_the_ray_package(tune):
    """``from adaptdl_ray.tune.adaptdl_trial_sched import AdaptDLScheduler``
    (the import line of the reference's Tune examples and tutorial) and the
    other module paths of ``ray/adaptdl_ray`` resolve to the classes here."""
    from adaptdl_ray.tune.adaptdl_trial_sched import AdaptDLScheduler
    from adaptdl_ray.tune.adaptdl_trial import AdaptDLTrial
    from adaptdl_ray.tune.adaptdl_trainable import AdaptDLTrainableCreator
    from adaptdl_ray.tune import AdaptDLScheduler as again
    assert AdaptDLScheduler is tune.AdaptDLScheduler is again
    assert AdaptDLTrial is tune.AdaptDLTrial
    assert AdaptDLTrainableCreator is tune.AdaptDLTrainableCreator
    from adaptdl_ray.adaptdl import AdaptDLAllocator, AdaptDLJobMixin
    from adaptdl_ray.adaptdl.adaptdl_allocator import AdaptDLAllocator as a2
    from adaptdl_ray.adaptdl.adaptdl_job_mixin import AdaptDLJobMixin as m2
    from adaptdl_ray.adaptdl.utils import pgf_to_allocation  # noqa: F401
    from adaptdl_ray.adaptdl.config import default_device  # noqa: F401
    import adaptdl_b200.ray.allocator
    assert a2 is AdaptDLAllocator is adaptdl_b200.ray.allocator.AdaptDLAllocator
    assert m2 is AdaptDLJobMixin
    for name in ("controller", "worker", "launch_job", "optimizer", "utils"):
        __import__("adaptdl_ray.aws." + name)
