"""Elastic worker groups (the Ray-independent core of the Tune integration):
rescale = checkpoint in memory + new group on another allocation."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from adaptdl_b200.ray.tune import workers  # noqa: E402
import tune_workload  # noqa: E402

ENV = {"PYTHONPATH": os.pathsep.join([ROOT, os.path.join(ROOT, "tests")]),
       "CUDA_VISIBLE_DEVICES": "", "OMP_NUM_THREADS": "1"}


def test_replica_env_layout():
    env = workers.replica_env(2, ["a", "b", "b"], "10.0.0.1", 1234, 3, "t/1")
    assert env["ADAPTDL_REPLICA_RANK"] == "2"
    assert env["ADAPTDL_NUM_REPLICAS"] == "3"
    assert env["ADAPTDL_NUM_NODES"] == "2"
    assert env["ADAPTDL_LOCAL_RANK"] == "1"      # second replica on "b"
    assert env["ADAPTDL_NUM_RESTARTS"] == "3"
    assert workers.replica_env(1, ["a", "b", "b"], "h", 1, 0, "j")[
        "ADAPTDL_LOCAL_RANK"] == "0"


@pytest.mark.timeout(300)
def test_rescale_2_to_1_through_an_in_memory_checkpoint():
    spawner = workers.ProcessSpawner(extra_env=ENV)
    config = {"lr": 0.05, "epochs": 40, "pause": 0.1}
    group = workers.WorkerGroup(tune_workload.train_fn, config,
                                ["n0", "n0"], spawner)
    try:
        seen = [group.next_result(timeout=120) for _ in range(3)]
        assert [r["epoch"] for r in seen] == [0, 1, 2]
        assert all(r["replicas"] == 2 and r["restarts"] == 0 for r in seen)
        assert "sched_hints" in seen[0]
        snapshot = group.checkpoint(timeout=120)
    finally:
        group.shutdown()
    assert not group.finished
    assert any(path.startswith("checkpoint-0" + os.sep) for path in snapshot)
    assert all(isinstance(blob, bytes) for blob in snapshot.values())

    group = workers.WorkerGroup(tune_workload.train_fn, config, ["n1"],
                                spawner, checkpoint=snapshot, generation=1)
    try:
        first = group.next_result(timeout=120)
        # resumed, not restarted: continues after the last finished epoch
        assert first["epoch"] >= 3 and first["replicas"] == 1
        assert first["restarts"] == 1
        last = first
        while True:
            result = group.next_result(timeout=120)
            if result is None:
                break
            last = result
        assert group.finished and last["epoch"] == config["epochs"] - 1
        assert last["loss"] < 0.5
        assert group.next_result() is None
    finally:
        group.shutdown()


@pytest.mark.timeout(120)
def test_a_failing_replica_surfaces_its_traceback():
    group = workers.WorkerGroup(tune_workload.failing_fn, {}, ["n0"],
                                workers.ProcessSpawner(extra_env=ENV))
    with pytest.raises(workers.WorkerGroupError, match="boom"):
        group.next_result(timeout=60)


@pytest.mark.timeout(300)
def test_elastic_trial_save_and_restore_in_a_clone():
    from adaptdl_b200.ray.tune.trainable import CHECKPOINT_KEY, ElasticTrial
    spawner = workers.ProcessSpawner(extra_env=ENV)
    config = {"lr": 0.05, "epochs": 30, "pause": 0.1}
    trial = ElasticTrial(tune_workload.train_fn, config, ["n0"], spawner)
    try:
        first = trial.step()
        assert first["epoch"] == 0 and first["num_replicas"] == 1
        assert first["generation"] == 0 and first["done"] is False
        state = trial.save()
    finally:
        trial.stop()
    assert state["generation"] == 1 and state[CHECKPOINT_KEY]
    # the clone runs on two replicas and picks up where the original stopped
    clone = ElasticTrial(tune_workload.train_fn, config, ["n0", "n1"],
                         spawner)
    clone.restore(state)
    try:
        result = clone.step()
        assert result["epoch"] >= 1 and result["replicas"] == 2
        assert result["generation"] == 1
        while not result["done"]:
            last, result = result, clone.step()
        assert last["epoch"] == config["epochs"] - 1
        assert clone.step() == {"done": True}
    finally:
        clone.stop()


@pytest.mark.timeout(300)
def test_ray_actor_spawner_rescales_through_an_in_memory_checkpoint(tmp_path):
    """``RayActorSpawner`` (replicas as Ray actors, the production path of
    the Tune trainable) on the in-process stand-in for Ray's actor calls:
    same scenario as the process spawner above."""
    import json
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    env = {k: v for k, v in os.environ.items()
           if not k.startswith("ADAPTDL_")}
    env.update(ENV)
    env["PYTHONPATH"] = os.pathsep.join(
        [ROOT, os.path.join(here, "fixtures", "fake_ray"), here])
    proc = subprocess.run(
        [sys.executable, os.path.join(here, "ray_tune_actor_job.py")],
        env=env, cwd=str(tmp_path), stdout=subprocess.PIPE,
        stderr=subprocess.PIPE, text=True, timeout=280)
    assert proc.returncode == 0, proc.stderr[-3000:]
    out = json.loads([ln for ln in proc.stdout.splitlines()
                      if ln.startswith("{")][-1])
    assert out["first_epochs"] == [0, 1, 2] and not out["finished_first"]
    assert out["snapshot_files"][0].startswith("checkpoint-0")
    assert out["resumed_at"] >= 3 and out["restarts"] == 1
    assert out["finished"] and out["last_epoch"] == 11
    assert out["calls"].count(["actor", "Replica"]) == 4
    # ... and the Tune trainable of AdaptDLTrainableCreator on top of it
    trainable = out["trainable"]
    assert trainable["name"] == "AdaptDL_train_fn"
    assert trainable["first"] == [0, 1]
    assert trainable["generation_saved"] == 1
    assert trainable["resumed_generation"] == 1
    assert trainable["resumed"][0] >= 2 and trainable["resumed"][1] == 7
    assert len(trainable["resources"]) == 2        # driver + one replica
