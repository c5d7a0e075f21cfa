"""``python -m adaptdl_b200.launch``: the replicas of a node pod."""
import os
import signal
import subprocess
import sys
import time

from adaptdl_b200 import launch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_replica_environments():
    envs = launch.replica_environments(
        {"ADAPTDL_LOCAL_REPLICAS": "3", "ADAPTDL_REPLICA_RANK": "4",
         "ADAPTDL_NUM_REPLICAS": "8", "OTHER": "x"})
    assert [e["ADAPTDL_REPLICA_RANK"] for e in envs] == ["4", "5", "6"]
    assert [e["ADAPTDL_LOCAL_RANK"] for e in envs] == ["0", "1", "2"]
    assert all(e["ADAPTDL_NUM_REPLICAS"] == "8" and e["OTHER"] == "x"
               for e in envs)
    # outside a node pod: one replica, rank as given (default 0)
    solo = launch.replica_environments({})
    assert len(solo) == 1 and solo[0]["ADAPTDL_REPLICA_RANK"] == "0"


def test_exit_code_rules():
    assert launch.exit_code([0, 0]) == 0
    assert launch.exit_code([143, 143]) == 143
    assert launch.exit_code([0, 143]) == 143
    assert launch.exit_code([143, 3, 143]) == 3
    assert launch.exit_code([-9, 0]) == 137          # killed by a signal
    assert launch.exit_code([-15, 143]) == 143       # SIGTERM before the handler


def _launch(code, tmp_path, replicas=3, **popen):
    env = dict(os.environ, PYTHONPATH=ROOT, ADAPTDL_LOCAL_REPLICAS=str(replicas),
               ADAPTDL_REPLICA_RANK="4", OUT=str(tmp_path))
    return subprocess.Popen([sys.executable, "-m", "adaptdl_b200.launch",
                             "-c", code], env=env, **popen)


WRITE = ("import os; open(os.path.join(os.environ['OUT'], "
         "os.environ['ADAPTDL_REPLICA_RANK']), 'w').write("
         "os.environ['ADAPTDL_LOCAL_RANK'])")


def test_runs_one_process_per_local_replica(tmp_path):
    assert _launch(WRITE, tmp_path).wait(60) == 0
    assert sorted(os.listdir(tmp_path)) == ["4", "5", "6"]
    assert [open(os.path.join(tmp_path, r)).read() for r in "456"] == \
        ["0", "1", "2"]


def test_one_failure_stops_the_group(tmp_path):
    code = ("import os, sys, time\n"
            "if os.environ['ADAPTDL_LOCAL_RANK'] == '1': sys.exit(3)\n"
            "time.sleep(60)")
    began = time.time()
    assert _launch(code, tmp_path).wait(60) == 3
    assert time.time() - began < 30          # did not wait for the sleepers


def test_sigterm_is_forwarded_and_preemption_reported(tmp_path):
    code = ("import os, signal, sys, time\n"
            "signal.signal(signal.SIGTERM, lambda *a: sys.exit(143))\n"
            + WRITE + "\n"
            "time.sleep(60)")
    proc = _launch(code, tmp_path, replicas=2)
    deadline = time.time() + 30
    while len(os.listdir(tmp_path)) < 2 and time.time() < deadline:
        time.sleep(0.1)
    proc.send_signal(signal.SIGTERM)
    assert proc.wait(30) == 143


def test_two_replicas_of_a_real_job_in_one_container(tmp_path):
    """What a node pod runs: the launcher + an unmodified training script;
    the two replicas find each other through the ADAPTDL_* variables."""
    from adaptdl_b200.utils import pick_unused_port
    script = os.path.join(ROOT, "examples", "linear_regression", "main.py")
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="",
               OMP_NUM_THREADS="1", ADAPTDL_LOCAL_REPLICAS="2",
               ADAPTDL_REPLICA_RANK="0", ADAPTDL_NUM_REPLICAS="2",
               ADAPTDL_NUM_NODES="1", ADAPTDL_MASTER_ADDR="127.0.0.1",
               ADAPTDL_MASTER_PORT=str(pick_unused_port()),
               ADAPTDL_CHECKPOINT_PATH=str(tmp_path), ADAPTDL_JOB_ID="ns/j")
    for stale in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(stale, None)
    done = subprocess.run(
        [sys.executable, "-m", "adaptdl_b200.launch", script, "--epochs",
         "2", "--size", "512"], env=env, timeout=240,
        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert done.returncode == 0, done.stdout[-2000:]


def test_replicas_flag_runs_a_whole_job_by_hand(tmp_path):
    """``--replicas N`` without a scheduler: the launcher supplies the
    job-level variables itself (N replicas, one node, a free port)."""
    script = os.path.join(ROOT, "examples", "linear_regression", "main.py")
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="",
               OMP_NUM_THREADS="1", ADAPTDL_CHECKPOINT_PATH=str(tmp_path))
    for stale in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "ADAPTDL_NUM_REPLICAS",
                  "ADAPTDL_LOCAL_REPLICAS", "ADAPTDL_REPLICA_RANK",
                  "ADAPTDL_MASTER_PORT"):
        env.pop(stale, None)
    done = subprocess.run(
        [sys.executable, "-m", "adaptdl_b200.launch", "--replicas", "2",
         script, "--epochs", "2", "--size", "512"], env=env, timeout=240,
        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert done.returncode == 0, done.stdout[-2000:]
    envs = launch.replica_environments({"ADAPTDL_LOCAL_REPLICAS": "3"})
    assert {e["ADAPTDL_NUM_REPLICAS"] for e in envs} == {"3"}
    assert len({e["ADAPTDL_MASTER_PORT"] for e in envs}) == 1
