"""Ray integration: the Ray-independent logic (Ray itself is optional)."""
import os
import threading
from http.server import BaseHTTPRequestHandler, HTTPServer

import pytest

from adaptdl_b200.ray import have_ray, utils as ray_utils
from adaptdl_b200.ray.allocator import AdaptDLAllocator
from adaptdl_b200.ray.aws import optimizer
from adaptdl_b200.ray.aws.controller import (JobState, cluster_ready,
                                             speedup_from_hints,
                                             trim_allocation)
from adaptdl_b200.ray.aws.utils import (Status, checkpoint_obj_to_dir,
                                        serialize_checkpoint)
from adaptdl_b200.ray.aws.worker import (poll_spot_termination, run_script,
                                         worker_environment)
from adaptdl_b200.ray.job_mixin import AdaptDLJobMixin
from adaptdl_b200.sched.policy import PolluxPolicy, SpeedupFunction

HINTS = {"initBatchSize": 128, "maxBatchSize": 1280,
         "localBszBounds": [64, 256], "maxProfiledReplicas": 4,
         "gradientAccumulation": False,
         "gradParams": {"norm": 0.00136, "var": 0.000502},
         "perfParams": dict(alpha_c=0.121, beta_c=0.00568, alpha_n=0.0236,
                            beta_n=0.00634, alpha_r=0.0118, beta_r=0.00317,
                            gamma=1.14)}


def test_bundle_allocation_roundtrip():
    alloc = ["10.0.0.1", "10.0.0.1", "10.0.0.2", "virtual-3"]
    bundles = ray_utils.allocation_to_bundles(alloc, {"CPU": 1, "GPU": 1})
    assert bundles[0] == {"CPU": 1, "GPU": 1, "node:10.0.0.1": 0.01}
    assert "node:virtual-3" not in bundles[3]
    assert ray_utils.bundles_to_allocation(bundles) == alloc
    assert ray_utils.unique_nodes(bundles) == 3


def test_checkpoint_serialisation_roundtrip(tmp_path):
    src = tmp_path / "src"
    (src / "checkpoint-0").mkdir(parents=True)
    (src / "checkpoint-0" / "a").write_bytes(b"123")
    (src / "checkpoint-0" / "b").write_bytes(b"\x00\x01")
    obj = serialize_checkpoint(str(src))
    assert obj == {"checkpoint-0/a": b"123", "checkpoint-0/b": b"\x00\x01"}
    dst = tmp_path / "dst"
    checkpoint_obj_to_dir(str(dst), obj)
    assert (dst / "checkpoint-0" / "b").read_bytes() == b"\x00\x01"


def test_single_job_optimizer_hysteresis():
    fn = speedup_from_hints(HINTS)
    nodes = [("10.0.0.{}".format(i), {"CPU": 8, "GPU": 4}) for i in range(2)]
    res = {"CPU": 1, "GPU": 1}
    assert optimizer.optimize(None, None, nodes, res, 4, 1) == \
        ["adaptdl_virtual_node_0"]
    alloc = optimizer.greedy_allocation(nodes, res, 3)
    assert alloc[:8] == ["10.0.0.0"] * 4 + ["10.0.0.1"] * 4
    assert alloc[8:] == ["adaptdl_virtual_node_0"]
    grown = optimizer.optimize(HINTS, fn, nodes, res, 2, 1)
    assert 1 < len(grown) <= 8
    again = optimizer.optimize(HINTS, fn, nodes, res, 2, len(grown))
    assert len(again) == len(grown)          # already at the optimum


def test_cluster_ready_and_trim():
    res = {"CPU": 1, "GPU": 1}
    nodes = {"a": {"CPU": 4, "GPU": 2}, "b": {"CPU": 4, "GPU": 1}}
    assert cluster_ready(["a", "a", "b"], nodes, res) == (True, 3)
    ready, n = cluster_ready(["a", "a", "adaptdl_virtual_node_0",
                              "adaptdl_virtual_node_1"], nodes, res)
    assert (ready, n) == (False, 3)
    assert trim_allocation(["adaptdl_virtual_node_0", "a", "a"], 2) == \
        ["a", "a"]
    job = JobState(res, 60, {"path": "x.py"})
    job.register_hints(dict(HINTS, junk=1))
    assert "junk" not in job.hints
    job.register_status(Status.SUCCEEDED.value)
    job.register_status(Status.FAILED.value)
    assert job.status is Status.SUCCEEDED


def test_spot_termination_poller():
    state = {"calls": 0}

    class Handler(BaseHTTPRequestHandler):
        def log_message(self, *a):
            pass

        def do_GET(self):
            state["calls"] += 1
            if state["calls"] < 3:
                self.send_response(404)
                self.end_headers()
                return
            body = b'{"action": "terminate", "time": "soon"}'
            self.send_response(200)
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)
    server = HTTPServer(("127.0.0.1", 0), Handler)
    threading.Thread(target=server.serve_forever, daemon=True).start()
    endpoint = "127.0.0.1:{}".format(server.server_port)
    assert poll_spot_termination(endpoint, timeout=10, period=0.01) == \
        "terminate"
    assert state["calls"] == 3
    state["calls"] = -100
    assert poll_spot_termination(endpoint, timeout=0.05, period=0.01) is None
    server.shutdown()


def test_worker_environment_and_script_runner(tmp_path):
    env = worker_environment("ns/job", "uid1", 2, 4, 3, "http://c:8080",
                             offset=10, base_dir=str(tmp_path))
    assert env["ADAPTDL_MASTER_PORT"] == "47013"
    assert env["ADAPTDL_REPLICA_RANK"] == "2"
    assert env["ADAPTDL_NUM_REPLICAS"] == "4"
    assert env["ADAPTDL_NUM_RESTARTS"] == "3"
    script = tmp_path / "train.py"
    script.write_text(
        "import os, sys\n"
        "ck = os.environ['ADAPTDL_CHECKPOINT_PATH']\n"
        "prev = os.path.exists(os.path.join(ck, 'state'))\n"
        "open(os.path.join(ck, 'state'), 'w').write(sys.argv[1])\n"
        "sys.exit(0 if prev else 143)\n")
    saved, saved_argv = dict(os.environ), list(__import__("sys").argv)
    try:
        status, ckpt = run_script(str(script), ["first"], env)
        assert status is Status.RUNNING and ckpt == {"state": b"first"}
        status, ckpt = run_script(str(script), ["second"], env, ckpt)
        assert status is Status.SUCCEEDED and ckpt is None
    finally:
        os.environ.clear()
        os.environ.update(saved)
        __import__("sys").argv = saved_argv


def test_ray_allocator_with_job_mixin():
    class Job(AdaptDLJobMixin):
        rescale_resources = {"CPU": 1, "GPU": 1}

        def __init__(self, job_id, alloc):
            super().__init__(job_id=job_id)
            self._alloc = alloc

        def _allocation_in_use(self):
            return self._alloc
    nodes = {"n0": {"CPU": 8, "GPU": 4}, "n1": {"CPU": 8, "GPU": 4}}
    allocator = AdaptDLAllocator(nodes, PolluxPolicy(generations=20, seed=0))
    assert allocator.default_allocation(2) == ["n0", "n0"]
    jobs = [Job("a", ["n0"]), Job("b", [])]
    jobs[0].update_hints(HINTS)
    info = jobs[0].job_info
    assert isinstance(info.speedup_fn, SpeedupFunction)
    assert info.max_replicas == 8 and jobs[1].job_info.max_replicas == 1
    changed, desired = allocator.allocate(jobs)
    assert desired >= 1
    assert set(changed) <= {"a", "b"}
    for alloc in changed.values():
        assert set(alloc) <= set(nodes)


@pytest.mark.skipif(have_ray(), reason="only meaningful without ray")
def test_ray_missing_is_reported_clearly():
    from adaptdl_b200.ray import require_ray
    with pytest.raises(ImportError):
        require_ray()
    with pytest.raises(ImportError):
        from adaptdl_b200.ray.tune import AdaptDLScheduler  # noqa: F401


def test_launch_job_parses_the_reference_command_line(tmp_path):
    """Every flag of the reference's ``adaptdl_on_ray_aws`` (ray/adaptdl_ray/
    aws/launch_job.py) is accepted; parsing ends at the missing script."""
    from adaptdl_b200.ray.aws.launch_job import main
    with pytest.raises(SystemExit) as exc:
        main(["-f", "train.py", "-u", "auto", "-m", "4", "--gpus", "1",
              "--cpus", "2", "--port-offset", "100", "-d", str(tmp_path),
              "--checkpoint-timeout", "30", "--cluster-rescale-timeout", "5",
              "--", "--epochs", "1"])
    assert "train.py not found" in str(exc.value)


def _run_ray_aws_job(tmp_path, *extra, epochs="30"):
    import json
    import socket
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = {k: v for k, v in os.environ.items()
           if not k.startswith("ADAPTDL_")}
    env["PYTHONPATH"] = os.pathsep.join(
        [root, os.path.join(here, "fixtures", "fake_ray")])
    env["OMP_NUM_THREADS"] = "1"
    script = os.path.join(root, "examples", "linear_regression", "main.py")
    proc = subprocess.run(
        [sys.executable, os.path.join(here, "ray_aws_job.py"), script,
         str(port)] + list(extra) + ["--", "--epochs", epochs, "--size",
                                     "2000"],
        env=env, cwd=str(tmp_path), stdout=subprocess.PIPE,
        stderr=subprocess.PIPE, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-3000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert lines, proc.stdout[-2000:] + proc.stderr[-2000:]
    return json.loads(lines[-1]), proc.stderr


def test_ray_aws_controller_runs_a_job_to_completion(tmp_path):
    """The controller actor end to end on the in-process stand-in for Ray
    (reference scenario: ray/adaptdl_ray/aws/test_controller.py): worker
    task started on the first allocation, the training script finds rank 0
    through the controller's ``/discover``, its hints arrive at ``/hints``,
    the job ends ``SUCCEEDED``."""
    out, _ = _run_ray_aws_job(tmp_path)
    assert out["status"] == 1, out                  # Status.SUCCEEDED
    assert out["generations"] >= 1
    assert ["actor", "Controller"] in out["calls"]
    assert ["task", "run_adaptdl"] in out["calls"]
    assert ["task", "listen_for_spot_termination"] in out["calls"]
    assert out["resource_requests"] >= 1            # autoscaler was asked


def test_ray_aws_controller_moves_the_job_off_a_spot_node(tmp_path):
    """A spot-termination notice on the worker's node: the node is excluded,
    the worker is cancelled (SIGINT -> checkpoint -> exit 143), its
    checkpoint travels through the object store, and the next generation
    resumes from it on the other node and finishes."""
    out, err = _run_ray_aws_job(tmp_path, "--spot-after", "4", epochs="1500")
    assert out["status"] == 1, (out, err[-2000:])
    assert out["terminating"] == ["127.0.0.1"], out
    assert out["generations"] >= 2 and out["had_checkpoint"], out
    assert ["cancel", None] in out["calls"]


def test_adaptdl_on_ray_aws_command_runs_a_job(tmp_path):
    """The command a user types (reference: ``adaptdl_on_ray_aws -f
    train.py -m 2 --cpus 1 -- args``) from argument parsing to the exit
    code, on the in-process stand-in for Ray."""
    import socket
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    workdir = os.path.join(root, "examples", "linear_regression")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = {k: v for k, v in os.environ.items()
           if not k.startswith("ADAPTDL_")}
    env["PYTHONPATH"] = os.pathsep.join(
        [root, os.path.join(here, "fixtures", "fake_ray")])
    env["OMP_NUM_THREADS"] = "1"
    proc = subprocess.run(
        [sys.executable, os.path.join(here, "ray_aws_launch.py"), str(port),
         "-f", "main.py", "-d", workdir, "-m", "1", "--cpus", "1",
         "--port-offset", str(port % 400), "--cluster-rescale-timeout", "5",
         "--", "--epochs", "20", "--size", "2000"],
        # Ray's runtime_env working_dir: the workers run inside it
        env=env, cwd=workdir, stdout=subprocess.PIPE,
        stderr=subprocess.PIPE, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-3000:]
    assert "job finished with status 1" in proc.stdout, proc.stdout
    assert "EXIT 0" in proc.stdout
    assert "['actor', 'Controller']" in proc.stdout
