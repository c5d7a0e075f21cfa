"""Driver of tests/test_ray.py::test_adaptdl_on_ray_aws_command...: the
``adaptdl_on_ray_aws`` entry point itself (``launch_job.main``) on the
in-process stand-in for Ray: argument parsing, ``ray.init``, the named
controller actor, ``create_job`` and the exit code."""
import os
import sys

import adaptdl_b200.torch  # noqa: F401 - signal handlers: main thread only
import ray
import adaptdl_b200.ray.aws.controller as controller
from adaptdl_b200.ray.aws.launch_job import main

controller.MIN_RESCHEDULE_PERIOD_S = 0
os.environ["ADAPTDL_B200_RAY_CONTROLLER_PORT"] = sys.argv[1]
os.environ["ADAPTDL_B200_RAY_CONTROLLER_HOST"] = "127.0.0.1"
ray._NODES[:] = [
    {"NodeManagerAddress": "127.0.0.1", "Alive": True,
     "Resources": {"CPU": 1.0}},
    {"NodeManagerAddress": ray._actors.CONTROLLER_IP, "Alive": True,
     "Resources": {"CPU": 8.0}}]
code = main(sys.argv[2:])
print("EXIT", code, [[k, n] for k, n, _ in ray._actors.CALLS][:3])
sys.stdout.flush()
os._exit(code)
