"""Driver of tests/test_ray.py::test_ray_aws_controller_runs_a_job...: the
Ray-on-AWS controller actor and its worker tasks on the in-process stand-in
for Ray (tests/fixtures/fake_ray), one real training script as the job.

    python ray_aws_job.py <script> <port> [--spot-after SECONDS] -- <argv>

Prints one JSON line: final status, worker generations, what was started."""
import asyncio
import json
import os
import sys
import threading
import time
from http.server import BaseHTTPRequestHandler, HTTPServer

import adaptdl_b200.torch  # noqa: F401 - signal handlers: main thread only
import ray
import adaptdl_b200.ray.aws.controller as controller

script, port = sys.argv[1], sys.argv[2]
rest = sys.argv[3:]
spot_after = None
if rest[:1] == ["--spot-after"]:
    spot_after, rest = float(rest[1]), rest[2:]
argv = rest[1:] if rest[:1] == ["--"] else rest

controller.MIN_RESCHEDULE_PERIOD_S = 0
os.environ["ADAPTDL_B200_RAY_CONTROLLER_PORT"] = port
os.environ["ADAPTDL_B200_RAY_CONTROLLER_HOST"] = "127.0.0.1"
# two worker nodes with room for one worker each; the controller's own node
# is never used for workers
ray._NODES[:] = [
    {"NodeManagerAddress": "127.0.0.1", "Alive": True,
     "Resources": {"CPU": 1.0}},
    {"NodeManagerAddress": "127.0.0.2", "Alive": True,
     "Resources": {"CPU": 1.0}},
    {"NodeManagerAddress": ray._actors.CONTROLLER_IP, "Alive": True,
     "Resources": {"CPU": 8.0}}]

if spot_after is not None:
    # EC2 metadata stand-in (MOCK=true makes the poller ask <ip>:8234):
    # 404 until the deadline, then the two-minute warning
    os.environ["MOCK"] = "true"
    t0 = time.time()

    class Metadata(BaseHTTPRequestHandler):
        def log_message(self, *args):
            pass

        def do_GET(self):
            due = time.time() - t0 >= spot_after
            body = json.dumps({"action": "terminate"}).encode() if due \
                else b"{}"
            self.send_response(200 if due else 404)
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)
    server = HTTPServer(("127.0.0.1", 8234), Metadata)
    threading.Thread(target=server.serve_forever, daemon=True).start()


async def main():
    Controller = controller.make_controller_actor()
    actor = Controller.options(name="AdaptDLController").remote(2, 5)
    actor.run_controller.remote()
    status = await actor.create_job.remote(
        worker_resources={"CPU": 1}, worker_port_offset=int(port) % 500,
        checkpoint_timeout=60, path=script, argv=argv)
    job = actor._obj._job
    print(json.dumps({
        "status": status, "generations": job.iteration,
        "had_checkpoint": job.checkpoint is not None,
        "hints": sorted(job.hints or {}),
        "terminating": sorted(actor._obj._terminating),
        "calls": [[kind, name] for kind, name, _ in ray._actors.CALLS],
        "resource_requests": len(ray.autoscaler.sdk.REQUESTS)}))

asyncio.run(main())
os._exit(0)            # daemon task threads may still be polling
