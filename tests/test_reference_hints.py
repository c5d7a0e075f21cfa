"""Supervisor protocol against the unmodified reference: rendezvous through
``GET /discover/<job>/<group>`` and the scheduling hints the trainer PUTs to
``/hints/<job>``. A capturing fake supervisor serves the same
2-replica job under both implementations; the requests must be the same."""

import json
import os
import subprocess
import sys
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

import pytest

from adaptdl_b200.utils import pick_unused_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
JOB = os.path.join(ROOT, "tests", "hints_job.py")

pytestmark = pytest.mark.skipif(
    not os.path.isdir(os.path.join(REF, "adaptdl")),
    reason="reference package not installed (baseline/install_reference.sh)")


class Supervisor(object):
    def __init__(self, replicas):
        outer = self
        self.requests = []           # (method, path, body)

        class Handler(BaseHTTPRequestHandler):
            def log_message(self, *args):
                pass

            def _reply(self, body):
                data = json.dumps(body).encode()
                self.send_response(200)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(data)))
                self.end_headers()
                self.wfile.write(data)

            def do_GET(self):
                outer.requests.append(("GET", self.path, None))
                self._reply(["127.0.0.1"] * replicas)

            def do_PUT(self):
                size = int(self.headers.get("Content-Length", 0))
                outer.requests.append(
                    ("PUT", self.path, json.loads(self.rfile.read(size))))
                self._reply({})
        self.server = ThreadingHTTPServer(("127.0.0.1", 0), Handler)
        self.url = "http://127.0.0.1:{}".format(self.server.server_port)
        threading.Thread(target=self.server.serve_forever,
                         daemon=True).start()

    def close(self):
        self.server.shutdown()


def _run(impl, workdir, replicas=2):
    supervisor = Supervisor(replicas)
    base = {k: v for k, v in os.environ.items()
            if not k.startswith("ADAPTDL_") and k != "PYTHONPATH"}
    if impl == "reference":
        path = [REF, os.path.join(ROOT, "baseline", "shims")]
        base["TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD"] = "1"
    else:
        path = [ROOT]
    port = pick_unused_port()
    procs = []
    try:
        for rank in range(replicas):
            env = dict(base, PYTHONPATH=os.pathsep.join(path),
                       OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="",
                       ADAPTDL_JOB_ID="team/job-7",
                       ADAPTDL_SUPERVISOR_URL=supervisor.url,
                       ADAPTDL_MASTER_PORT=str(port),
                       ADAPTDL_NUM_REPLICAS=str(replicas),
                       ADAPTDL_REPLICA_RANK=str(rank),
                       ADAPTDL_NUM_NODES="1", ADAPTDL_NUM_RESTARTS="3")
            procs.append(subprocess.Popen(
                [sys.executable, JOB], env=env, stdout=subprocess.PIPE,
                stderr=subprocess.PIPE, text=True, cwd=str(workdir)))
        outs = [p.communicate(timeout=300) for p in procs]
        for p, (_, err) in zip(procs, outs):
            assert p.returncode == 0, err[-3000:]
    finally:
        supervisor.close()
    grad = [line.split()[1:] for line in outs[0][0].splitlines()
            if line.startswith("GRAD ")][0]
    return supervisor.requests, [float(v) for v in grad]


def test_discovery_and_hints_requests_match_the_reference(tmp_path):
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(2) as pool:          # the two arms side by side
        theirs = pool.submit(_run, "reference", tmp_path)
        ours = pool.submit(_run, "own", tmp_path)
        (theirs, their_grad), (ours, our_grad) = theirs.result(), \
            ours.result()

    def discover(requests):
        return sorted(path for method, path, _ in requests
                      if method == "GET")
    # every replica asks for the same job / restart-group path
    assert set(discover(ours)) == set(discover(theirs)) == \
        {"/discover/team/job-7/3"}
    assert len(discover(ours)) == len(discover(theirs)) == 2

    def hints(requests):
        puts = [(path, body) for method, path, body in requests
                if method == "PUT"]
        assert puts and all(path == "/hints/team/job-7" for path, _ in puts)
        return puts[-1][1]
    mine, ref = hints(ours), hints(theirs)
    assert set(mine) == set(ref)
    for key in ("initBatchSize", "maxBatchSize", "localBszBounds",
                "gradientAccumulation", "maxProfiledReplicas"):
        assert mine[key] == ref[key], key
    assert mine["initBatchSize"] == 64 and mine["maxBatchSize"] == 1024
    assert list(mine["localBszBounds"]) == [16, 256]
    assert set(mine["perfParams"]) == set(ref["perfParams"])
    assert all(isinstance(v, float) and v > 0
               for v in mine["perfParams"].values())
    # the gradient statistics are deterministic and end at the same numbers
    # in both; this framework reports the current ones, the reference those
    # of the step before (its hint is refreshed in a backward callback that
    # runs ahead of the statistics update, like its ``gain`` attribute)
    assert their_grad == pytest.approx(our_grad, rel=1e-4)
    assert [mine["gradParams"]["norm"], mine["gradParams"]["var"]] == \
        pytest.approx(our_grad, rel=1e-6)
    assert ref["gradParams"]["norm"] == pytest.approx(our_grad[0], rel=0.1)
