"""``KubernetesCluster`` (the ``kubernetes_asyncio`` adapter of the
scheduler's cluster backend) against an in-memory stand-in for the client
library (``tests/fixtures/fake_k8s``): the real package is not installable
here, so the adapter's request / error translation is pinned down against
the call signatures it uses, and the controller is run on top of it."""

import asyncio
import os
import sys

import pytest

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                       "fixtures", "fake_k8s")


@pytest.fixture
def k8s(monkeypatch):
    try:
        import kubernetes_asyncio
        if "fixtures" not in (kubernetes_asyncio.__file__ or ""):
            pytest.skip("the real kubernetes_asyncio is installed")
    except ImportError:
        pass
    monkeypatch.syspath_prepend(FIXTURE)
    for name in [m for m in sys.modules
                 if m.split(".")[0] == "kubernetes_asyncio"]:
        monkeypatch.delitem(sys.modules, name)
    import kubernetes_asyncio.client as client
    client.reset()
    yield client
    client.reset()
    for name in [m for m in sys.modules
                 if m.split(".")[0] == "kubernetes_asyncio"]:
        sys.modules.pop(name, None)


def _pod(name, job="j1", containers=True):
    return {"metadata": {"name": name, "labels": {"adaptdl/job": job}},
            "spec": {"containers": [{"name": "main", "image": "x"}]
                     if containers else []}}


def test_adapter_translates_requests_and_errors(k8s):
    from adaptdl_b200.sched import config
    from adaptdl_b200.sched.kube import ApiError, KubernetesCluster, NotFound
    cluster = KubernetesCluster()

    async def scenario():
        k8s.STATE["nodes"]["n0"] = {"metadata": {"name": "n0"},
                                    "status": {"allocatable": {"cpu": "8"}}}
        k8s.STATE["jobs"][("ns", "j1")] = {
            "apiVersion": "{}/{}".format(config.GROUP, config.VERSION),
            "metadata": {"namespace": "ns", "name": "j1"}, "spec": {}}
        created = await cluster.create_pod("ns", _pod("p0"))
        assert created["metadata"]["namespace"] == "ns"
        assert created["status"]["phase"] == "Pending"
        # dry run: validated, not stored
        await cluster.create_pod("ns", _pod("p-dry"), dry_run=True)
        assert ("ns", "p-dry") not in k8s.STATE["pods"]
        with pytest.raises(ApiError) as err:
            await cluster.create_pod("ns", _pod("bad", containers=False))
        assert err.value.status == 422
        with pytest.raises(ApiError) as err:
            await cluster.create_pod("ns", _pod("p0"))
        assert err.value.status == 409
        pods = await cluster.list_pods("ns", label_selector="adaptdl/job=j1")
        assert [p["metadata"]["name"] for p in pods] == ["p0"]
        assert await cluster.list_pods("ns",
                                       label_selector="adaptdl/job=zz") == []
        assert len(await cluster.list_pods()) == 1
        assert (await cluster.get_job("ns", "j1"))["metadata"]["name"] == "j1"
        with pytest.raises(NotFound):
            await cluster.get_job("ns", "nope")
        patched = await cluster.patch_job_status(
            "ns", "j1", {"status": {"phase": "Running"}})
        assert patched["status"]["phase"] == "Running"
        assert await cluster.patch_job_status("ns", "gone", {}) is None
        assert [j["metadata"]["name"] for j in await cluster.list_jobs()] \
            == ["j1"]
        assert (await cluster.list_nodes())[0]["metadata"]["name"] == "n0"
        assert (await cluster.read_node("n0"))["status"]["allocatable"]
        await cluster.delete_pod("ns", "p0")
        await cluster.delete_pod("ns", "p0")        # 404 is not an error
        assert await cluster.list_pods("ns") == []
        # the watch multiplexes job and pod events as (kind, plain dict)
        seen = []

        async def consume():
            async for kind, obj in cluster.watch():
                seen.append((kind, obj["metadata"]["name"]))
                if len(seen) >= 3:
                    return
        await asyncio.wait_for(consume(), 5)
        assert ("job", "j1") in seen and ("pod", "p0") in seen
    asyncio.run(scenario())


def test_validator_without_a_cluster_backend_asks_the_api_server(k8s):
    """``Validator()`` (how the reference constructs it): the job's template
    is checked with a dry-run PodTemplate creation and the API server's own
    message comes back in the admission response."""
    from adaptdl_b200.sched.validator import Validator, verdict_for_create
    validator = Validator()
    assert validator._core_api is validator._cluster.core_api

    async def scenario():
        bad = {"spec": {"template": {"spec": {"containers": []}}}}
        verdict = await verdict_for_create(validator._cluster, "ns", bad)
        assert not verdict["allowed"]
        assert verdict["status"]["reason"] == "Invalid"
        assert verdict["status"]["message"].startswith("PodTemplate is")
        good = {"spec": {"minReplicas": 1, "maxReplicas": 2, "template": {
            "spec": {"containers": [{"name": "main", "image": "x"}]}}}}
        assert (await verdict_for_create(validator._cluster, "ns",
                                         good))["allowed"]
        good["spec"]["maxReplicas"] = 0
        assert not (await verdict_for_create(validator._cluster, "ns",
                                             good))["allowed"]
        assert not k8s.STATE["pods"]          # nothing was persisted
    asyncio.run(scenario())


def _seed_file(tmp_path):
    import json
    job = {"apiVersion": "adaptdl.petuum.com/v1", "kind": "AdaptDLJob",
           "metadata": {"name": "j1", "namespace": "ns", "uid": "u1",
                        "creationTimestamp": "2026-01-01T00:00:00Z"},
           "spec": {"template": {"spec": {"containers": [
               {"name": "main", "image": "x", "resources": {
                   "limits": {"nvidia.com/gpu": 1}}}]}}},
           "status": {"phase": "Running", "group": 0,
                      "allocation": ["n0", "n0"], "replicas": 2}}
    pods = [{"metadata": {"name": "j1-{}".format(rank), "namespace": "ns",
                          "labels": {"adaptdl/job": "j1"},
                          "annotations": {"adaptdl/group": "0",
                                          "adaptdl/replicas": "2",
                                          "adaptdl/rank": str(rank)}},
             "spec": {"nodeName": "n0", "containers": [
                 {"name": "main", "image": "x"}]},
             "status": {"phase": "Running",
                        "podIP": "10.0.0.{}".format(rank + 1)}}
            for rank in range(2)]
    node = {"metadata": {"name": "n0", "labels": {}},
            "status": {"allocatable": {"nvidia.com/gpu": "8", "pods": "110",
                                       "cpu": "64", "memory": "512Gi"}},
            "spec": {}}
    path = tmp_path / "state.json"
    path.write_text(json.dumps({"nodes": [node], "jobs": [job],
                                "pods": pods}))
    return str(path)


@pytest.mark.parametrize("module", ["adaptdl_sched", "adaptdl_sched.allocator",
                                    "adaptdl_sched.supervisor"])
def test_scheduler_containers_start_with_the_reference_commands(tmp_path,
                                                                module):
    """The three long-running scheduler processes, started the way the
    reference's chart starts them (``python -m adaptdl_sched[...]``), come up
    against the stand-in API server, stay up, and the supervisor answers a
    replica's rendezvous and hints requests."""
    import json
    import signal
    import socket
    import subprocess
    import time
    import urllib.request
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([root, FIXTURE]),
               FAKE_K8S_STATE=_seed_file(tmp_path),
               ADAPTDL_SUPERVISOR_SERVICE_PORT=str(port),
               ADAPTDL_NAMESPACE="ns")
    log = open(str(tmp_path / "out.log"), "w")
    proc = subprocess.Popen([sys.executable, "-m", module], env=env,
                            cwd=str(tmp_path), stdout=log,
                            stderr=subprocess.STDOUT)
    try:
        if module.endswith("supervisor"):
            base = "http://127.0.0.1:{}".format(port)
            deadline = time.time() + 60
            while True:
                try:
                    urllib.request.urlopen(base + "/healthz", timeout=2)
                    break
                except OSError:
                    assert proc.poll() is None and time.time() < deadline, \
                        open(log.name).read()[-3000:]
                    time.sleep(0.2)
            with urllib.request.urlopen(
                    base + "/discover/ns/j1/0?timeout=5") as resp:
                assert json.loads(resp.read()) == ["10.0.0.1", "10.0.0.2"]
            body = json.dumps({"initBatchSize": 128}).encode()
            req = urllib.request.Request(base + "/hints/ns/j1", data=body,
                                         method="PUT")
            with urllib.request.urlopen(req) as resp:
                assert resp.status == 200
        else:
            time.sleep(6.0)           # imports + a few loop iterations
        assert proc.poll() is None, open(log.name).read()[-3000:]
    finally:
        proc.send_signal(signal.SIGTERM)
        try:
            proc.wait(timeout=20)
        except subprocess.TimeoutExpired:
            proc.kill()
        log.close()
    text = open(log.name).read()
    assert "Traceback" not in text, text[-3000:]


def test_controller_process_creates_the_pods_of_an_allocated_job(tmp_path):
    """``python -m adaptdl_sched`` (the controller container) against the
    stand-in API server: a job that has an allocation but no pods gets one
    pod per replica, with the replica's rank / world size in its
    environment (the contract of controller.py:310-420 in the reference)."""
    import json
    import signal
    import subprocess
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    seed = json.loads(open(_seed_file(tmp_path)).read())
    seed["pods"] = []
    seed["jobs"][0]["status"] = {"phase": "Pending", "group": 0,
                                 "allocation": ["n0", "n0"]}
    (tmp_path / "state.json").write_text(json.dumps(seed))
    dump = tmp_path / "pods.json"
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([root, FIXTURE]),
               FAKE_K8S_STATE=str(tmp_path / "state.json"),
               FAKE_K8S_DUMP=str(dump), ADAPTDL_NAMESPACE="ns")
    log = open(str(tmp_path / "out.log"), "w")
    proc = subprocess.Popen([sys.executable, "-m", "adaptdl_sched"], env=env,
                            cwd=str(tmp_path), stdout=log,
                            stderr=subprocess.STDOUT)
    try:
        deadline = time.time() + 90
        pods = []
        while time.time() < deadline and len(pods) < 2:
            assert proc.poll() is None, open(log.name).read()[-3000:]
            if dump.exists():
                pods = json.loads(dump.read_text())["pods"]
            time.sleep(0.3)
        assert len(pods) == 2, open(log.name).read()[-3000:]
    finally:
        proc.send_signal(signal.SIGTERM)
        try:
            proc.wait(timeout=20)
        except subprocess.TimeoutExpired:
            proc.kill()
        log.close()
    ranks = []
    for pod in pods:
        assert pod["metadata"]["labels"]["adaptdl/job"] == "j1"
        assert "n0" in json.dumps(pod["spec"])        # pinned to its node
        env_vars = {e["name"]: e.get("value")
                    for e in pod["spec"]["containers"][0]["env"]}
        assert env_vars["ADAPTDL_NUM_REPLICAS"] == "2"
        ranks.append(env_vars["ADAPTDL_REPLICA_RANK"])
    assert sorted(ranks) == ["0", "1"]


def test_allocator_process_gives_a_new_job_its_first_allocation(tmp_path):
    """``python -m adaptdl_sched.allocator`` against the stand-in API server:
    a job without an allocation is placed right away (the fast path that
    does not wait for the next optimisation cycle; allocator.py:48-80 in the
    reference) and the decision is written to the job's status."""
    import json
    import signal
    import subprocess
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    seed = json.loads(open(_seed_file(tmp_path)).read())
    seed["pods"] = []
    seed["jobs"][0]["status"] = {"phase": "Pending"}
    (tmp_path / "state.json").write_text(json.dumps(seed))
    dump = tmp_path / "state-now.json"
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([root, FIXTURE]),
               FAKE_K8S_STATE=str(tmp_path / "state.json"),
               FAKE_K8S_DUMP=str(dump), ADAPTDL_NAMESPACE="ns")
    log = open(str(tmp_path / "out.log"), "w")
    proc = subprocess.Popen([sys.executable, "-m", "adaptdl_sched.allocator"],
                            env=env, cwd=str(tmp_path), stdout=log,
                            stderr=subprocess.STDOUT)
    try:
        deadline = time.time() + 90
        allocation = None
        while time.time() < deadline and not allocation:
            assert proc.poll() is None, open(log.name).read()[-3000:]
            if dump.exists():
                job = json.loads(dump.read_text())["jobs"][0]
                allocation = (job.get("status") or {}).get("allocation")
            time.sleep(0.3)
        assert allocation and set(allocation) == {"n0"}, \
            open(log.name).read()[-3000:]
    finally:
        proc.send_signal(signal.SIGTERM)
        try:
            proc.wait(timeout=20)
        except subprocess.TimeoutExpired:
            proc.kill()
        log.close()


def test_validator_process_uses_cluster_credentials_and_rejects_bad_jobs(
        tmp_path):
    """``python -m adaptdl_sched.validator`` (the webhook container): loads
    the in-cluster credentials before talking to the API server, answers
    ``/healthz``, and turns the API server's dry-run verdict into an
    admission response."""
    import json
    import signal
    import socket
    import subprocess
    import time
    import urllib.request
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    marker = tmp_path / "config-loaded"
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([root, FIXTURE]),
               FAKE_K8S_CONFIG_MARKER=str(marker))
    log = open(str(tmp_path / "out.log"), "w")
    proc = subprocess.Popen(
        [sys.executable, "-m", "adaptdl_sched.validator", "--host",
         "127.0.0.1", "--port", str(port)], env=env, cwd=str(tmp_path),
        stdout=log, stderr=subprocess.STDOUT)
    base = "http://127.0.0.1:{}".format(port)

    def review(spec):
        body = json.dumps({"request": {
            "uid": "u1", "operation": "CREATE", "namespace": "ns",
            "object": {"spec": spec}}}).encode()
        req = urllib.request.Request(
            base + "/validate", data=body, method="POST",
            headers={"Content-Type": "application/json"})
        with urllib.request.urlopen(req, timeout=10) as resp:
            return json.loads(resp.read())["response"]
    try:
        deadline = time.time() + 60
        while True:
            try:
                urllib.request.urlopen(base + "/healthz", timeout=2)
                break
            except OSError:
                assert proc.poll() is None and time.time() < deadline, \
                    open(log.name).read()[-3000:]
                time.sleep(0.2)
        assert marker.read_text() == "incluster"
        bad = review({"template": {"spec": {"containers": []}}})
        assert bad["uid"] == "u1" and not bad["allowed"]
        assert bad["status"]["code"] == 422
        good = review({"minReplicas": 1, "maxReplicas": 4, "template": {
            "spec": {"containers": [{"name": "main", "image": "x"}]}}})
        assert good["allowed"], good
    finally:
        proc.send_signal(signal.SIGTERM)
        try:
            proc.wait(timeout=20)
        except subprocess.TimeoutExpired:
            proc.kill()
        log.close()
