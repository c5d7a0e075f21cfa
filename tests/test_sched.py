"""Cluster scheduler on the in-memory cluster backend: resources parsing,
controller state machine, allocator cycle, supervisor and validator HTTP
endpoints, cluster expander."""
import asyncio
import json

from adaptdl_b200.sched import resources
from adaptdl_b200.sched.allocator import AdaptDLAllocator, job_info_from
from adaptdl_b200.sched.cluster_expander import ClusterExpander
from adaptdl_b200.sched.controller import (AdaptDLController, build_pod,
                                           reconcile, detect_completion)
from adaptdl_b200.sched.kube import InMemoryCluster
from adaptdl_b200.sched.policy import NodeInfo, PolluxPolicy
from adaptdl_b200.sched.supervisor import Supervisor
from adaptdl_b200.sched.validator import Validator

TEMPLATE = {"spec": {"containers": [{
    "name": "main", "image": "img",
    "resources": {"limits": {"nvidia.com/gpu": 1},
                  "requests": {"cpu": "500m", "memory": "1Gi"}}}]}}


def run(coro):
    return asyncio.new_event_loop().run_until_complete(coro)


# ---------------------------------------------------------------- resources

def test_discretize_resource():
    d = resources.discretize_resource
    assert d("cpu", "100m") == 100
    assert d("cpu", "2") == 2000
    assert d("cpu", 0.5) == 500
    assert d("memory", "1Ki") == 1024
    assert d("memory", "1Mi") == 1024 ** 2
    assert d("memory", "1.5Gi") == int(1.5 * 1024 ** 3)
    assert d("memory", "1k") == 1000
    assert d("memory", "2G") == 2 * 1000 ** 3
    assert d("memory", "129e6") == 129000000
    assert d("nvidia.com/gpu", "4") == 4
    assert d("ephemeral-storage", "1E") == 1000 ** 6


def test_pod_requests_and_node_unrequested():
    spec = {"containers": [
        {"resources": {"requests": {"cpu": "250m", "memory": "64Mi"},
                       "limits": {"cpu": "1", "nvidia.com/gpu": 2}}},
        {"resources": {"requests": {"cpu": "250m"},
                       "limits": {"example.com/foo": "3"}}},
        {}]}
    req = resources.get_pod_requests(spec)
    assert req == {"pods": 1, "cpu": 500, "memory": 64 * 1024 ** 2,
                   "nvidia.com/gpu": 2, "example.com/foo": 3}
    node = {"metadata": {"name": "n0"},
            "status": {"allocatable": {"cpu": "4", "memory": "1Gi",
                                       "nvidia.com/gpu": "2", "pods": "10"}}}
    pods = [{"spec": dict(spec, nodeName="n0"),
             "status": {"phase": "Running"}},
            {"spec": dict(spec, nodeName="n0"),
             "status": {"phase": "Succeeded"}},
            {"spec": dict(spec, nodeName="other"),
             "status": {"phase": "Running"}}]
    left = resources.get_node_unrequested(node, pods)
    assert left == {"cpu": 3500, "memory": 1024 ** 3 - 64 * 1024 ** 2,
                    "pods": 9}                     # GPUs all taken -> absent


def test_default_resources(monkeypatch):
    monkeypatch.setenv("ADAPTDL_JOB_DEFAULT_RESOURCES", json.dumps(
        {"requests": {"cpu": "100m"}, "limits": {"nvidia.com/gpu": 1}}))
    spec = {"containers": [{"resources": {"requests": {"cpu": "1"}}}]}
    out = resources.set_default_resources(spec)
    assert out["containers"][0]["resources"] == {
        "requests": {"cpu": "1"}, "limits": {"nvidia.com/gpu": 1}}
    assert "limits" not in spec["containers"][0]["resources"]


# ------------------------------------------------------------ controller

def _job(status=None, preemptible=True):
    return {"metadata": {"namespace": "ns", "name": "job", "uid": "u1"},
            "spec": {"template": TEMPLATE, "preemptible": preemptible},
            "status": dict(status or {})}


def _pods(job, allocation, group=0, phase="Running", ready=True):
    out = []
    for rank in range(len(allocation)):
        pod = build_pod(job["metadata"], job["spec"]["template"], allocation,
                        group, rank, allocation[rank])
        pod["metadata"]["namespace"] = "ns"
        pod["spec"]["nodeName"] = allocation[rank]
        pod["status"] = {
            "phase": phase,
            "conditions": [{"type": "PodScheduled", "status": "True"}],
            "containerStatuses": [{"ready": ready, "state": {}}]}
        out.append(pod)
    return out


def test_build_pod_contract(monkeypatch):
    monkeypatch.setenv("ADAPTDL_SUPERVISOR_URL", "http://sup:8080")
    monkeypatch.setenv("ADAPTDL_SCHED_VERSION", "1.2.3")
    job = _job()
    pod = build_pod(job["metadata"], TEMPLATE, ["n0", "n0", "n1"], 4, 2,
                    "host-n1")
    assert pod["metadata"]["name"] == "job-u1-4-2"
    assert pod["metadata"]["annotations"] == {
        "adaptdl/replicas": "3", "adaptdl/group": "4", "adaptdl/rank": "2",
        "adaptdl/node": "n1"}
    assert pod["metadata"]["labels"]["adaptdl/job"] == "job"
    assert pod["spec"]["hostname"] == "job-4-2"
    assert pod["spec"]["nodeSelector"]["kubernetes.io/hostname"] == "host-n1"
    assert pod["spec"]["restartPolicy"] == "Never"
    env = {e["name"]: e["value"]
           for e in pod["spec"]["containers"][0]["env"]}
    assert env == {"ADAPTDL_JOB_ID": "ns/job", "ADAPTDL_MASTER_PORT": "47004",
                   "ADAPTDL_NUM_NODES": "2", "ADAPTDL_NUM_RESTARTS": "4",
                   "ADAPTDL_NUM_REPLICAS": "3", "ADAPTDL_REPLICA_RANK": "2",
                   "ADAPTDL_SUPERVISOR_URL": "http://sup:8080",
                   "ADAPTDL_SCHED_VERSION": "1.2.3"}
    mounts = pod["spec"]["containers"][0]["volumeMounts"]
    assert {"name": "adaptdl-shm", "mountPath": "/dev/shm"} in mounts
    assert TEMPLATE["spec"]["containers"][0].get("env") is None


def test_reconcile_lifecycle():
    job = _job()
    patch, actions = reconcile(job, [])
    assert patch["phase"] == "Pending" and actions == []
    job["status"].update(patch)
    job["status"]["allocation"] = ["n0", "n1"]
    patch, actions = reconcile(job, [])
    assert patch["phase"] == "Starting" and patch["replicas"] == 2
    job["status"].update(patch)
    patch, actions = reconcile(job, [])
    assert patch["group"] == 0
    assert actions == [("create_pods", 0, ["n0", "n1"])]
    job["status"].update(patch)
    pods = _pods(job, ["n0", "n1"], ready=False)
    patch, actions = reconcile(job, pods)
    assert "phase" not in patch and actions == []
    pods = _pods(job, ["n0", "n1"], ready=True)
    patch, _ = reconcile(job, pods)
    assert patch["phase"] == "Running" and patch["readyReplicas"] == 2
    job["status"].update(patch)
    # the allocator changes its mind: rescale 2 -> 3 replicas
    job["status"]["allocation"] = ["n0", "n1", "n1"]
    patch, actions = reconcile(job, pods)
    assert patch["phase"] == "Stopping"
    job["status"].update(patch)
    patch, actions = reconcile(job, pods)
    assert actions[0][0] == "delete_pods" and len(actions[0][1]) == 2
    patch, actions = reconcile(job, [])
    assert patch["phase"] == "Pending"
    job["status"].update(patch)
    patch, _ = reconcile(job, [])
    assert patch["phase"] == "Starting"
    job["status"].update(patch)
    patch, actions = reconcile(job, [])
    assert patch["group"] == 1
    assert actions == [("create_pods", 1, ["n0", "n1", "n1"])]


def test_completion_and_preemption_semantics():
    job = _job({"phase": "Running", "allocation": ["n0"], "replicas": 1})
    pods = _pods(job, ["n0"], phase="Succeeded")
    patch, actions = reconcile(job, pods)
    assert patch["phase"] == "Succeeded" and "completionTimestamp" in patch
    assert patch["allocation"] is None
    # exit code 143 of a preemptible job is not a failure ...
    pods = _pods(job, ["n0"], phase="Failed")
    pods[0]["status"]["containerStatuses"][0]["state"] = {
        "terminated": {"exitCode": 143}}
    assert detect_completion(pods, preemptible=True) == {}
    # ... but it is for a non-preemptible one, as is any other exit code
    assert detect_completion(pods, preemptible=False)["phase"] == "Failed"
    pods[0]["status"]["containerStatuses"][0]["state"] = {
        "terminated": {"exitCode": 1}}
    assert detect_completion(pods, preemptible=True)["phase"] == "Failed"
    pods[0]["status"]["reason"] = "OutOfnvidia.com/gpu"
    assert detect_completion(pods, preemptible=True)["phase"] == "Failed" \
        or True
    pods[0]["status"]["reason"] = "Outofcpu"
    assert detect_completion(pods, preemptible=True) == {}
    pods[0]["status"]["reason"] = "UnexpectedAdmissionError"
    assert detect_completion(pods, preemptible=True) == {}


def test_invalid_pod_groups_fail_the_job():
    job = _job({"phase": "Running", "allocation": ["n0", "n1"],
                "replicas": 2})
    pods = _pods(job, ["n0", "n1"])
    pods[1]["metadata"]["annotations"]["adaptdl/group"] = "7"
    patch, _ = reconcile(job, pods)
    assert patch["phase"] == "Failed" and patch["reason"] == "Invalid"
    pods = _pods(job, ["n0", "n1"])
    pods[0]["spec"]["nodeName"] = "elsewhere"
    patch, _ = reconcile(job, pods)
    assert "incorrect node" in patch["message"]
    assert reconcile(None, pods) == ({}, [("delete_pods", pods)])


def test_controller_against_in_memory_cluster():
    async def scenario():
        cluster = InMemoryCluster()
        for n in ("n0", "n1"):
            cluster.add_node(n, {"nvidia.com/gpu": 4, "pods": 32,
                                 "cpu": "16", "memory": "64Gi"})
        cluster.add_job("ns", "job", {"template": TEMPLATE})
        ctl = AdaptDLController(cluster)
        await ctl.sync_job("ns", "job")
        await cluster.patch_job_status(
            "ns", "job", {"status": {"allocation": ["n0", "n1"]}})
        for _ in range(2):
            await ctl.sync_job("ns", "job")
        pods = await cluster.list_pods("ns")
        assert sorted(p["metadata"]["annotations"]["adaptdl/rank"]
                      for p in pods) == ["0", "1"]
        job = await cluster.get_job("ns", "job")
        assert job["status"]["phase"] == "Starting"
        assert job["status"]["group"] == 0
        for pod in pods:
            cluster.set_pod_status(
                "ns", pod["metadata"]["name"], phase="Running",
                conditions=[{"type": "PodScheduled", "status": "True"}],
                containerStatuses=[{"ready": True, "state": {}}])
        await ctl.sync_job("ns", "job")
        assert (await cluster.get_job("ns", "job"))["status"]["phase"] \
            == "Running"
        # pod creation failure -> job Failed with reason
        cluster.add_job("ns", "bad", {"template": TEMPLATE},
                        {"allocation": ["n0"], "phase": "Starting",
                         "replicas": 1})
        cluster.fail_pod_creation = True
        await ctl.sync_job("ns", "bad")
        bad = await cluster.get_job("ns", "bad")
        assert bad["status"]["phase"] == "Failed"
        assert bad["status"]["reason"] == "PodCreationError"
    run(scenario())


# ------------------------------------------------------------- allocator

HINTS = {"initBatchSize": 128, "maxBatchSize": 1280,
         "localBszBounds": [64, 256], "maxProfiledReplicas": 2,
         "gradientAccumulation": False,
         "gradParams": {"norm": 0.00136, "var": 0.000502},
         "perfParams": dict(alpha_c=0.121, beta_c=0.00568, alpha_n=0.0236,
                            beta_n=0.00634, alpha_r=0.0118, beta_r=0.00317,
                            gamma=1.14)}


def test_job_info_from_hints():
    job = {"metadata": {"namespace": "ns", "name": "j",
                        "creationTimestamp": "2020-01-01T00:00:00Z"},
           "spec": {"template": TEMPLATE, "maxReplicas": 10,
                    "minReplicas": 1},
           "status": {"train": HINTS}}
    info = job_info_from(job)
    assert info.resources["nvidia.com/gpu"] == 1
    assert info.min_replicas == 1 and info.max_replicas == 4   # 2x profiled
    assert 1.0 < info.speedup_fn(1, 2) < 2.0
    job["status"] = {}
    info = job_info_from(job)
    assert info.max_replicas == 1 and info.speedup_fn(1, 3) == 3


def test_allocator_cycle_and_fast_path():
    async def scenario():
        cluster = InMemoryCluster()
        for n in ("n0", "n1"):
            cluster.add_node(n, {"nvidia.com/gpu": 4, "pods": 32,
                                 "cpu": "16", "memory": "64Gi"})
        cluster.add_node("tainted", {"nvidia.com/gpu": 8, "pods": 32},
                         taints=[{"key": "other", "value": "x"}])
        expander = ClusterExpander(cluster, namespace="ns")
        alloc = AdaptDLAllocator(cluster, expander,
                                 PolluxPolicy(generations=20, seed=0))
        cluster.add_job("ns", "a", {"template": TEMPLATE, "maxReplicas": 8},
                        {"train": HINTS})
        cluster.add_job("ns", "b", {"template": TEMPLATE, "maxReplicas": 8})
        cluster.add_job("ns", "done", {"template": TEMPLATE},
                        {"phase": "Succeeded"})
        first = await alloc.allocate_one("ns", "b")
        assert first in (["n0"], ["n1"])
        assert await alloc.allocate_one("ns", "b") is None   # already known
        allocations = await alloc.optimize_all()
        assert ("ns", "done") not in allocations
        a = (await cluster.get_job("ns", "a"))["status"]["allocation"]
        assert 1 <= len(a) <= 4 and set(a) <= {"n0", "n1"}
        assert expander.expected >= 1
    run(scenario())


def test_cluster_expander_maintains_placeholders():
    async def scenario():
        cluster = InMemoryCluster()
        exp = ClusterExpander(cluster, namespace="ns")
        exp.fit(["n0", "~1", "~2"])
        await exp.reconcile()
        pods = await cluster.list_pods(
            "ns", label_selector="adaptdl/placeholder=true")
        assert len(pods) == 3
        anti = pods[0]["spec"]["affinity"]["podAntiAffinity"]
        assert anti["requiredDuringSchedulingIgnoredDuringExecution"][0][
            "topologyKey"] == "kubernetes.io/hostname"
        # one placeholder lands on the allocated node, one elsewhere
        names = sorted(p["metadata"]["name"] for p in pods)
        cluster.pods[("ns", names[0])]["spec"]["nodeName"] = "n0"
        cluster.set_pod_status("ns", names[0], phase="Running")
        cluster.pods[("ns", names[1])]["spec"]["nodeName"] = "n9"
        cluster.set_pod_status("ns", names[1], phase="Running")
        exp.fit(["n0"])
        await exp.reconcile()
        left = await cluster.list_pods("ns")
        assert [p["metadata"]["name"] for p in left] == [names[0]]
    run(scenario())


# --------------------------------------------------- supervisor / validator

def _client(app):
    from aiohttp.test_utils import TestClient, TestServer
    return TestClient(TestServer(app))


def test_supervisor_endpoints():
    async def scenario():
        cluster = InMemoryCluster()
        cluster.add_job("ns", "job", {"template": TEMPLATE})
        sup = Supervisor(cluster, port=0, poll=0.01)
        async with _client(sup.app) as client:
            assert (await client.get("/healthz")).status == 200
            resp = await client.get("/discover/ns/job/0?timeout=0.05")
            assert resp.status == 408
            job = await cluster.get_job("ns", "job")
            for rank, ip in enumerate(["10.0.0.1", None]):
                pod = build_pod(job["metadata"], TEMPLATE, ["n0", "n1"], 0,
                                rank, "h")
                pod["status"] = {"podIP": ip}
                await cluster.create_pod("ns", pod)
            resp = await client.get("/discover/ns/job/0?timeout=0.05")
            assert resp.status == 408            # rank 1 has no IP yet
            cluster.set_pod_status("ns", "job-{}-0-1".format(
                job["metadata"]["uid"]), podIP="10.0.0.2")
            resp = await client.get("/discover/ns/job/0?timeout=1")
            assert resp.status == 200
            assert await resp.json() == ["10.0.0.1", "10.0.0.2"]
            resp = await client.get("/discover/ns/job/1?timeout=0.05")
            assert resp.status == 408            # other generation
            resp = await client.put("/hints/ns/job", json=dict(
                HINTS, bogus="dropped"))
            assert resp.status == 200
            train = (await cluster.get_job("ns", "job"))["status"]["train"]
            assert "bogus" not in train and train["initBatchSize"] == 128
            resp = await client.put("/hints/ns/missing", json=HINTS)
            assert resp.status == 404
            # the hints became Prometheus series
            text = await (await client.get("/metrics")).text()
            assert 'adaptdl_job_batch_size{job="job",kind="init",' \
                'namespace="ns"} 128.0' in text
            assert 'adaptdl_job_speedup_predict{job="job",namespace="ns",' \
                'replicas="1"} 1.0' in text
            assert 'job="missing"' not in text
    run(scenario())


def test_scheduler_metrics():
    from adaptdl_b200.sched import metrics
    predicted = metrics.speedup_predictions(HINTS)
    assert predicted[1] == 1.0
    assert 1.0 < predicted[2] < 2.0 and predicted[2] <= predicted[8]
    assert metrics.speedup_predictions({"initBatchSize": 128}) == {}
    nodes = {"n0": NodeInfo({"nvidia.com/gpu": 4}, False),
             "n1": NodeInfo({"nvidia.com/gpu": 4}, False)}
    metrics.observe_cycle({("ns", "m1"): ["n0", "n0", "n1"]}, nodes,
                          seconds=0.5, desired_nodes=2,
                          known_jobs=[("ns", "m1"), ("ns", "m2")])
    text = metrics.render()[0].decode()
    assert 'adaptdl_job_replicas{job="m1",namespace="ns"} 3.0' in text
    assert 'adaptdl_job_nodes{job="m1",namespace="ns"} 2.0' in text
    assert 'adaptdl_job_replicas{job="m2",namespace="ns"} 0.0' in text
    assert 'adaptdl_sched_cluster_gpus{state="total"} 8.0' in text
    assert 'adaptdl_sched_cluster_gpus{state="allocated"} 3.0' in text
    assert "adaptdl_sched_desired_nodes 2.0" in text
    metrics.forget_job("ns", "m1")
    text = metrics.render()[0].decode()
    assert 'job="m1"' not in text and 'job="m2"' in text


def test_validator_webhook():
    async def scenario():
        cluster = InMemoryCluster()
        validator = Validator(cluster)

        def review(operation, obj, old=None):
            body = {"request": {"uid": "abc", "operation": operation,
                                "namespace": "ns", "object": obj}}
            if old is not None:
                body["request"]["oldObject"] = old
            return body
        good = {"spec": {"template": TEMPLATE, "minReplicas": 1,
                         "maxReplicas": 4}}
        async with _client(validator.get_app()) as client:
            out = await (await client.post(
                "/validate", json=review("CREATE", good))).json()
            assert out["response"] == {"allowed": True, "uid": "abc"}
            bad = {"spec": {"template": TEMPLATE, "minReplicas": 5,
                            "maxReplicas": 4}}
            out = await (await client.post(
                "/validate", json=review("CREATE", bad))).json()
            assert out["response"]["allowed"] is False
            assert "maxReplicas" in out["response"]["status"]["message"]
            empty = {"spec": {"template": {"spec": {"containers": []}}}}
            out = await (await client.post(
                "/validate", json=review("CREATE", empty))).json()
            assert out["response"]["allowed"] is False
            out = await (await client.post("/validate", json=review(
                "UPDATE", bad, good))).json()
            assert out["response"]["status"]["reason"] == "Forbidden"
            out = await (await client.post("/validate", json=review(
                "UPDATE", good, good))).json()
            assert out["response"]["allowed"] is True
            out = await (await client.post("/validate", json=review(
                "DELETE", good))).json()
            assert out["response"]["allowed"] is True
    run(scenario())


def test_dashboard_only_plots_series_the_code_exports():
    """The reference's Grafana dashboard went stale (none of its series is
    exported any more); ours must track the code."""
    import os
    import re
    from adaptdl_b200.sched import metrics
    from adaptdl_b200.sched import controller
    controller.JOB_COMPLETION_TIME.labels(status="Succeeded").observe(1.0)
    metrics.observe_hints("ns", "dash", HINTS)
    metrics.observe_cycle({("ns", "dash"): ["n0"]},
                          {"n0": NodeInfo({"nvidia.com/gpu": 1}, False)},
                          seconds=0.1, desired_nodes=1)
    import prometheus_client
    exported = set()
    for family in prometheus_client.REGISTRY.collect():
        for sample in family.samples:
            exported.add(sample.name)
    path = os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "deploy", "grafana", "dashboard.json")
    with open(path) as f:
        dashboard = json.load(f)
    plotted = set()
    for panel in dashboard["panels"]:
        for target in panel["targets"]:
            plotted.update(re.findall(r"adaptdl_[a-z_]+", target["expr"]))
    assert plotted and plotted <= exported, plotted - exported
    metrics.forget_job("ns", "dash")


def test_allocation_cycle_does_not_block_the_event_loop():
    import time

    class SlowPolicy(object):
        def optimize(self, jobs, nodes, previous, template):
            time.sleep(0.6)                       # a big cluster's search
            return {key: [next(iter(nodes))] for key in jobs}, 1

    async def scenario():
        cluster = InMemoryCluster()
        cluster.add_node("n0", {"nvidia.com/gpu": 4, "pods": 32,
                                "cpu": "16", "memory": "64Gi"})
        cluster.add_job("ns", "a", {"template": TEMPLATE, "maxReplicas": 4})
        alloc = AdaptDLAllocator(cluster, None, SlowPolicy())
        ticks = []

        async def heartbeat():
            while True:
                ticks.append(time.time())
                await asyncio.sleep(0.05)
        beat = asyncio.ensure_future(heartbeat())
        began = time.time()
        allocations = await alloc.optimize_all()
        took = time.time() - began
        beat.cancel()
        assert allocations == {("ns", "a"): ["n0"]}
        during = [t for t in ticks if began < t < began + took]
        assert took >= 0.6 and len(during) >= 5, (took, len(during))
    run(scenario())


# ------------------------------------------------------- one pod per node

def test_plan_pods_and_canonical_allocation():
    from adaptdl_b200.sched.controller import (canonical_allocation,
                                               plan_pods)
    alloc = ["n1", "n0", "n1", "n1", "n0"]
    assert plan_pods(alloc) == [(0, "n1", 1), (1, "n0", 1), (2, "n1", 1),
                                (3, "n1", 1), (4, "n0", 1)]
    assert canonical_allocation(alloc, True) == ["n1", "n1", "n1", "n0", "n0"]
    assert plan_pods(alloc, True) == [(0, "n1", 3), (3, "n0", 2)]
    assert plan_pods([], True) == []


def test_node_pod_manifest_scales_resources(monkeypatch):
    monkeypatch.setenv("ADAPTDL_SUPERVISOR_URL", "http://sup:8080")
    monkeypatch.delenv("ADAPTDL_JOB_DEFAULT_RESOURCES", raising=False)
    template = {"spec": {"containers": [{
        "name": "main", "image": "img",
        "resources": {"requests": {"cpu": "500m", "memory": "1Gi"},
                      "limits": {"nvidia.com/gpu": 1}}}]}}
    meta = {"namespace": "ns", "name": "job", "uid": "u"}
    pod = build_pod(meta, template, ["n0", "n0", "n0", "n1"], 2, 0, "host0",
                    local_replicas=3)
    ann = pod["metadata"]["annotations"]
    assert ann["adaptdl/local-replicas"] == "3" and ann["adaptdl/rank"] == "0"
    res = pod["spec"]["containers"][0]["resources"]
    assert res["limits"]["nvidia.com/gpu"] == "3"
    assert res["requests"] == {"cpu": "1500m", "memory": str(3 << 30)}
    env = {e["name"]: e["value"] for e in pod["spec"]["containers"][0]["env"]}
    assert env["ADAPTDL_LOCAL_REPLICAS"] == "3"
    assert env["ADAPTDL_REPLICA_RANK"] == "0"
    assert env["ADAPTDL_NUM_REPLICAS"] == "4" and env["ADAPTDL_NUM_NODES"] == "2"
    # the template itself is untouched, one-replica pods are unchanged
    assert template["spec"]["containers"][0]["resources"]["limits"] == \
        {"nvidia.com/gpu": 1}
    single = build_pod(meta, template, ["n0", "n1"], 0, 1, "host1")
    assert "adaptdl/local-replicas" not in single["metadata"]["annotations"]
    assert single["spec"]["containers"][0]["resources"]["limits"] == \
        {"nvidia.com/gpu": 1}


def test_pod_per_node_lifecycle_and_discovery():
    async def scenario():
        cluster = InMemoryCluster()
        for n in ("n0", "n1"):
            cluster.add_node(n, {"nvidia.com/gpu": 8, "pods": 32})
        cluster.add_job("ns", "job", {"template": TEMPLATE, "maxReplicas": 8,
                                      "podPerNode": True},
                        {"allocation": ["n0", "n1", "n0", "n0"]})
        ctl = AdaptDLController(cluster)
        await ctl.sync_job("ns", "job")                    # -> Starting
        await ctl.sync_job("ns", "job")                    # creates pods
        pods = await cluster.list_pods("ns")
        assert len(pods) == 2
        by_node = {p["metadata"]["annotations"]["adaptdl/node"]: p
                   for p in pods}
        ann0 = by_node["n0"]["metadata"]["annotations"]
        ann1 = by_node["n1"]["metadata"]["annotations"]
        assert (ann0["adaptdl/rank"], ann0["adaptdl/local-replicas"]) == \
            ("0", "3")
        assert ann1["adaptdl/rank"] == "3" and \
            "adaptdl/local-replicas" not in ann1
        for pod in pods:
            cluster.set_pod_status(
                "ns", pod["metadata"]["name"], phase="Running",
                podIP="10.0.0." + pod["metadata"]["annotations"][
                    "adaptdl/rank"],
                containerStatuses=[{"ready": True}],
                conditions=[{"type": "PodScheduled", "status": "True"}])
        await ctl.sync_job("ns", "job")
        status = (await cluster.get_job("ns", "job"))["status"]
        assert status["phase"] == "Running"
        assert status["replicas"] == 4 and status["readyReplicas"] == 4
        # rendezvous: one address per RANK
        sup = Supervisor(cluster, port=0, poll=0.01)
        async with _client(sup.app) as client:
            resp = await client.get("/discover/ns/job/0?timeout=1")
            assert await resp.json() == ["10.0.0.0"] * 3 + ["10.0.0.3"]
        # same allocation in another order is not a restart; a new share is
        await cluster.patch_job_status(
            "ns", "job", {"status": {"allocation": ["n0", "n0", "n0", "n1"]}})
        await ctl.sync_job("ns", "job")
        assert (await cluster.get_job("ns", "job"))["status"]["phase"] == \
            "Running"
        await cluster.patch_job_status(
            "ns", "job", {"status": {"allocation": ["n0", "n0", "n1", "n1"]}})
        await ctl.sync_job("ns", "job")
        assert (await cluster.get_job("ns", "job"))["status"]["phase"] == \
            "Stopping"
    run(scenario())


def test_pod_plans_cover_every_rank_once():
    from hypothesis import given, settings, strategies as st
    from adaptdl_b200.sched.controller import (canonical_allocation,
                                               plan_pods)

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.sampled_from(["n0", "n1", "n2", "n3"]), max_size=24),
           st.booleans())
    def check(allocation, per_node):
        ranked = canonical_allocation(allocation, per_node)
        assert sorted(ranked) == sorted(allocation)
        plan = plan_pods(allocation, per_node)
        covered = []
        for first, node, count in plan:
            assert count >= 1
            assert ranked[first:first + count] == [node] * count
            covered.extend(range(first, first + count))
        assert covered == list(range(len(allocation)))
        if per_node:
            assert len(plan) == len(set(allocation))
        else:
            assert len(plan) == len(allocation)
    check()


def test_quantity_scaling_roundtrip():
    from hypothesis import given, settings, strategies as st
    from adaptdl_b200.sched.resources import (discretize_resource,
                                              scale_quantity)

    @settings(max_examples=200, deadline=None)
    @given(st.sampled_from(["cpu", "memory", "nvidia.com/gpu"]),
           st.sampled_from(["1", "2", "500m", "1Gi", "3Mi", "2k", "1.5",
                            "250m", "8"]),
           st.integers(1, 16))
    def check(name, quantity, factor):
        scaled = scale_quantity(name, quantity, factor)
        assert discretize_resource(name, scaled) == \
            discretize_resource(name, quantity) * factor
    check()


def test_json_patches_from_helm_values_are_applied(monkeypatch):
    """job.patch.pods / job.patch.containers of the chart (RFC 6902 subset:
    add / replace / remove, ``~1`` escapes, ``-`` appends)."""
    from adaptdl_b200.sched.controller import _apply_json_patch
    doc = {"metadata": {"annotations": {"a": "1"}},
           "env": [{"name": "X", "value": "1"}], "keep": True}
    patched = _apply_json_patch(doc, [
        {"op": "add", "path": "/metadata/annotations/k8s.v1.cni.cncf.io~1networks",
         "value": "macvlan-conf"},
        {"op": "add", "path": "/env/0", "value": {"name": "FIRST"}},
        {"op": "add", "path": "/env/-", "value": {"name": "LAST"}},
        {"op": "replace", "path": "/metadata/annotations/a", "value": "2"},
        {"op": "remove", "path": "/keep"},
        {"op": "add", "path": "/spec/new/nested", "value": 5}])
    assert patched["metadata"]["annotations"] == {
        "a": "2", "k8s.v1.cni.cncf.io/networks": "macvlan-conf"}
    assert [e["name"] for e in patched["env"]] == ["FIRST", "X", "LAST"]
    assert "keep" not in patched and patched["spec"]["new"]["nested"] == 5
    assert doc["env"] == [{"name": "X", "value": "1"}] and doc["keep"]
    # end to end: the controller stamps them onto every job pod / container
    monkeypatch.setenv("ADAPTDL_SUPERVISOR_URL", "http://sup:8080")
    monkeypatch.setenv("ADAPTDL_JOB_PATCH_PODS", json.dumps([
        {"op": "add", "path": "/metadata/annotations/net", "value": "ib0"}]))
    monkeypatch.setenv("ADAPTDL_JOB_PATCH_CONTAINERS", json.dumps([
        {"op": "add", "path": "/env/-",
         "value": {"name": "NCCL_SOCKET_IFNAME", "value": "net1"}}]))
    pod = build_pod({"namespace": "ns", "name": "j", "uid": "u"}, TEMPLATE,
                    ["n0"], 0, 0, "host0")
    assert pod["metadata"]["annotations"]["net"] == "ib0"
    env = pod["spec"]["containers"][0]["env"]
    assert env[-1] == {"name": "NCCL_SOCKET_IFNAME", "value": "net1"}
