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
