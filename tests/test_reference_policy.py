"""The reference's UNMODIFIED scheduling policy, run on the pymoo stand-in of
baseline/shims (baseline/ref_policy.py), next to this framework's policy: the
baseline arm of tools/sched_sim.py --reference-policy and
tools/policy_bench.py --search reference."""
import os
import sys
from collections import Counter
from datetime import datetime, timedelta

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_policy  # noqa: E402

pytestmark = pytest.mark.skipif(
    not ref_policy.available(),
    reason="reference policy not installed (baseline/install_reference.sh)")

GPU = "nvidia.com/gpu"


@pytest.fixture()
def reference(monkeypatch):
    before = list(sys.path)
    modules = {n: m for n, m in sys.modules.items()
               if n.split(".")[0] in ("adaptdl_sched", "pymoo")}
    monkeypatch.setattr(np, "int", int, raising=False)
    monkeypatch.setattr(np, "float", float, raising=False)
    loaded = ref_policy.load()
    yield loaded
    # leave no trace: the alias package of the same name must keep working
    sys.path[:] = before
    for name in [n for n in sys.modules
                 if n.split(".")[0] in ("adaptdl_sched", "pymoo")]:
        del sys.modules[name]
    sys.modules.update(modules)


def _jobs(job_class, speedup_fn, count, max_replicas=8):
    now = datetime.now()
    return {i: job_class({GPU: 1, "pods": 1}, speedup_fn,
                         now + timedelta(minutes=i), 0, max_replicas)
            for i in range(count)}


def _check(allocations, jobs, nodes):
    per_node = Counter()
    for key, placement in allocations.items():
        assert len(placement) <= jobs[key].max_replicas
        per_node.update(placement)
    for node, count in per_node.items():
        assert count <= nodes[node].resources[GPU]
    spread = {k: set(v) for k, v in allocations.items() if len(set(v)) > 1}
    for node in nodes:
        assert sum(node in s for s in spread.values()) <= 1


def test_reference_policy_runs_unmodified_and_respects_the_rules(reference):
    policy_class, job_class, node_class, speedup_class = reference
    assert "_ref_sched" in sys.modules[policy_class.__module__].__file__
    from adaptdl_b200.goodput import GoodputFunction, GradParams, PerfParams
    goodput = GoodputFunction(
        PerfParams(0.121, 0.00568, 0.0236, 0.00634, 0.0118, 0.00317, 1.14),
        GradParams(sqr=0.00136, var=0.000502), 128)
    speedup = speedup_class(goodput, max_batch_size=1280,
                            atomic_bsz_range=(64, 256))
    jobs = _jobs(job_class, speedup, 4)
    nodes = {i: node_class({GPU: 4, "pods": 32}, preemptible=False)
             for i in range(2)}
    template = node_class({GPU: 4, "pods": 32}, preemptible=True)
    np.random.seed(0)
    policy = policy_class()
    previous = {}
    for _ in range(2):
        allocations, desired = policy.optimize(jobs, nodes, previous,
                                               template)
        _check(allocations, jobs, nodes)
        assert desired >= 1
        previous = allocations
    assert sum(len(v) for v in allocations.values()) > 0


def test_this_policy_is_at_least_as_good_on_the_same_cluster(reference):
    """Same jobs and nodes for both policies, three warm-started cycles: the
    sum of speedups of the allocation each settles on."""
    policy_class = reference[0]
    from adaptdl_b200.sched.policy import JobInfo, NodeInfo, PolluxPolicy
    speedup = lambda n, r: np.asarray(r, dtype=float) ** 0.8   # noqa: E731
    jobs = _jobs(JobInfo, speedup, 10)
    nodes = {i: NodeInfo({GPU: 4, "pods": 32}, False) for i in range(4)}
    template = NodeInfo({GPU: 4, "pods": 32}, True)

    def settle(policy):
        previous = {}
        for _ in range(3):
            previous, _ = policy.optimize(jobs, nodes, previous, template)
            _check(previous, jobs, nodes)
        return (sum(len(a) ** 0.8 for a in previous.values() if a),
                sum(len(a) for a in previous.values()))
    np.random.seed(1)
    theirs, _ = settle(policy_class())
    ours, gpus_in_use = settle(PolluxPolicy(seed=1))
    assert ours >= theirs - 1e-9
    assert gpus_in_use == 16                  # nothing idles next to a job
