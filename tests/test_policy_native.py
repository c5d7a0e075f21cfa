"""The C++ core of the Pollux search (csrc/host/adl_pollux.cpp) against the
numpy implementation it mirrors: repair rules candidate by candidate,
mutation statistics, and the whole search (feasibility, quality,
reproducibility, independence of the thread count)."""
import collections
from collections import Counter
from datetime import datetime, timedelta

import numpy as np
import pytest

from adaptdl_b200.goodput import GoodputFunction, GradParams, PerfParams
from adaptdl_b200.sched.policy import (JobInfo, NodeInfo, PolluxPolicy,
                                       SpeedupFunction)
from adaptdl_b200.sched.policy import native, nsga2
from adaptdl_b200.sched.policy.pollux import ClusterProblem

pytestmark = pytest.mark.skipif(not native.available(),
                                reason="no C++ compiler on this machine")

GPU = "nvidia.com/gpu"


def _random_problem(rng, J=9, N=6, max_replicas=1000):
    jobs = [JobInfo({"gpu": int(rng.integers(1, 3)), "pods": 1,
                     "mem": int(rng.integers(0, 3)) * (1 << 33)},
                    lambda n, r: r, j, int(rng.integers(0, 3)), max_replicas,
                    preemptible=bool(j % 4))
            for j in range(J)]
    nodes = [NodeInfo({"gpu": int(rng.integers(2, 9)), "pods": 5,
                       "mem": 6 << 33}, False) for _ in range(N)]
    base = np.zeros((J, N), dtype=np.int64)
    base[0, 0] = 1                     # job 0 is not preemptible: pinned
    return ClusterProblem(jobs, nodes, base, rng=rng)


def test_native_repair_equals_the_numpy_rules():
    """Rules 1, 2, 4 and 5 are deterministic: the C++ repair must return
    exactly what ClusterProblem.repair returns (rule 3 never fires here:
    max_replicas is out of reach)."""
    rng = np.random.default_rng(11)
    for trial in range(6):
        problem = _random_problem(rng, J=int(rng.integers(3, 12)),
                                  N=int(rng.integers(2, 9)))
        J, N = problem.base.shape
        population = rng.integers(0, 5, size=(50, J, N)) * \
            (rng.random((50, J, N)) < 0.5)
        want = problem.repair(population)
        search = native.NativeSearch(problem, 10, 1, seed=trial)
        for i, candidate in enumerate(population):
            np.testing.assert_array_equal(search.repair(candidate, i),
                                          want[i])
        search.close()


def test_native_repair_trims_rows_over_max_replicas():
    rng = np.random.default_rng(3)
    jobs = [JobInfo({"gpu": 1}, lambda n, r: r, j, 0, 3 + j) for j in range(4)]
    nodes = [NodeInfo({"gpu": 64}, False) for _ in range(5)]
    problem = ClusterProblem(jobs, nodes, np.zeros((4, 5), dtype=np.int64),
                             rng=rng)
    search = native.NativeSearch(problem, 10, 1, seed=5)
    caps = np.array([3, 4, 5, 6])
    trimmed_somewhere = set()
    for i in range(200):
        candidate = rng.integers(0, 4, size=(4, 5))
        fixed = search.repair(candidate, i)
        assert (fixed <= candidate).all() and (fixed >= 0).all()
        assert (fixed.sum(axis=1) <= caps).all()
        # nothing is given away: a trimmed row sits exactly at its cap unless
        # the one-multi-node-job-per-node rule emptied entries first
        after_rule2 = problem.repair(candidate[None])[0]
        for j in range(4):
            if (after_rule2[j] > 0).sum() == (candidate[j] > 0).sum() and \
                    candidate[j].sum() > caps[j]:
                assert fixed[j].sum() == caps[j]
                trimmed_somewhere.update(
                    np.flatnonzero(fixed[j] < candidate[j]).tolist())
    assert trimmed_somewhere == set(range(5))     # random node order
    search.close()


def test_native_mutation_statistics():
    """One expected re-draw among the non-zero entries of a row and one
    among its zero entries left of the growth limit; values inside
    [min_fill, max_fit]; the minimum spread of guaranteed replicas is
    restored; nothing beyond size + growth is touched."""
    rng = np.random.default_rng(5)
    J, N = 6, 16
    jobs = [JobInfo({"gpu": 1}, lambda n, r: r, j, 2 if j == 1 else 0, 64)
            for j in range(J)]
    nodes = [NodeInfo({"gpu": 8}, False) for _ in range(N)]
    base = np.zeros((J, N), dtype=np.int64)
    problem = ClusterProblem(jobs, nodes, base, rng=rng)
    assert problem.min_fill[1, 0] == 2
    search = native.NativeSearch(problem, 10, 1, seed=9)
    state = np.zeros((J, N), dtype=np.int32)
    state[:, :8] = rng.integers(0, 4, size=(J, 8)) * \
        (rng.random((J, 8)) < 0.6)
    state[1, 0] = 0                      # below the job's guaranteed spread
    state[3] = 0                         # an empty row
    size = int(np.flatnonzero(state.any(axis=0)).max()) + 1
    trials = 4000
    changed_pos = np.zeros(J)
    hit_zero = np.zeros(J)
    furthest = 0
    for i in range(trials):
        out = search.mutate(state, i)
        assert (out >= problem.min_fill).all()
        assert (out <= problem.max_fit).all()
        assert out[1, 0] >= 2
        cols = np.flatnonzero((out != state).any(axis=0))
        if len(cols):
            furthest = max(furthest, int(cols.max()))
        positive = state > 0
        changed_pos += ((out != state) & positive).sum(axis=1)
        hit_zero += ((out != state) & ~positive).sum(axis=1)
    assert size <= furthest < N
    for j in range(J):
        nz = int((state[j] > 0).sum())
        if nz and j != 1:
            # each non-zero entry is re-drawn with probability 1/nz and the
            # new value differs from the old one with probability 8/9
            assert changed_pos[j] / trials == pytest.approx(8 / 9, rel=0.15)
        # zero entries: (size + E[growth] - nz) / (N - nz) hits, of which 8/9
        # draw a non-zero value; the growth is 1 + geometric, capped by N
        if j != 1:
            low = (size + 1 - nz) / (N - nz) * 8 / 9
            high = (size + 3 - nz) / (N - nz) * 8 / 9
            assert 0.85 * low <= hit_zero[j] / trials <= 1.15 * high
    search.close()


def _cluster(num_jobs, num_nodes, seed, gpus_per_node=4):
    rng = np.random.RandomState(seed)
    jobs = {}
    for i in range(num_jobs):
        perf = PerfParams(0.1 * rng.uniform(.5, 2), 0.01 * rng.uniform(.5, 2),
                          0.05, 0.002, 0.02, 0.001, 1.2)
        grad = GradParams(sqr=rng.uniform(0.001, 0.1),
                          var=rng.uniform(0.01, 1.0))
        speedup = SpeedupFunction(
            GoodputFunction(perf, grad, 128), max_batch_size=4096,
            atomic_bsz_range=(32, 512), accumulation=True, mem_size=32)
        jobs["job-%d" % i] = JobInfo({GPU: 1, "pods": 1}, speedup, i,
                                     int(rng.randint(0, 2)),
                                     min(16, 2 ** rng.randint(1, 5)))
    resources = {GPU: gpus_per_node, "pods": 32}
    nodes = {"node-%d" % i: NodeInfo(dict(resources), False)
             for i in range(num_nodes)}
    return jobs, nodes, NodeInfo(dict(resources), True)


def _check_feasible(allocations, jobs, nodes):
    per_node = Counter()
    for key, placement in allocations.items():
        assert len(placement) <= jobs[key].max_replicas
        assert len(placement) == 0 or len(placement) >= jobs[key].min_replicas
        per_node.update(placement)
    for node, count in per_node.items():
        assert count <= nodes[node].resources[GPU]
    spread = {k: set(v) for k, v in allocations.items() if len(set(v)) > 1}
    for node in nodes:
        assert sum(node in s for s in spread.values()) <= 1


def _value(allocations, jobs):
    return sum(float(jobs[k].speedup_fn(len(set(a)), len(a)))
               for k, a in allocations.items() if a)


def test_native_search_matches_numpy_search_quality():
    jobs, nodes, template = _cluster(24, 8, seed=1)
    totals = collections.defaultdict(list)
    for seed in range(3):
        for flavour in (True, False):
            policy = PolluxPolicy(generations=40, seed=seed, native=flavour)
            assert policy._native is flavour
            previous = {}
            for _ in range(3):
                allocations, desired = policy.optimize(jobs, nodes, previous,
                                                       template)
                _check_feasible(allocations, jobs, nodes)
                assert desired >= 1
                previous = allocations
            totals[flavour].append(_value(allocations, jobs))
    # the same search with other random numbers: same quality of the
    # allocation it settles on (sum of speedups), within a few per cent
    assert np.mean(totals[True]) >= 0.95 * np.mean(totals[False])


def test_native_search_is_reproducible_and_thread_independent(monkeypatch):
    jobs, nodes, template = _cluster(300, 12, seed=2)   # above the threshold
    results = []
    for threads in ("1", "1", "4"):
        monkeypatch.setenv("ADAPTDL_B200_POLICY_THREADS", threads)
        policy = PolluxPolicy(generations=15, seed=123, native=True)
        allocations, desired = policy.optimize(jobs, nodes, {}, template)
        results.append((allocations, desired, policy._prev_states.copy()))
    for other in results[1:]:
        assert other[0] == results[0][0] and other[1] == results[0][1]
        np.testing.assert_array_equal(other[2], results[0][2])


def test_native_population_is_feasible_and_scored_like_numpy():
    """Every member of the final population passes the numpy repair
    unchanged and carries the objectives ClusterProblem.evaluate gives."""
    rng = np.random.default_rng(0)
    jobs, nodes, template = _cluster(12, 4, seed=3)
    job_list = list(jobs.values())
    node_list = list(nodes.values()) + [template] * len(nodes)
    base = np.zeros((len(job_list), len(node_list)), dtype=np.int64)
    base[2, 1] = 2
    base[5, 0] = 1
    problem = ClusterProblem(job_list, node_list, base, rng=rng)
    states, values = native.minimize(problem, base[None], 40, 30, rng)
    assert 1 < len(states) <= 40
    np.testing.assert_array_equal(problem.repair(states), states)
    np.testing.assert_allclose(problem.evaluate(states), values, rtol=1e-12)
    assert len({s.tobytes() for s in states}) == len(states)   # no twins
    front = nsga2.non_dominated_fronts(values)[0]
    assert len(front) >= 2             # several cluster sizes on the front


def test_native_search_with_plain_callables_and_pinned_jobs():
    now = datetime.now()
    jobs = {i: JobInfo({GPU: 1, "pods": 1}, lambda n, r: r ** 0.5,
                       now + timedelta(minutes=i), 0, 8,
                       preemptible=(i != 0))
            for i in range(6)}
    nodes = {i: NodeInfo({GPU: 4, "pods": 32}, False) for i in range(3)}
    template = NodeInfo({GPU: 4, "pods": 32}, True)
    policy = PolluxPolicy(generations=20, seed=0, native=True)
    previous = {0: [1, 1]}
    for _ in range(3):
        allocations, _ = policy.optimize(jobs, nodes, previous, template)
        assert allocations[0] == [1, 1]      # non-preemptible: stays put
        _check_feasible(allocations, jobs, nodes)
        previous = allocations
    assert sum(len(v) for v in allocations.values()) == 12   # all GPUs busy


def test_numpy_path_can_be_forced(monkeypatch):
    monkeypatch.setenv("ADAPTDL_B200_NATIVE_POLICY", "0")
    assert PolluxPolicy(seed=0)._native is False
    monkeypatch.delenv("ADAPTDL_B200_NATIVE_POLICY")
    assert PolluxPolicy(seed=0)._native is True
