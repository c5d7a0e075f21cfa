"""Scheduling policy: NSGA-II, speedup memoisation, Pollux allocation
invariants (ideas from the reference's policy/*_test.py)."""
import collections
import copy
import random
from collections import Counter
from datetime import datetime, timedelta
from unittest.mock import Mock

import numpy as np
import pytest

from adaptdl_b200.goodput import GoodputFunction, GradParams, PerfParams
from adaptdl_b200.sched.policy import (JobInfo, NodeInfo, PolluxPolicy,
                                       SpeedupFunction)
from adaptdl_b200.sched.policy import nsga2

PERF = PerfParams(0.121, 0.00568, 0.0236, 0.00634, 0.0118, 0.00317, 1.14)
GRAD = GradParams(sqr=0.00136, var=0.000502)


def _speedup_fn():
    return SpeedupFunction(GoodputFunction(PERF, GRAD, 128),
                           max_batch_size=1280, atomic_bsz_range=(64, 256))


def test_non_dominated_fronts_and_crowding():
    F = np.array([[1, 5], [2, 2], [5, 1], [3, 3], [6, 6], [2, 2]], float)
    fronts = nsga2.non_dominated_fronts(F)
    assert sorted(fronts[0].tolist()) == [0, 1, 2, 5]
    assert sorted(fronts[1].tolist()) == [3]
    assert sorted(fronts[2].tolist()) == [4]
    d = nsga2.crowding_distance(F[[0, 1, 2]])
    assert np.isinf(d[0]) and np.isinf(d[2]) and np.isfinite(d[1])


def test_nsga2_finds_pareto_front_of_toy_problem():
    # genome: one integer x in [0, 20]; objectives (x^2, (x-10)^2):
    # the Pareto set is x in [0, 10]
    class Toy:
        def evaluate(self, X):
            x = X[:, 0].astype(float)
            return np.column_stack([x ** 2, (x - 10) ** 2])

        def crossover(self, a, b, rng):
            mix = (a + b) // 2
            return np.concatenate([mix, np.maximum(a, b)])

        def mutate(self, X, rng):
            return X + rng.integers(-2, 3, size=X.shape)

        def repair(self, X):
            return np.clip(X, 0, 20)
    X, F = nsga2.minimize(Toy(), np.array([[20]]), pop_size=24, n_gen=40,
                          rng=np.random.default_rng(0))
    front = nsga2.non_dominated_fronts(F)[0]
    xs = set(X[front, 0].tolist())
    assert xs <= set(range(0, 11)) and len(xs) >= 8


def test_speedup_memoization():
    goodput_fn = Mock()
    calls = []

    def optimize(num_nodes, num_replicas, **kw):
        calls.append(np.size(num_replicas))
        r = np.asarray(num_replicas, dtype=float)
        return r, r.astype(int), np.zeros_like(r, dtype=int)
    goodput_fn.optimize.side_effect = optimize
    fn = SpeedupFunction(goodput_fn, mem_size=8)
    assert calls == [1]                       # the (1, 1) baseline
    assert fn(1, 3) == pytest.approx(3.0)
    assert fn(1, 3) == pytest.approx(3.0)     # memoised: no new call
    assert len(calls) == 2
    out = fn(np.array([1, 1, 2, 0]), np.array([3, 4, 4, 0]))
    assert out.tolist() == pytest.approx([3.0, 4.0, 4.0, 0.0])
    assert calls[-1] == 2                     # only the 2 unseen pairs
    assert fn(2, 16) == pytest.approx(16.0)   # beyond the table: computed
    assert fn(2, 16) == pytest.approx(16.0)
    assert len(calls) == 5


def test_speedup_realistic_is_sublinear_and_monotone():
    fn = _speedup_fn()
    replicas = np.arange(1, 17)
    single = fn(np.ones_like(replicas), replicas)
    assert single[0] == pytest.approx(1.0)
    assert np.all(np.diff(single[:8]) >= -1e-9)   # retrogression later
    assert np.all(single <= replicas + 1e-9)
    multi = fn(np.minimum(replicas, 4), replicas)
    assert np.all(multi[1:] <= single[1:] + 1e-9)   # crossing nodes costs


@pytest.mark.parametrize("num_nodes", [1, 2, 4, 8, 16])
def test_optimize_respects_capacities(num_nodes, total_devices=16):
    num_devices = total_devices // num_nodes
    speedup_fn = _speedup_fn()
    now = datetime.now()
    jobs = {}
    for i in range(16):
        jobs[i] = JobInfo({"nvidia.com/gpu": 1, "pods": 1}, speedup_fn,
                          now + timedelta(minutes=i), 0, 8)
    node_resources = {"nvidia.com/gpu": num_devices, "pods": 32}
    nodes = {i: NodeInfo(node_resources, preemptible=False)
             for i in range(num_nodes)}
    template = NodeInfo(node_resources, preemptible=True)
    policy = PolluxPolicy(generations=30, seed=0)
    prev = {}
    for cycle in range(3):
        allocations, desired = policy.optimize(jobs, nodes, prev, template)
        assert desired >= 1
        per_node = Counter()
        for key, placement in allocations.items():
            assert len(placement) <= jobs[key].max_replicas
            per_node.update(placement)
            # at most one multi-node job per node is checked below
        for node, count in per_node.items():
            assert count <= nodes[node].resources["nvidia.com/gpu"]
        spread = {k: set(v) for k, v in allocations.items()
                  if len(set(v)) > 1}
        for node in nodes:
            assert sum(node in s for s in spread.values()) <= 1
        assert sum(len(v) for v in allocations.values()) > 0
        prev = allocations


def test_allocate_job():
    nodes = {
        "0": NodeInfo({"gpu": 1, "cpu": 500, "pods": 32}, preemptible=False),
        "1": NodeInfo({"gpu": 2, "cpu": 2000, "pods": 32}, preemptible=False),
        "2": NodeInfo({"gpu": 2, "cpu": 3000, "pods": 32}, preemptible=True),
    }
    fn = _speedup_fn()
    now = datetime.now()
    job_1 = JobInfo({"gpu": 1, "cpu": 500, "pods": 1}, fn, now, 0, 1)
    job_2 = JobInfo({"gpu": 1, "cpu": 1000, "pods": 1}, fn, now, 0, 1)
    job_3 = JobInfo({"gpu": 1, "cpu": 1000, "pods": 1}, fn, now, 2, 2)
    job_4 = JobInfo({"gpu": 1, "cpu": 2000, "pods": 1}, fn, now, 2, 2)
    policy = PolluxPolicy()
    assert policy.allocate_job(job_1, nodes) == ["0"]
    assert policy.allocate_job(job_2, nodes) == ["1"]
    assert policy.allocate_job(job_3, nodes) == ["1", "1"]
    assert policy.allocate_job(job_4, nodes) == []


def test_unusable_node_asks_for_more_nodes():
    nodes = {
        0: NodeInfo({"gpu": 1, "cpu": 500, "pods": 32}, preemptible=False),
        1: NodeInfo({"gpu": 1, "cpu": 8000, "pods": 32}, preemptible=False),
        2: NodeInfo({"gpu": 1, "cpu": 8000, "pods": 32}, preemptible=False),
    }
    template = NodeInfo({"gpu": 1, "cpu": 8000, "pods": 32}, preemptible=True)
    fn = _speedup_fn()
    now = datetime.now()
    jobs = {i: JobInfo({"gpu": 1, "cpu": 1000, "pods": 1}, fn,
                       now + timedelta(minutes=i), 0, 1) for i in range(3)}
    policy = PolluxPolicy(seed=1)
    allocations, desired = policy.optimize(jobs, nodes, {}, template)
    assert desired > 3
    assert max(len(a) for a in allocations.values()) == 1
    assert sum(len(a) for a in allocations.values()) == 2


@pytest.mark.parametrize("num_nodes", [1, 2, 4, 8])
def test_non_preemptible_jobs_keep_their_allocation(num_nodes,
                                                    total_devices=16):
    random.seed(num_nodes)
    ids = list(range(10))
    random.shuffle(ids)
    preemptible, fixed = ids[:5], ids[5:]
    num_devices = total_devices // num_nodes
    fn = _speedup_fn()
    now = datetime.now()
    policy = PolluxPolicy(generations=30, seed=num_nodes)
    res = {"nvidia.com/gpu": 1, "pods": 1}
    node_res = {"nvidia.com/gpu": num_devices, "pods": 32}
    nodes = {i: NodeInfo(node_res, preemptible=False)
             for i in range(num_nodes)}
    template = NodeInfo(node_res, preemptible=True)
    prev = {i: [] for i in ids}
    for cycle in range(3):
        jobs = {}
        for i in preemptible:
            jobs[i] = JobInfo(res, fn, now + timedelta(minutes=i), 0, 8)
        for i in fixed:
            jobs[i] = JobInfo(res, fn, now + timedelta(minutes=i), 2, 4,
                              preemptible=False)
        allocations, _ = policy.optimize(jobs, nodes, prev, template)
        per_node = Counter()
        for key, placement in allocations.items():
            assert len(placement) <= jobs[key].max_replicas
            if placement:
                assert len(placement) >= jobs[key].min_replicas
            per_node.update(placement)
        for node, count in per_node.items():
            assert count <= nodes[node].resources["nvidia.com/gpu"]
        for i in fixed:
            if i in allocations and prev.get(i):
                assert sorted(allocations[i]) == sorted(prev[i])
        prev = copy.deepcopy(allocations)
        victim = random.choice(sorted(allocations))
        (fixed if victim in fixed else preemptible).remove(victim)
        prev.pop(victim)


def test_starved_jobs_get_their_minimum_next_to_idle_capacity():
    from adaptdl_b200.sched.policy import JobInfo, NodeInfo, PolluxPolicy
    gpu = "nvidia.com/gpu"
    nodes = {"a": NodeInfo({gpu: 4, "pods": 8}, False),
             "b": NodeInfo({gpu: 4, "pods": 8}, False)}
    mk = lambda t, lo=0: JobInfo({gpu: 1, "pods": 1}, lambda n, r: r, t, lo, 8)
    jobs = {"old": mk(0), "new": mk(1), "pair": mk(2, lo=2), "big": mk(3, lo=4)}
    alloc = {"old": ["a", "a", "a"], "new": [], "pair": [], "big": []}
    PolluxPolicy._place_starved(alloc, jobs, nodes)
    assert alloc["old"] == ["a", "a", "a"]          # untouched
    assert alloc["new"] == ["a"]                      # packs onto the used node
    assert alloc["pair"] == ["b", "b"]                # minimum, on one node
    assert alloc["big"] == []                         # 4 do not fit anywhere
    used = {}
    for placement in alloc.values():
        for node in placement:
            used[node] = used.get(node, 0) + 1
    assert all(used[n] <= nodes[n].resources[gpu] for n in used)


def _repair_oracle(problem, candidate):
    """Straightforward per-candidate version of ClusterProblem.repair
    (without the max_replicas rule, whose trimming order is random)."""
    state = np.array(candidate, dtype=np.int64)
    J, N = state.shape
    for j in problem.pinned:
        state[j] = problem.base[j]
    multi = [np.count_nonzero(state[j]) > 1 for j in range(J)]
    for n in range(N):
        taken = False
        for j in range(J):
            if multi[j] and state[j, n] > 0:
                if taken:
                    state[j, n] = 0
                taken = True
    for r in range(problem.job_res.shape[1]):
        for n in range(N):
            need = problem.job_res[:, r]
            if (state[:, n] * need).sum() <= problem.node_res[n, r]:
                continue
            left = problem.node_res[n, r]
            for j in range(J):
                if need[j] == 0:
                    continue
                grant = min(state[j, n] * need[j], left)
                left -= grant
                state[j, n] = min(state[j, n], grant // need[j])
    for j in range(J):
        if state[j].sum() < problem.min_replicas[j]:
            state[j] = 0
    return state


def test_vectorised_repair_matches_a_plain_loop():
    from adaptdl_b200.sched.policy.pollux import ClusterProblem
    rng = np.random.default_rng(7)
    J, N = 9, 6
    jobs = [JobInfo({"gpu": int(rng.integers(1, 3)), "pods": 1,
                     "mem": int(rng.integers(0, 3)) * (1 << 33)},
                    lambda n, r: r, j, int(rng.integers(0, 3)), 1000,
                    preemptible=bool(j % 4))
            for j in range(J)]
    nodes = [NodeInfo({"gpu": int(rng.integers(2, 9)), "pods": 5,
                       "mem": 6 << 33}, False) for _ in range(N)]
    base = np.zeros((J, N), dtype=np.int64)
    base[0, 0] = 1                     # job 0 is not preemptible: pinned
    problem = ClusterProblem(jobs, nodes, base, rng=rng)
    assert list(problem.pinned) == [0]
    population = rng.integers(0, 5, size=(40, J, N)) * \
        (rng.random((40, J, N)) < 0.5)
    fixed = problem.repair(population)
    assert fixed.shape == population.shape
    for got, candidate in zip(fixed, population):
        np.testing.assert_array_equal(got, _repair_oracle(problem, candidate))
    # idempotent, and every resource within capacity
    np.testing.assert_array_equal(problem.repair(fixed), fixed)
    use = np.einsum("pjn,jr->pnr", fixed, problem.job_res)
    assert (use <= problem.node_res[None]).all()


def test_repair_trims_rows_over_max_replicas():
    from adaptdl_b200.sched.policy.pollux import ClusterProblem
    rng = np.random.default_rng(3)
    jobs = [JobInfo({"gpu": 1}, lambda n, r: r, j, 0, 3 + j) for j in range(4)]
    nodes = [NodeInfo({"gpu": 64}, False) for _ in range(5)]
    problem = ClusterProblem(jobs, nodes, np.zeros((4, 5), dtype=np.int64),
                             rng=rng)
    population = rng.integers(0, 4, size=(30, 4, 5))
    fixed = problem.repair(population)
    assert (fixed <= population).all() and (fixed >= 0).all()
    caps = np.array([3, 4, 5, 6])
    multi_node_rule = population.copy()          # rule 2 may empty entries
    assert (fixed.sum(axis=2) <= caps).all()
    # rows already within their cap and alone on their nodes are untouched
    ok = population.sum(axis=2) <= caps
    single = np.count_nonzero(population, axis=2) <= 1
    keep = ok & single
    np.testing.assert_array_equal(fixed[keep], multi_node_rule[keep])


def test_speedup_lookup_equals_call():
    fn = _speedup_fn()
    nodes = np.array([1, 1, 2, 2, 4])
    replicas = np.array([1, 3, 2, 8, 70])       # 70 is beyond the table
    np.testing.assert_allclose(fn.lookup(nodes, replicas),
                               fn(nodes, replicas))


def test_row_keys_distinguish_rows():
    rng = np.random.default_rng(0)
    X = rng.integers(0, 9, size=(300, 500)).astype(np.int32)
    X[17] = X[3]
    keep = nsga2._unique_rows(X)
    assert keep.sum() == 299 and not keep[17]
    assert not nsga2._unique_rows(X[:5], against=X).any()


def test_non_dominated_sorting_properties():
    from hypothesis import given, settings, strategies as st

    def dominates(a, b):
        return all(x <= y for x, y in zip(a, b)) and \
            any(x < y for x, y in zip(a, b))

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.tuples(st.integers(0, 6), st.integers(0, 6)),
                    min_size=1, max_size=25))
    def check(points):
        F = np.array(points, dtype=float)
        fronts = nsga2.non_dominated_fronts(F)
        flat = sorted(int(i) for front in fronts for i in front)
        assert flat == list(range(len(points)))          # a partition
        rank = {int(i): r for r, front in enumerate(fronts) for i in front}
        for i, a in enumerate(points):
            for j, b in enumerate(points):
                if dominates(a, b):
                    assert rank[i] < rank[j]
            # every point outside front 0 is dominated by someone one front up
            if rank[i] > 0:
                assert any(dominates(points[j], a) and rank[j] == rank[i] - 1
                           for j in range(len(points)))
        for front in fronts:
            dist = nsga2.crowding_distance(F[front])
            assert (dist >= 0).all()
            if len(front) > 2:
                for k in range(F.shape[1]):
                    col = F[front, k]
                    assert np.isinf(dist[np.argmin(col)]) or \
                        np.isinf(dist).sum() >= 2
    check()


def test_idle_capacity_goes_to_the_jobs_that_gain_most():
    """The greedy pass after the search: one replica at a time to the best
    marginal gain, inside the usable nodes, never breaking the rules."""
    gpu = "nvidia.com/gpu"
    nodes = collections.OrderedDict(
        (name, NodeInfo({gpu: 4, "pods": 8}, False)) for name in "abcd")

    def mk(t, fn, cap=8, lo=0, preemptible=True):
        return JobInfo({gpu: 1, "pods": 1}, fn, t, lo, cap,
                       preemptible=preemptible)
    # crossing nodes costs: replicas / nodes ** 0.25
    scal = lambda n, r: r / max(n, 1) ** 0.25 if r else 0.0      # noqa: E731
    flat = lambda n, r: min(r, 2)                                # noqa: E731
    jobs = collections.OrderedDict([
        ("scales", mk(0, scal)),
        ("flat", mk(1, flat)),                 # nothing to gain beyond 2
        ("capped", mk(2, scal, cap=2)),
        ("pinned", mk(3, scal, preemptible=False)),
        ("waiting", mk(4, scal, cap=16, lo=9)),   # not running: left alone
    ])
    base = {"scales": ["a"], "flat": ["a", "a"], "capped": ["b", "b"],
            "pinned": ["b"], "waiting": []}
    alloc = {k: list(v) for k, v in base.items()}
    PolluxPolicy._grow_into_idle(alloc, jobs, nodes, base, ["a", "b", "c"])
    assert alloc["flat"] == ["a", "a"]
    assert alloc["capped"] == ["b", "b"]
    assert alloc["pinned"] == ["b"]
    assert alloc["waiting"] == []
    # "scales" takes the free GPU of its own node first, then the free GPU
    # of "b" (nodes in use before empty ones), then all of "c"; "d" is not
    # usable
    assert Counter(alloc["scales"]) == Counter({"a": 2, "b": 1, "c": 4})
    per_node = Counter()
    for placement in alloc.values():
        per_node.update(placement)
    assert all(per_node[n] <= 4 for n in per_node)
    # a second multi-node job may not share a node with the first one
    jobs2 = collections.OrderedDict([("one", mk(0, scal)), ("two", mk(1, scal))])
    base2 = {"one": ["a", "a", "b", "b"], "two": ["c", "c", "c", "c"]}
    alloc2 = {k: list(v) for k, v in base2.items()}
    PolluxPolicy._grow_into_idle(alloc2, jobs2, nodes, {}, list("abcd"))
    spread = {k: set(v) for k, v in alloc2.items() if len(set(v)) > 1}
    for node in nodes:
        assert sum(node in s for s in spread.values()) <= 1
    assert len(alloc2["one"]) == 8 and set(alloc2["one"]) == {"a", "b"}
    assert Counter(alloc2["two"]) == Counter({"c": 4, "d": 4})


def test_growth_respects_the_restart_penalty():
    gpu = "nvidia.com/gpu"
    nodes = collections.OrderedDict(
        [("a", NodeInfo({gpu: 8, "pods": 8}, False))])
    # 5 % per extra replica: not worth a restart one replica at a time, but
    # four more at once are (1.2 x 0.9 > 1)
    slow = lambda n, r: 1 + 0.05 * (r - 1) if r else 0.0         # noqa: E731
    jobs = collections.OrderedDict(
        [("j", JobInfo({gpu: 1, "pods": 1}, slow, 0, 0, 5))])
    alloc = {"j": ["a"]}
    PolluxPolicy._grow_into_idle(alloc, jobs, nodes, {"j": ["a"]}, ["a"])
    assert alloc["j"] == ["a"] * 5
    jobs["j"] = JobInfo({gpu: 1, "pods": 1}, slow, 0, 0, 2)
    alloc = {"j": ["a"]}
    PolluxPolicy._grow_into_idle(alloc, jobs, nodes, {"j": ["a"]}, ["a"])
    assert alloc["j"] == ["a"]                  # 1.05 x 0.9 < 1: stays
    alloc = {"j": ["a"]}                        # already being restarted
    PolluxPolicy._grow_into_idle(alloc, jobs, nodes, {}, ["a"])
    assert alloc["j"] == ["a", "a"]
