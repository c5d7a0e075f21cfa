"""Pollux: goodput-driven co-adaptive allocation of replicas to nodes.

Capabilities match the reference's ``sched/adaptdl_sched/policy/pollux.py``
(``allocate_job`` first-fit for new arrivals; ``optimize`` = multi-objective
genetic search over job x node replica matrices, maximising the sum of
dominant-share-scaled speedups and minimising the number of nodes used, with
a restart penalty, warm start from the previous cycle, "virtual" nodes for
cluster autoscaling and a utilisation band that picks the desired cluster
size) -- but implemented on an in-house NSGA-II rather than on pymoo, with
the problem's operators written against this module's own state layout. The
search exists twice: vectorised numpy (``nsga2.py`` + the operators of
``ClusterProblem`` below, the executable specification) and the C++ core
``csrc/host/adl_pollux.cpp`` (``native.py``), which is what runs when the
host library can be built or loaded.

State layout: ``states[p, j, n]`` = replicas of job ``j`` on node ``n`` in
candidate ``p``. Nodes ``[0, N)`` are the real nodes (non-preemptible first),
nodes ``[N, 2N)`` are copies of the autoscaling node template ("what if the
cluster were bigger").
"""

import collections
import copy
import logging
import os

import numpy as np

from adaptdl_b200.sched.policy import native as native_search
from adaptdl_b200.sched.policy import nsga2

LOG = logging.getLogger(__name__)

POP_SIZE = 100
GENERATIONS = 100
RESTART_PENALTY = 0.1
MIN_UTIL, MAX_UTIL = 0.35, 0.65
# replica counts per (job, node): the population tensor is the bulk of the
# memory traffic of a cycle, resource arithmetic is done in int64
STATE_DTYPE = np.int32


def _sorted_nodes(nodes):
    """Non-preemptible nodes first, then by key."""
    return collections.OrderedDict(
        sorted(nodes.items(), key=lambda kv: (kv[1].preemptible, kv[0])))


class ClusterProblem(object):
    """Objectives and genetic operators for one optimisation cycle."""

    def __init__(self, jobs, nodes, base_state,
                 restart_penalty=RESTART_PENALTY, rng=None):
        assert base_state.shape == (len(jobs), len(nodes))
        self._rng = np.random.default_rng() if rng is None else rng
        self.jobs, self.nodes = jobs, nodes
        self.base = base_state
        self.restart_penalty = restart_penalty
        J, N = base_state.shape
        self.pinned = np.array(
            [j for j, job in enumerate(jobs)
             if not job.preemptible and base_state[j].any()], dtype=int)
        rtypes = sorted(set().union(*[set(job.resources) for job in jobs]))
        self.job_res = np.array(
            [[job.resources.get(r, 0) for r in rtypes] for job in jobs],
            dtype=np.int64)                                    # [J, R]
        self.node_res = np.array(
            [[node.resources.get(r, 0) for r in rtypes] for node in nodes],
            dtype=np.int64)                                    # [N, R]
        with np.errstate(divide="ignore", invalid="ignore"):
            share = self.job_res / self.node_res.sum(axis=0)
        self.dominant_share = np.nan_to_num(share, nan=0.0,
                                            posinf=0.0).max(axis=1)
        # capacity left on each node once pinned jobs are accounted for
        pinned_use = np.einsum("jn,jr->nr", base_state[self.pinned],
                               self.job_res[self.pinned]) \
            if len(self.pinned) else np.zeros_like(self.node_res)
        free = self.node_res - pinned_use
        assert (free >= 0).all()
        # most replicas of job j that fit on node n by itself
        with np.errstate(divide="ignore"):
            fit = np.where(self.job_res[:, None, :] > 0,
                           free[None, :, :] // np.maximum(
                               self.job_res[:, None, :], 1),
                           np.iinfo(np.int64).max)
        self.max_fit = fit.min(axis=2).astype(np.int64)        # [J, N]
        self.max_fit = np.minimum(self.max_fit, 1 << 20)
        # spread each job's min_replicas greedily over nodes: the lower
        # bound used by mutation
        self.min_fill = np.zeros((J, N), dtype=np.int64)
        for j, job in enumerate(jobs):
            need = job.min_replicas
            for n in range(N):
                take = min(need, int(self.max_fit[j, n]))
                self.min_fill[j, n] = take
                need -= take
        self.max_replicas = np.array([[job.max_replicas] for job in jobs])
        self.min_replicas = np.array([job.min_replicas for job in jobs])

    # -- objectives --------------------------------------------------------

    def speedups(self, states):
        nodes_used = np.count_nonzero(states, axis=2)
        replicas = states.sum(axis=2, dtype=np.int64)
        # SpeedupFunction objects answer from their table without the
        # argument checking of the public call (tens of thousands of calls
        # per cycle); any other callable is called as is
        cols = [getattr(job.speedup_fn, "lookup", job.speedup_fn)(
                    nodes_used[:, j], replicas[:, j])
                for j, job in enumerate(self.jobs)]
        return np.stack(cols, axis=1).astype(float)

    def cluster_sizes(self, states):
        """Index of the last used node + 1 (nodes are in preference order)."""
        used = states.any(axis=-2)
        ordinal = np.arange(states.shape[-1]) + 1
        return np.where(used, ordinal, 0).max(axis=-1)

    def evaluate(self, states):
        speedup = self.speedups(states)
        # a dominant share worth one node <=> speedup 1
        scaled = speedup * self.dominant_share * len(self.nodes)
        moved = (states != self.base).any(axis=2)
        scaled = np.where(moved, scaled * (1.0 - self.restart_penalty),
                          scaled)
        return np.column_stack([-scaled.sum(axis=1),
                                self.cluster_sizes(states)])

    def utilities(self, states):
        """Average fraction of ideal scaling (speedup / replicas) weighted by
        each job's share of the most contended resource of the nodes in
        use."""
        replicas = states.sum(axis=2)
        speedup = self.speedups(states)
        active = states.sum(axis=1) > 0                        # [P, N]
        total = (active[:, :, None] * self.node_res).sum(axis=1)  # [P, R]
        alloc = replicas[:, :, None] * self.job_res            # [P, J, R]
        with np.errstate(divide="ignore", invalid="ignore"):
            share = np.where(alloc > 0, alloc / total[:, None, :], 0.0)
            ideal = np.where(replicas > 0, speedup / replicas, 0.0)
        return (ideal[:, :, None] * share).sum(axis=1).max(axis=1)

    # -- genetic operators ---------------------------------------------------

    def crossover(self, a, b, rng):
        """Single cut over the job axis; children also inherit a cluster
        size drawn between the parents' sizes (nodes beyond it emptied)."""
        pairs, J, N = a.shape
        cut = rng.integers(J, size=(pairs, 1, 1))
        take_a = np.arange(J)[None, :, None] < cut
        children = np.concatenate([np.where(take_a, a, b),
                                   np.where(take_a, b, a)])
        size_a, size_b = self.cluster_sizes(a), self.cluster_sizes(b)
        lo, hi = np.minimum(size_a, size_b), np.maximum(size_a, size_b)
        lo, hi = np.tile(lo, 2), np.tile(hi, 2)
        size = lo + rng.integers(1 << 15, size=len(children)) % (hi - lo + 1)
        beyond = np.arange(N)[None, None, :] >= size[:, None, None]
        return np.where(beyond, 0, children)

    def mutate(self, states, rng):
        """Re-draw a few entries within their feasible range; zero and
        non-zero entries of a row are equally likely to be touched."""
        states = np.asarray(states)
        width = states.shape[2]
        positive = states > 0
        nonzero = positive.sum(axis=2, keepdims=True, dtype=np.int32)
        with np.errstate(divide="ignore"):            # per-row thresholds
            p_nonzero = (1.0 / nonzero).astype(np.float32)
            p_zero = (1.0 / (width - nonzero)).astype(np.float32)
        hit = rng.random(states.shape, dtype=np.float32) < \
            np.where(positive, p_nonzero, p_zero)
        # a mutation grows the cluster by a geometrically distributed number
        # of nodes (usually one): unrestricted hits in far columns would erase
        # every small-cluster candidate and the Pareto front would lose the
        # sizes that fit the real nodes
        ordinal = np.arange(width)
        size = np.where(positive.any(axis=1), ordinal + 1, 0).max(axis=1)
        limit = size + rng.geometric(0.5, size=len(states))
        hit &= (ordinal[None, :] < limit[:, None])[:, None, :]
        # new values are drawn only where a mutation happens (a few entries
        # per row), not for the whole population tensor
        p, j, n = np.nonzero(hit)
        out = np.array(states, dtype=STATE_DTYPE, copy=True)
        out[p, j, n] = rng.integers(self.min_fill[j, n],
                                    self.max_fit[j, n] + 1)
        return np.maximum(out, self.min_fill, out=out)

    def repair(self, states):
        """Make every candidate feasible. The population tensor is
        ``[P, J, N]`` with thousands of jobs x nodes entries per candidate
        and this runs once per generation, so each rule touches only what it
        must (rows over their cap, resources somebody requests) and works
        resource by resource instead of on a ``[P, J, N, R]`` tensor."""
        states = np.array(states, dtype=STATE_DTYPE, copy=True)
        # 1. non-preemptible jobs that already run keep their placement
        if len(self.pinned):
            states[:, self.pinned] = self.base[self.pinned]
        # 2. a node hosts at most one multi-node job (first in job order);
        #    only nodes with several of them are looked at
        positive = states > 0
        spread = positive.sum(axis=2, dtype=np.int32) > 1       # [P, J]
        positive &= spread[:, :, None]
        crowd_p, crowd_n = np.nonzero(
            positive.sum(axis=1, dtype=np.int32) > 1)
        if len(crowd_p):
            cols = positive[crowd_p, :, crowd_n]                # [K, J]
            later = cols & (np.cumsum(cols, axis=1, dtype=np.int32) > 1)
            kept = states[crowd_p, :, crowd_n]
            kept[later] = 0
            states[crowd_p, :, crowd_n] = kept
        # 3. no more than max_replicas per job: rows above their cap are
        #    trimmed in a random node order
        over_p, over_j = np.nonzero(
            states.sum(axis=2) > self.max_replicas[:, 0])
        if len(over_p):
            rows = states[over_p, over_j]                       # [K, N]
            order = np.argsort(self._rng.random(rows.shape), axis=1)
            shuffled = np.take_along_axis(rows, order, axis=1)
            capped = np.minimum(shuffled.cumsum(axis=1),
                                self.max_replicas[over_j])
            trimmed = np.empty_like(rows)
            np.put_along_axis(trimmed, order,
                              np.diff(capped, axis=1, prepend=0), axis=1)
            states[over_p, over_j] = trimmed
        # 4. node capacities, one resource after the other: on every
        #    oversubscribed (candidate, node) the jobs claim the resource in
        #    priority order and keep what it can still grant. Columns within
        #    capacity (most of them for every resource but the scarcest) are
        #    not touched.
        for r in range(self.job_res.shape[1]):
            need = self.job_res[:, r]
            if not need.any():
                continue
            capacity = self.node_res[:, r]
            used = np.einsum("pjn,j->pn", states, need)
            over_p, over_n = np.nonzero(used > capacity)
            if not len(over_p):
                continue
            cols = states[over_p, :, over_n]                    # [K, J]
            claim = cols * need
            np.cumsum(claim, axis=1, out=claim)
            np.minimum(claim, capacity[over_n][:, None], out=claim)
            granted = np.diff(claim, axis=1, prepend=0)
            granted //= np.maximum(need, 1)
            states[over_p, :, over_n] = np.where(need > 0,
                                                np.minimum(cols, granted),
                                                cols)
        # 5. all-or-nothing below min_replicas
        short = states.sum(axis=2) < self.min_replicas
        states[short] = 0
        return states


class PolluxPolicy(object):

    def __init__(self, pop_size=POP_SIZE, generations=GENERATIONS, seed=None,
                 native=None):
        """``native``: run the search on the C++ core
        (``csrc/host/adl_pollux.cpp``, ~20x faster at cluster scale) rather
        than on numpy. ``None`` = when it can be loaded, unless
        ``ADAPTDL_B200_NATIVE_POLICY=0``."""
        self._pop_size = pop_size
        self._generations = generations
        self._rng = np.random.default_rng(seed)
        if native is None:
            native = os.environ.get("ADAPTDL_B200_NATIVE_POLICY", "1") != "0"
        self._native = bool(native) and native_search.available()
        self._grow = os.environ.get("ADAPTDL_B200_POLICY_GROW", "1") != "0"
        self._prev_states = None
        self._prev_jobs = None
        self._prev_nodes = None
        self._min_util = MIN_UTIL
        self._max_util = MAX_UTIL

    # -- single-job fast path ------------------------------------------------

    def allocate_job(self, job_info, nodes):
        """First node (non-preemptible first) that fits ``min_replicas`` (at
        least one) replicas of a newly arrived job; ``[]`` if none does.
        ``nodes`` must already account for every running pod."""
        want = max(job_info.min_replicas, 1)
        for name, node in _sorted_nodes(nodes).items():
            fits = min(node.resources.get(key, 0) // val
                       for key, val in job_info.resources.items())
            if fits >= want:
                return [name] * want
        return []

    # -- state <-> allocation ------------------------------------------------

    @staticmethod
    def _allocations_to_state(allocations, jobs, nodes):
        job_idx = {key: i for i, key in enumerate(jobs)}
        node_idx = {key: i for i, key in enumerate(nodes)}
        state = np.zeros((len(jobs), len(nodes)), dtype=np.int64)
        for key, alloc in allocations.items():
            if key not in job_idx:
                continue
            for node in alloc:
                if node in node_idx:
                    state[job_idx[key], node_idx[node]] += 1
        return state

    @staticmethod
    def _state_to_allocations(state, jobs, nodes):
        node_keys = list(nodes)
        out = {key: [] for key in jobs}
        job_keys = list(jobs)
        rows, cols = np.nonzero(state)        # row-major: nodes in order
        for j, n in zip(rows.tolist(), cols.tolist()):
            out[job_keys[j]].extend([node_keys[n]] * int(state[j, n]))
        return out

    def _warm_start(self, jobs, nodes):
        """Re-index last cycle's final population onto this cycle's jobs and
        nodes; nodes that appeared since take over a virtual node's column."""
        prev = self._prev_states
        P = prev.shape[0]
        N = len(nodes)
        out = np.zeros((P, len(jobs), 2 * N), dtype=STATE_DTYPE)
        prev_job = {key: i for i, key in enumerate(self._prev_jobs)}
        src_rows = [prev_job[k] for k in jobs if k in prev_job]
        dst_rows = [i for i, k in enumerate(jobs) if k in prev_job]
        if not src_rows:
            return out
        prev_node = {key: i for i, key in enumerate(self._prev_nodes)}
        spare = len(self._prev_nodes)         # next unused virtual column
        width = prev.shape[2]
        node_keys = list(nodes)
        src_cols, dst_cols = [], []
        for col in range(2 * N):
            key = node_keys[col] if col < N else None
            if key is not None and key in prev_node:
                src = prev_node[key]
            elif spare < width:
                src = spare
                spare += 1
            else:
                continue
            src_cols.append(src)
            dst_cols.append(col)
        # one gather / scatter for the whole population
        out[:, np.array(dst_rows)[:, None], np.array(dst_cols)[None, :]] = \
            prev[:, np.array(src_rows)[:, None], np.array(src_cols)[None, :]]
        return out

    # -- choosing from the Pareto front -------------------------------------------

    @staticmethod
    def _best_within(values, max_nodes):
        ok = values[:, 1] <= max_nodes
        if not ok.any():
            return None
        return int(np.argmin(np.where(ok, values[:, 0], 0.0)))

    def _desired_nodes(self, utilities, values, num_nodes):
        """Cluster size whose utilisation sits closest to the middle of the
        band, unless the best allocation on the present nodes is already
        inside it. The search for that size only looks in the direction the
        present utilisation points to: an over-used cluster never shrinks
        and an under-used one never grows. (The reference compares every
        size on the Pareto front, ``pollux.py:121-142``: with hundreds of
        jobs a two-node candidate that happens to sit at 50 % wins against
        the 80 % of the full cluster, and since the allocation is then
        chosen within the desired size, a cycle evicts nearly every job --
        seen at 1000 jobs x 256 nodes with ``tools/policy_bench.py``.)"""
        idx = self._best_within(values, num_nodes)
        if idx is not None and \
                self._min_util <= utilities[idx] <= self._max_util:
            return num_nodes
        grow = idx is not None and utilities[idx] > self._max_util
        shrink = idx is not None and utilities[idx] < self._min_util
        target = (self._min_util + self._max_util) / 2
        best_util, best_nodes = np.inf, num_nodes
        for util, (_, size) in zip(utilities, values):
            if util < self._min_util:
                continue
            if (grow and size < num_nodes) or (shrink and size > num_nodes):
                continue
            if np.isclose(util, best_util) and size > best_nodes:
                best_nodes = size
            if abs(util - target) < abs(best_util - target):
                best_util, best_nodes = util, size
        return int(best_nodes)

    # -- the optimisation cycle ------------------------------------------------

    def optimize(self, jobs, nodes, base_allocations, node_template):
        """One scheduling cycle.

        Arguments:
            jobs (dict): job key -> :class:`JobInfo` of every incomplete job.
            nodes (dict): node key -> :class:`NodeInfo`; resources net of
                non-adaptdl pods only.
            base_allocations (dict): job key -> current allocation (list with
                one node key per replica).
            node_template (NodeInfo): a node the cluster autoscaler could add.

        Returns ``(allocations, desired_nodes)``.
        """
        def pinned(key, job):
            return not job.preemptible and bool(base_allocations.get(key))

        # pinned jobs first (their placement is fixed), then fewer guaranteed
        # replicas first, then FIFO
        jobs = collections.OrderedDict(sorted(
            jobs.items(), key=lambda kv: (not pinned(kv[0], kv[1]),
                                          kv[1].min_replicas,
                                          kv[1].creation_timestamp)))
        nodes = _sorted_nodes(nodes)
        N = len(nodes)
        base = np.concatenate(
            [self._allocations_to_state(base_allocations, jobs, nodes),
             np.zeros((len(jobs), N), dtype=np.int64)], axis=1)
        if self._prev_states is None:
            initial = base[None]
        else:
            initial = np.concatenate([self._warm_start(jobs, nodes),
                                      base[None]])
        problem = ClusterProblem(list(jobs.values()),
                                 list(nodes.values()) + [node_template] * N,
                                 base, rng=self._rng)
        search = native_search.minimize if self._native else nsga2.minimize
        states, values = search(problem, initial, self._pop_size,
                                self._generations, self._rng)
        self._prev_states = states       # a fresh array: ours to keep
        self._prev_jobs = list(jobs)
        self._prev_nodes = list(nodes)
        front = nsga2.non_dominated_fronts(values)[0]
        states, values = states[front], values[front]
        utilities = problem.utilities(states)
        desired = self._desired_nodes(utilities, values, N)
        idx = self._best_within(values, min(N, desired))
        if LOG.isEnabledFor(logging.DEBUG):
            for i, state in enumerate(states):
                LOG.debug("solution %d value=%s utility=%.3f\n%s", i,
                          values[i].tolist(), utilities[i], state)
        if idx is None:
            return {}, desired
        allocations = self._state_to_allocations(states[idx][:, :N], jobs,
                                                 nodes)
        self._place_starved(allocations, jobs, nodes)
        if self._grow:
            self._grow_into_idle(allocations, jobs, nodes, base_allocations,
                                 list(nodes)[:min(N, desired)],
                                 RESTART_PENALTY)
        return allocations, desired

    @staticmethod
    def _place_starved(allocations, jobs, nodes):
        """No job waits while capacity sits idle. The genetic search limits
        itself to the desired cluster size (utilisation band), which on a
        busy, fixed-size cluster can leave recently arrived jobs without a
        single replica next to free GPUs. Give each such job its minimum
        (at least one replica) on one node, preferring nodes that are already
        in use so that empty nodes stay releasable. This goes beyond the
        reference policy (``pollux.py:200-215`` returns the selected state
        as is); a job at one replica has speedup 1 instead of 0 and nobody
        else's allocation changes, so the cycle's utility only improves."""
        free = {key: dict(node.resources) for key, node in nodes.items()}
        for key, placement in allocations.items():
            for node in placement:
                for rtype, amount in jobs[key].resources.items():
                    free[node][rtype] = free[node].get(rtype, 0) - amount
        in_use = {node for placement in allocations.values()
                  for node in placement}
        for key, job in jobs.items():           # policy order (FIFO inside)
            if allocations.get(key):
                continue
            want = max(job.min_replicas, 1)
            for name in sorted(free, key=lambda n: n not in in_use):
                fits = min((free[name].get(rtype, 0) // amount
                            for rtype, amount in job.resources.items()
                            if amount > 0), default=0)
                if fits >= want:
                    allocations[key] = [name] * want
                    for rtype, amount in job.resources.items():
                        free[name][rtype] -= amount * want
                    in_use.add(name)
                    break

    @staticmethod
    def _grow_into_idle(allocations, jobs, nodes, base_allocations,
                        usable_nodes, restart_penalty=RESTART_PENALTY):
        """Hand idle capacity of the nodes the cycle may use to the jobs that
        gain most from it, one replica at a time (lazy greedy on the marginal
        gain of the search's own objective: speedup x dominant share, times
        ``1 - restart_penalty`` for a job whose allocation differs from the
        one it runs with).

        The genetic search re-draws about two entries of EVERY job in every
        child (the reference's mutation too, ``pollux.py:377-392``). With tens
        of jobs that explores well; with hundreds, no child is a small step
        away from a good allocation any more, the restart penalty on every
        job outweighs what a few lucky re-draws win, and the search returns
        its starting point with a third of the GPUs idle (1000 jobs x 256
        nodes, ``tools/policy_bench.py``). This pass is the small-step
        improvement the search cannot make at that scale; at small scale it
        picks up the last idle GPUs. It never takes anything away, never
        touches a pinned job, keeps a node to one multi-node job and stays
        inside ``usable_nodes`` (the cluster size the cycle decided on)."""
        import heapq
        usable = [n for n in usable_nodes if n in nodes]
        if not usable:
            return
        free = {key: dict(node.resources) for key, node in nodes.items()}
        placed = {key: collections.Counter(allocations.get(key) or [])
                  for key in jobs}
        for key, count in placed.items():
            for node, reps in count.items():
                for rtype, amount in jobs[key].resources.items():
                    free[node][rtype] = free[node].get(rtype, 0) - \
                        amount * reps
        # nodes that host a multi-node job (rule: at most one per node)
        spread_on = {}
        for key, count in placed.items():
            if len(count) > 1:
                for node in count:
                    spread_on[node] = key
        totals = collections.Counter()
        for node in nodes.values():
            totals.update(node.resources)
        in_use = {node for count in placed.values() for node in count}

        def weight(job):
            return max((amount / totals[rtype]
                        for rtype, amount in job.resources.items()
                        if totals.get(rtype, 0) > 0), default=0.0)

        def fits(job, node, reps=1):
            return all(free[node].get(rtype, 0) >= amount * reps
                       for rtype, amount in job.resources.items()
                       if amount > 0)

        def value(key, num_nodes, replicas, moved):
            if replicas == 0:
                return 0.0
            speedup = float(jobs[key].speedup_fn(num_nodes, replicas))
            return speedup * (1.0 - restart_penalty if moved else 1.0)

        moved = {key: placed[key] != collections.Counter(
                     base_allocations.get(key) or []) for key in jobs}
        weights = {}

        def is_moved(key):
            return moved[key]

        def proposal(key):
            """Best single step for this job: ``(gain, node, replicas)``."""
            job, count = jobs[key], placed[key]
            replicas = sum(count.values())
            if replicas == 0 or replicas >= job.max_replicas:
                return None
            if not job.preemptible and base_allocations.get(key):
                return None                      # pinned
            now = value(key, len(count), replicas, is_moved(key))
            best = None
            own = [n for n in count if fits(job, n)]
            if own:
                node = max(own, key=lambda n: count[n])
                steps = [1]
                if not is_moved(key):
                    # a job that has not been disturbed yet pays the restart
                    # penalty once: also look at the largest jump
                    room = min((free[node].get(rtype, 0) // amount
                                for rtype, amount in job.resources.items()
                                if amount > 0), default=0)
                    steps.append(min(room, job.max_replicas - replicas))
                for reps in sorted(set(k for k in steps if k >= 1)):
                    gain = value(key, len(count), replicas + reps, True) - now
                    if best is None or gain > best[0]:
                        best = (gain, node, reps)
            else:
                # a further node: neither end may break the one-multi-node-
                # job-per-node rule
                if any(spread_on.get(n, key) != key for n in count):
                    return None
                for node in sorted(usable, key=lambda n: n not in in_use):
                    if node in count or node in spread_on or \
                            not fits(job, node):
                        continue
                    gain = value(key, len(count) + 1, replicas + 1, True) - now
                    best = (gain, node, 1)
                    break
            if best is None or best[0] <= 0:
                return None
            if key not in weights:
                weights[key] = weight(job)
            return (best[0] * weights[key], best[1], best[2])

        heap = []
        order = {key: i for i, key in enumerate(jobs)}
        for key in jobs:
            found = proposal(key)
            if found:
                heapq.heappush(heap, (-found[0], order[key], key, found[1],
                                      found[2]))
        while heap:
            neg_gain, _, key, node, reps = heapq.heappop(heap)
            fresh = proposal(key)
            if fresh is None:
                continue
            if fresh[1] != node or fresh[2] != reps or \
                    fresh[0] < -neg_gain - 1e-12:
                # the node filled up meanwhile: queue the job's current best
                heapq.heappush(heap, (-fresh[0], order[key], key, fresh[1],
                                      fresh[2]))
                continue
            job, count = jobs[key], placed[key]
            count[node] += reps
            moved[key] = True
            for rtype, amount in job.resources.items():
                free[node][rtype] = free[node].get(rtype, 0) - amount * reps
            in_use.add(node)
            if len(count) > 1:
                for n in count:
                    spread_on[n] = key
            again = proposal(key)
            if again:
                heapq.heappush(heap, (-again[0], order[key], key, again[1],
                                      again[2]))
        node_order = {key: i for i, key in enumerate(nodes)}
        for key, count in placed.items():
            if sum(count.values()) != len(allocations.get(key) or []):
                allocations[key] = [
                    node for node in sorted(count, key=node_order.get)
                    for _ in range(count[node])]


_ = copy
