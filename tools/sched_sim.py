"""Cluster-scheduling simulator for the Pollux policy (CPU only).

The reference documents its scheduler with one chart -- average job completion
time under increasing load, AdaptDL vs a static gang scheduler
(``docs/README.rst:61``, values in BASELINE.md) -- produced on a real cluster.
This tool replays the same experiment against THIS repo's scheduler stack
(:class:`adaptdl_b200.sched.policy.PolluxPolicy`, :class:`SpeedupFunction`,
:class:`adaptdl_b200.goodput.GoodputFunction`) in simulated time:

* jobs arrive as a Poisson process; each is drawn from a small zoo of model
  profiles (a goodput model = throughput parameters + gradient-noise
  statistics that drift as training progresses, as they do in practice) and
  needs a fixed amount of *scale-invariant progress* to finish, expressed as
  the hours it would take on one GPU at its initial batch size (a heavy-tailed
  mix of 15-minute to 10-hour jobs);
* **adaptive**: every ``--interval`` seconds the policy re-allocates GPUs from
  the jobs' speedup functions; a job that is re-allocated pays a restart
  penalty; it trains at the batch size that maximises its goodput;
* **static** (baseline): every job asks for a fixed, per-model TUNED number of
  GPUs and keeps its initial batch size; FIFO with backfilling, no
  pre-emption. **static_whole_node**: the same scheduler when every user
  simply asks for one whole node (the untuned habit).

Reported per load level: average / p90 job completion time, restarts per job,
GPU-hours and node-hours held (what an autoscaled cloud cluster would bill: the
policy sizes the cluster to a utilisation band, ``pollux.py`` MIN/MAX_UTIL, so
it deliberately leaves nodes empty when the jobs' speedups have saturated),
and the policy's wall-clock cost per optimisation cycle.

    python tools/sched_sim.py --nodes 16 --gpus-per-node 4 --hours 8 \
        --rates 3,7,10,15 --out profiles/sched_sim.json
"""

import argparse
import collections
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptdl_b200.goodput import (GoodputFunction, GradParams,  # noqa: E402
                                  PerfParams)
from adaptdl_b200.sched.policy import (JobInfo, NodeInfo,  # noqa: E402
                                       PolluxPolicy, SpeedupFunction)

GPU = "nvidia.com/gpu"

# name -> (PerfParams, initial batch, max batch, local bsz bounds,
#          (grad sqr, var) at start, (sqr, var) at end, static GPU request)
# Throughput parameters are in the range this repo measures on B200 for the
# corresponding workloads (ResNet-18: ~2 ms / 128 samples; the slope / network
# terms follow the reference's fixture ratios, pollux_test.py:33-37).
ZOO = {
    "resnet18": (PerfParams(1.2e-3, 6.5e-6, 2.0e-4, 4.0e-5, 1.2e-4, 2.0e-5,
                            1.3), 128, 4096, (32, 1024),
                 (0.0014, 0.0005), (0.00025, 0.0011), 2),
    "bert": (PerfParams(4.0e-3, 2.4e-4, 1.6e-3, 3.0e-4, 9.0e-4, 1.5e-4, 1.2),
             32, 1024, (8, 128), (0.02, 0.004), (0.004, 0.012), 4),
    "ncf": (PerfParams(2.2e-4, 1.0e-7, 6.0e-5, 1.0e-5, 4.0e-5, 6.0e-6, 1.1),
            256, 32768, (128, 8192), (0.3, 0.05), (0.05, 0.2), 1),
    "transformer": (PerfParams(2.5e-3, 4.0e-5, 5.0e-4, 8.0e-5, 3.0e-4, 4.0e-5,
                               1.2), 20, 640, (5, 80),
                    (0.05, 0.01), (0.01, 0.04), 1),
}
RESTART_SECONDS = 30.0


class SimJob(object):
    def __init__(self, key, kind, arrival, rng):
        (self.perf, self.init_bsz, self.max_bsz, self.bounds, g0, g1,
         self.static_gpus) = ZOO[kind]
        self.key, self.kind, self.arrival = key, kind, arrival
        self.g0, self.g1 = g0, g1
        # work = single-GPU hours at the initial batch size (efficiency 1)
        hours = float(rng.choice([0.25, 1.0, 4.0, 10.0],
                                 p=[0.40, 0.35, 0.20, 0.05]))
        hours *= float(rng.uniform(0.7, 1.3))
        base = GoodputFunction(self.perf, GradParams(*g0), self.init_bsz)
        self.target = float(base.evaluate(1, 1, self.init_bsz, 0)) * \
            hours * 3600.0
        self.single_gpu_hours = hours
        self.progress = 0.0
        self.finish = None
        self.allocation = []
        self.penalty_until = 0.0
        self.restarts = 0
        self.max_profiled = 1
        self._speedup = None
        self._speedup_at = -1.0

    def grad_params(self):
        # noise scale grows as training converges (larger batches pay off late)
        f = min(self.progress / self.target, 1.0)
        sqr = self.g0[0] + f * (self.g1[0] - self.g0[0])
        var = self.g0[1] + f * (self.g1[1] - self.g0[1])
        return GradParams(sqr, var)

    def goodput_fn(self):
        return GoodputFunction(self.perf, self.grad_params(), self.init_bsz)

    def speedup_fn(self):
        bucket = int(10 * self.progress / self.target)
        if self._speedup is None or bucket != self._speedup_at:
            self._speedup = SpeedupFunction(
                self.goodput_fn(), self.max_bsz, self.bounds,
                accumulation=True)
            self._speedup_at = bucket
        return self._speedup

    def rate(self, adaptive):
        """scale-invariant samples per second with the current allocation"""
        n = len(self.allocation)
        if n == 0:
            return 0.0
        nodes = len(set(self.allocation))
        fn = self.goodput_fn()
        if adaptive:
            goodput, _, _ = fn.optimize(nodes, n, self.max_bsz, self.bounds,
                                        accumulation=True)
            return float(goodput)
        # static: fixed global batch = initial batch (split over the GPUs)
        atomic = max(-(-self.init_bsz // n), 1)       # ceil: never below it
        return float(fn.evaluate(nodes, n, atomic, 0))


def reference_policy(seed):
    """The UNMODIFIED reference policy (``baseline/ref_policy.py``: its
    pollux.py on a stand-in NSGA-II engine, population 100 x 100 generations
    as hard-coded there); ``None`` when the reference is not installed."""
    from baseline import ref_policy
    if not ref_policy.available():
        return None
    import logging
    policy_class = ref_policy.load()[0]
    logging.getLogger("adaptdl_sched.policy.pollux").setLevel(logging.WARNING)
    np.random.seed(seed)                 # the reference draws from the global
    return policy_class()


def simulate(rate_per_hour, args, adaptive, seed, static_gpus=None,
             policy_name="own"):
    rng = np.random.default_rng(seed)
    kinds = list(ZOO)
    weights = np.array([0.5, 0.1, 0.2, 0.2])
    horizon = args.hours * 3600.0
    arrivals, t = [], 0.0
    while True:
        t += rng.exponential(3600.0 / rate_per_hour)
        if t >= horizon:
            break
        arrivals.append(t)
    jobs = [SimJob("job-{}".format(i), kinds[rng.choice(len(kinds),
                                                        p=weights)], a, rng)
            for i, a in enumerate(arrivals)]
    if static_gpus is not None:          # every user asks for the same size
        for j in jobs:
            j.static_gpus = static_gpus
    nodes = {"node-{:02d}".format(i): NodeInfo({GPU: args.gpus_per_node,
                                                "pods": 32}, False)
             for i in range(args.nodes)}
    template = NodeInfo({GPU: args.gpus_per_node, "pods": 32}, True)
    if policy_name == "reference":
        policy = reference_policy(seed)
    else:
        policy = PolluxPolicy(pop_size=args.pop, generations=args.generations,
                              seed=seed)
    now, pending, active = 0.0, collections.deque(jobs), []
    dt = args.interval
    policy_seconds = 0.0
    gpu_seconds = 0.0            # GPU time held by jobs (what a cloud bills)
    nodes_seconds = 0.0          # nodes with at least one replica
    while (pending or active) and now < horizon * 6:
        while pending and pending[0].arrival <= now:
            active.append(pending.popleft())
        if adaptive and active:
            infos = {}
            for j in active:
                cap = max(2 * j.max_profiled, 2)
                infos[j.key] = JobInfo({GPU: 1, "pods": 1}, j.speedup_fn(),
                                       j.arrival, 0,
                                       min(cap, args.nodes *
                                           args.gpus_per_node))
            base = {j.key: list(j.allocation) for j in active}
            t0 = time.perf_counter()
            alloc, _ = policy.optimize(infos, nodes, base, template)
            policy_seconds += time.perf_counter() - t0
            for j in active:
                new = sorted(alloc.get(j.key, []))
                if new != sorted(j.allocation):
                    if j.allocation or j.progress > 0:
                        j.restarts += 1
                    j.allocation = new
                    j.penalty_until = now + RESTART_SECONDS
                j.max_profiled = max(j.max_profiled, len(new))
        elif active:
            free = {k: n.resources[GPU] for k, n in nodes.items()}
            for j in active:
                for node in j.allocation:
                    free[node] -= 1
            for j in sorted(active, key=lambda j: j.arrival):
                if j.allocation:
                    continue
                want, got = j.static_gpus, []
                for node in sorted(free, key=lambda k: -free[k]):
                    take = min(free[node], want - len(got))
                    got += [node] * take
                    if len(got) == want:
                        break
                if len(got) == want:            # gang: all or nothing
                    for node in got:
                        free[node] -= 1
                    j.allocation = got
                    j.penalty_until = now + RESTART_SECONDS
        gpu_seconds += dt * sum(len(j.allocation) for j in active)
        nodes_seconds += dt * len({n for j in active for n in j.allocation})
        for j in list(active):
            run = max(0.0, now + dt - max(now, j.penalty_until))
            j.progress += j.rate(adaptive) * run
            if j.progress >= j.target:
                j.finish = now + dt
                j.allocation = []
                active.remove(j)
        now += dt
    done = [j for j in jobs if j.finish is not None]
    jct = [j.finish - j.arrival for j in done]
    return {
        "jobs": len(jobs), "finished": len(done),
        "avg_jct_hours": float(np.mean(jct)) / 3600.0 if jct else None,
        "p90_jct_hours": float(np.percentile(jct, 90)) / 3600.0
        if jct else None,
        "makespan_hours": (max(j.finish for j in done) / 3600.0)
        if done else None,
        "restarts_per_job": float(np.mean([j.restarts for j in jobs]))
        if jobs else 0.0,
        "gpu_hours": gpu_seconds / 3600.0,
        "node_hours": nodes_seconds / 3600.0,
        "single_gpu_hours_of_work": float(sum(j.single_gpu_hours
                                              for j in jobs)),
        "policy_seconds_per_cycle": policy_seconds / max(now / dt, 1.0),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=16)
    ap.add_argument("--gpus-per-node", type=int, default=4)
    ap.add_argument("--hours", type=float, default=8.0,
                    help="length of the arrival window")
    ap.add_argument("--rates", default="3,7,10,15",
                    help="job submissions per hour")
    ap.add_argument("--interval", type=float, default=60.0)
    ap.add_argument("--pop", type=int, default=50)
    ap.add_argument("--generations", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--reference-policy", action="store_true",
                    help="add an arm that schedules with the reference's "
                         "unmodified policy (baseline/ref_policy.py)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rows = []
    for rate in (float(r) for r in args.rates.split(",")):
        row = {"jobs_per_hour": rate}
        for name, adaptive, fixed in (
                ("static", False, None),
                ("static_whole_node", False, args.gpus_per_node),
                ("adaptive", True, None)):
            row[name] = simulate(rate, args, adaptive, args.seed, fixed)
        if args.reference_policy and reference_policy(args.seed) is not None:
            row["reference_policy"] = simulate(rate, args, True, args.seed,
                                               None, "reference")
            r = row["reference_policy"]["avg_jct_hours"]
            own = row["adaptive"]["avg_jct_hours"]
            row["avg_jct_ratio_reference_policy_over_adaptive"] = \
                (r / own) if (r and own) else None
        a, s = row["adaptive"]["avg_jct_hours"], row["static"]["avg_jct_hours"]
        row["avg_jct_ratio_static_over_adaptive"] = \
            (s / a) if (a and s) else None
        w = row["static_whole_node"]["avg_jct_hours"]
        row["avg_jct_ratio_whole_node_over_adaptive"] = \
            (w / a) if (a and w) else None
        row["node_hours_ratio_static_over_adaptive"] = \
            row["static"]["node_hours"] / max(row["adaptive"]["node_hours"],
                                              1e-9)
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"config": vars(args), "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
