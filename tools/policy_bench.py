#!/usr/bin/env python
"""Cost of one Pollux optimisation cycle at cluster scale (CPU only).

    python tools/policy_bench.py --out profiles/policy_bench.json

The allocator runs the genetic search every 60 s over a ``[population, jobs,
2 x nodes]`` integer tensor (100 candidates x 100 generations, as in the
reference), so the cycle has to stay far below the period on the clusters the
scheduler is meant for. Jobs get randomised but realistic goodput models
(``SpeedupFunction`` over fitted-looking performance parameters); each
configuration runs ``--cycles`` warm-started cycles and reports the wall time
of each, the GPUs allocated and the sum of speedups of the chosen allocation
(the quantity the search maximises; useful to compare code versions).
``--search both`` runs the C++ core and the numpy implementation on the same
clusters.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from adaptdl_b200.goodput import GoodputFunction, GradParams, PerfParams  # noqa: E402
from adaptdl_b200.sched.policy import (JobInfo, NodeInfo, PolluxPolicy,  # noqa: E402
                                       SpeedupFunction)

GPU = "nvidia.com/gpu"


def make_cluster(num_jobs, num_nodes, gpus_per_node, seed):
    rng = np.random.RandomState(seed)
    jobs = {}
    for i in range(num_jobs):
        perf = PerfParams(0.1 * rng.uniform(.5, 2), 0.01 * rng.uniform(.5, 2),
                          0.05, 0.002, 0.02, 0.001, 1.2)
        grad = GradParams(sqr=rng.uniform(0.001, 0.1),
                          var=rng.uniform(0.01, 1.0))
        speedup = SpeedupFunction(
            GoodputFunction(perf, grad, 128), max_batch_size=4096,
            atomic_bsz_range=(32, 512), accumulation=True, mem_size=64)
        jobs["job-%d" % i] = JobInfo({GPU: 1, "pods": 1}, speedup, i, 0,
                                     min(64, 2 ** rng.randint(1, 7)))
    resources = {GPU: gpus_per_node, "pods": 32}
    nodes = {"node-%d" % i: NodeInfo(dict(resources), False)
             for i in range(num_nodes)}
    return jobs, nodes, NodeInfo(dict(resources), True)


def run(num_jobs, num_nodes, gpus_per_node, cycles, seed, native=None):
    jobs, nodes, template = make_cluster(num_jobs, num_nodes, gpus_per_node,
                                         seed)
    if native == "reference":
        # the reference's unmodified pollux.py on the stand-in NSGA-II engine
        # of baseline/shims/pymoo (baseline/ref_policy.py)
        import logging
        from baseline import ref_policy
        policy = ref_policy.load()[0]()
        logging.getLogger("adaptdl_sched.policy.pollux").setLevel(
            logging.WARNING)
        np.random.seed(seed)
        policy._native = None
    else:
        policy = PolluxPolicy(seed=seed, native=native)
    previous, seconds = {}, []
    for _ in range(cycles):
        start = time.perf_counter()
        allocations, desired = policy.optimize(jobs, nodes, previous,
                                               template)
        seconds.append(time.perf_counter() - start)
        previous = allocations
    per_node = {}
    for placement in allocations.values():
        for node in placement:
            per_node[node] = per_node.get(node, 0) + 1
    assert all(count <= gpus_per_node for count in per_node.values())
    value = sum(float(jobs[k].speedup_fn(len(set(a)), len(a)))
                for k, a in allocations.items() if a)
    return {"jobs": num_jobs, "nodes": num_nodes,
            "search": "reference policy (stand-in engine)"
            if policy._native is None
            else ("native" if policy._native else "numpy"),
            "gpus": num_nodes * gpus_per_node,
            "cycle_seconds": [round(s, 3) for s in seconds],
            "gpus_allocated": sum(per_node.values()),
            "jobs_running": sum(1 for a in allocations.values() if a),
            "sum_speedup": round(value, 2), "desired_nodes": int(desired)}


def main():
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("--sizes", default="16x4,50x16,100x32,200x64",
                        help="comma-separated JOBSxNODES")
    parser.add_argument("--gpus-per-node", type=int, default=8)
    parser.add_argument("--cycles", type=int, default=3)
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--search", default="native",
                        choices=["native", "numpy", "both", "reference"],
                        help="C++ core (csrc/host/adl_pollux.cpp), the numpy "
                             "implementation, or one after the other")
    parser.add_argument("--numpy-max-jobs", type=int, default=200,
                        help="with --search both: largest job count the "
                             "numpy search is run on")
    parser.add_argument("--out")
    args = parser.parse_args()
    rows = []
    for size in args.sizes.split(","):
        num_jobs, num_nodes = (int(v) for v in size.split("x"))
        flavours = {"native": [True], "numpy": [False],
                    "both": [True, False],
                    "reference": ["reference"]}[args.search]
        for native in flavours:
            if not native and args.search == "both" and \
                    num_jobs > args.numpy_max_jobs:
                continue
            row = run(num_jobs, num_nodes, args.gpus_per_node, args.cycles,
                      args.seed, native)
            rows.append(row)
            print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"config": vars(args), "cpu_count": os.cpu_count(),
                       "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
