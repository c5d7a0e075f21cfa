"""Several elastic jobs sharing the GPUs of ONE machine, scheduled by the
Pollux policy -- the cluster scheduler without a cluster.

On Kubernetes the allocator (``sched/allocator.py``) turns every job's
scheduling hints into a speedup function and lets :class:`PolluxPolicy`
divide the nodes' GPUs between the jobs once a minute; the controller then
restarts jobs whose allocation changed. This module does the same for an
8-GPU box: each job is a :class:`LocalElasticJob` (its own embedded
supervisor receives the hints), the "cluster" is one node with ``gpus``
devices, and a rescale is SIGTERM -> checkpoint -> exit 143 -> respawn on the
job's new set of device indices.

    python -m adaptdl_b200.sched.local_cluster --gpus 8 --interval 30 \\
        -- examples/pytorch-cifar/main.py --autoscale-bsz --epochs 30 \\
        -- examples/transformer/transformer.py --autoscale-bsz --epochs 20 \\
        -- examples/NCF/main.py --autoscale-bsz

(one ``--``-separated command per job; ``name=`` / ``min=`` / ``max=``
prefixes set a job's name and replica bounds).
"""

import argparse
import json
import logging
import os
import sys
import time

from adaptdl_b200.sched.allocator import job_info_from
from adaptdl_b200.sched.local import (LocalElasticJob,
                                      scratch_checkpoint_dir)
from adaptdl_b200.sched.policy import NodeInfo, PolluxPolicy

LOG = logging.getLogger(__name__)
GPU = "nvidia.com/gpu"
NODE = "localhost"


class LocalCluster(object):
    """``jobs``: list of dicts ``{"name", "command", "min_replicas",
    "max_replicas", "env"}``."""

    def __init__(self, jobs, gpus, checkpoint_root=None, policy=None,
                 resource=GPU):
        self.gpus = gpus
        self.resource = resource
        self.policy = policy or PolluxPolicy(pop_size=50, generations=40)
        root = checkpoint_root or scratch_checkpoint_dir("adaptdl-b200-lc-")
        self.jobs = {}
        self.specs = {}
        for i, spec in enumerate(jobs):
            name = spec.get("name") or "job-{}".format(i)
            ckpt = os.path.join(root, name)
            os.makedirs(ckpt, exist_ok=True)
            self.jobs[name] = LocalElasticJob(
                spec["command"], gpus, checkpoint_dir=ckpt,
                job_id="local/" + name, env=spec.get("env"))
            self.specs[name] = spec
        self.created = {name: time.time() + 1e-3 * i
                        for i, name in enumerate(self.jobs)}
        self.devices = {name: [] for name in self.jobs}   # held GPU indices
        self.done = {}
        self.events = []

    # -- the allocator's view of a job ---------------------------------------

    def _job_object(self, name):
        spec, job = self.specs[name], self.jobs[name]
        return {
            "metadata": {"name": name,
                         "creationTimestamp": self.created[name]},
            "spec": {
                "minReplicas": spec.get("min_replicas", 0),
                "maxReplicas": spec.get("max_replicas", self.gpus),
                "template": {"spec": {"containers": [{
                    "name": "main",
                    "resources": {"limits": {self.resource: 1}}}]}},
            },
            "status": {"train": job.server.hints or {}},
        }

    def _assign_devices(self, targets):
        """Device indices per job for the new replica counts: jobs keep the
        devices they hold where possible (no restart if nothing changes)."""
        keep = {name: self.devices[name][:targets.get(name, 0)]
                for name in self.jobs}
        used = {d for ids in keep.values() for d in ids}
        free = [d for d in range(self.gpus) if d not in used]
        out = {}
        for name in self.jobs:
            ids = list(keep[name])
            while len(ids) < targets.get(name, 0):
                ids.append(free.pop(0))
            out[name] = ids
        return out

    # -- one scheduling cycle ------------------------------------------------

    def step(self):
        """Reap finished jobs, run the policy, apply the new allocation."""
        for name, job in self.jobs.items():
            if name in self.done or not job.procs:
                continue
            state = job.poll()
            if state in ("finished", "failed"):
                self.done[name] = state
                self.devices[name] = []
                self._log(state, job=name)
            elif state == "preempted":          # left on its own signal
                job.num_restarts += 1
                job.start(len(self.devices[name]), self.devices[name])
        active = [n for n in self.jobs if n not in self.done]
        if not active:
            return False
        infos = {n: job_info_from(self._job_object(n)) for n in active}
        nodes = {NODE: NodeInfo({self.resource: self.gpus, "pods": 1000},
                                False)}
        base = {n: [NODE] * len(self.devices[n]) for n in active}
        alloc, _ = self.policy.optimize(infos, nodes, base, nodes[NODE])
        targets = {n: len(alloc.get(n, [])) for n in active}
        devices = self._assign_devices(targets)
        # shrink first so that growing jobs find their devices free
        order = sorted(active, key=lambda n: targets[n]
                       - len(self.devices[n]))
        for name in order:
            job, want = self.jobs[name], devices[name]
            if want == self.devices[name] and (job.procs or not want):
                continue
            self._log("allocate", job=name, replicas=len(want), devices=want)
            if job.procs:
                state = job.rescale(len(want), gpu_ids=want)
                if state in ("finished", "failed"):
                    self.done[name] = state
                    want = []
            elif want:
                job.start(len(want), want)
            self.devices[name] = list(want)
        return True

    def run(self, interval=30.0, timeout=None):
        deadline = time.time() + timeout if timeout else None
        try:
            while self.step():
                if deadline is not None and time.time() > deadline:
                    self._log("timeout")
                    break
                t_next = time.time() + interval
                while time.time() < t_next:
                    time.sleep(0.2)
                    if all((n in self.done) or
                           (self.jobs[n].procs and
                            all(p.poll() is not None
                                for p in self.jobs[n].procs))
                           for n in self.jobs):
                        break               # everything exited: reap now
        finally:
            for job in self.jobs.values():
                job.kill()
                job.server.close()
        return dict(self.done)

    def _log(self, what, **detail):
        self.events.append((time.time(), what, detail))
        LOG.info("%s %s", what, detail)


def _parse_jobs(argv):
    jobs, cur = [], None
    for tok in argv:
        if tok == "--":
            cur = {"command": []}
            jobs.append(cur)
        elif cur is None:
            raise SystemExit("jobs must be introduced by '--'")
        elif not cur["command"] and "=" in tok and \
                tok.split("=")[0] in ("name", "min", "max"):
            key, value = tok.split("=", 1)
            cur[{"name": "name", "min": "min_replicas",
                 "max": "max_replicas"}[key]] = \
                value if key == "name" else int(value)
        else:
            cur["command"].append(tok)
    for job in jobs:
        if not job["command"]:
            raise SystemExit("empty job command")
        if job["command"][0].endswith(".py"):
            job["command"] = [sys.executable] + job["command"]
    return jobs


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    split = argv.index("--") if "--" in argv else len(argv)
    parser = argparse.ArgumentParser(
        description="schedule several elastic jobs on this box's GPUs")
    parser.add_argument("--gpus", type=int, default=None)
    parser.add_argument("--interval", type=float, default=30.0)
    parser.add_argument("--timeout", type=float, default=None)
    parser.add_argument("--checkpoint-root", default=None)
    parser.add_argument("--report", default=None)
    args = parser.parse_args(argv[:split])
    logging.basicConfig(level=logging.INFO)
    jobs = _parse_jobs(argv[split:])
    if not jobs:
        parser.error("no jobs given")
    gpus = args.gpus
    if gpus is None:
        import torch
        gpus = max(torch.cuda.device_count(), 1)
    cluster = LocalCluster(jobs, gpus, args.checkpoint_root)
    done = cluster.run(args.interval, args.timeout)
    if args.report:
        with open(args.report, "w") as f:
            json.dump({"result": done,
                       "events": [{"t": t, "event": w, **d}
                                  for t, w, d in cluster.events]}, f,
                      indent=1)
    return 0 if done and all(v == "finished" for v in done.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
