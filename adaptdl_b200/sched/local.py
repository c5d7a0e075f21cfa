"""Single-box elastic launcher: AdaptDL scheduling without Kubernetes.

Runs ONE elastic job on the GPUs of this machine. Each replica is a process
with the same ``ADAPTDL_*`` contract a pod would get
(:func:`adaptdl_b200.sched.controller.build_pod`); a rescale is what it is
on a cluster -- SIGTERM every replica, they agree on the step, checkpoint and
exit with code 143, and the next generation starts at the new replica count
with ``ADAPTDL_NUM_RESTARTS + 1`` (peer GPU mappings are rebuilt by the new
processes at the new world size).

The replica count comes either from a fixed schedule (``--schedule
2,4,8,4 --interval 20``: benchmark config "elastic rescale 2->4->8->4") or
from the job's own scheduling hints: an embedded supervisor receives
``PUT /hints``, builds the goodput speedup function and picks the count that
maximises speedup among the GPUs available (the single-job specialisation of
the Pollux policy; 5 % hysteresis like the reference's Ray-AWS optimiser,
``ray/adaptdl_ray/aws/optimizer.py:70-94``).

    python -m adaptdl_b200.sched.local --gpus 8 --schedule 2,4,8,4 \
        --interval 20 examples/transformer/transformer.py --epochs 5
"""

import argparse
import atexit
import json
import logging
import os
import shutil
import signal
import subprocess
import sys
import tempfile
import threading
import time

from adaptdl_b200.goodput import GoodputFunction, GradParams, PerfParams
from adaptdl_b200.sched.policy import SpeedupFunction
from adaptdl_b200.sched_hints import PERF_PARAMS, SCHED_HINTS
from adaptdl_b200.utils import pick_unused_port

LOG = logging.getLogger(__name__)
EXIT_PREEMPTED = 143


class HintsServer(object):
    """Tiny supervisor: ``PUT /hints/<job>``, ``GET /discover/...`` (all
    replicas are local, so discovery always answers 127.0.0.1) and
    ``GET /metrics`` (Prometheus series derived from the hints)."""

    def __init__(self):
        from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
        outer = self
        self.hints = None
        self.replicas = 1

        class Handler(BaseHTTPRequestHandler):
            def log_message(self, *args):     # quiet
                pass

            def do_GET(self):
                if self.path.startswith("/discover/"):
                    body = json.dumps(["127.0.0.1"] * outer.replicas)
                    self._reply(200, body)
                elif self.path == "/metrics":     # same series as the
                    from adaptdl_b200.sched import metrics  # K8s supervisor
                    body, ctype = metrics.render()
                    self._reply(200, body.decode(), ctype)
                else:
                    self._reply(200, "")

            def do_PUT(self):
                length = int(self.headers.get("Content-Length", 0))
                try:
                    hints = json.loads(self.rfile.read(length) or b"{}")
                    outer.hints = {k: hints[k] for k in SCHED_HINTS
                                   if k in hints}
                    self._reply(200, "")
                    from adaptdl_b200.sched import metrics
                    metrics.observe_hints(
                        "local", self.path.rsplit("/", 1)[-1], outer.hints)
                except ValueError:
                    self._reply(400, "")

            def _reply(self, code, body, ctype="application/json"):
                data = body.encode()
                self.send_response(code)
                self.send_header("Content-Type", ctype)
                self.send_header("Content-Length", str(len(data)))
                self.end_headers()
                self.wfile.write(data)
        self._server = ThreadingHTTPServer(("127.0.0.1", 0), Handler)
        self.url = "http://127.0.0.1:{}".format(self._server.server_port)
        threading.Thread(target=self._server.serve_forever,
                         daemon=True).start()

    def close(self):
        self._server.shutdown()


def best_replicas(hints, max_replicas, current, hysteresis=0.05):
    """Replica count in ``[1, max_replicas]`` with the best goodput speedup
    according to the job's hints; keeps ``current`` unless the best is more
    than ``hysteresis`` better. Exploration is capped at twice the largest
    profiled count (like the cluster allocator)."""
    if not hints or not hints.get("perfParams") or \
            not hints.get("initBatchSize"):
        return current
    perf = PerfParams(*[hints["perfParams"][k] for k in PERF_PARAMS])
    grad = hints.get("gradParams")
    grad = GradParams(grad["norm"], grad["var"]) if grad \
        else GradParams(0.0, 1.0)
    bounds = hints.get("localBszBounds")
    fn = SpeedupFunction(
        GoodputFunction(perf, grad, hints["initBatchSize"]),
        hints.get("maxBatchSize"), tuple(bounds) if bounds else None,
        hints.get("gradientAccumulation", False))
    cap = min(max_replicas, max(2 * hints.get("maxProfiledReplicas", 1), 1))
    speedups = {n: fn(1, n) for n in range(1, cap + 1)}
    best = max(speedups, key=lambda n: (speedups[n], -n))
    if current in speedups and \
            speedups[best] <= speedups[current] * (1 + hysteresis):
        return current
    return best


SCRATCH_MIN_FREE = 8 << 30        # bytes /dev/shm must have free to be used


def scratch_checkpoint_dir(prefix):
    """Directory for checkpoints that only have to live as long as the
    launcher (a rescale hands them from one generation to the next): memory
    backed when the machine has a writable ``/dev/shm``, so that a large
    model is written and read at memory speed and never waits for a disk
    (``ADAPTDL_B200_SCRATCH`` names another place). Pass ``--checkpoint-dir``
    for checkpoints that must survive the launcher."""
    base = os.environ.get("ADAPTDL_B200_SCRATCH")
    if base is None and os.path.isdir("/dev/shm") and \
            os.access("/dev/shm", os.W_OK | os.X_OK):
        try:       # containers often get a 64 MB /dev/shm: not for checkpoints
            roomy = shutil.disk_usage("/dev/shm").free >= SCRATCH_MIN_FREE
        except OSError:
            roomy = False
        if roomy:
            base = "/dev/shm"
    path = tempfile.mkdtemp(prefix=prefix, dir=base or None)
    # scratch means scratch: gone when the launcher process ends (memory-
    # backed files would otherwise hold RAM until the next reboot)
    atexit.register(shutil.rmtree, path, ignore_errors=True)
    return path


class LocalElasticJob(object):
    """One elastic job on this machine's GPUs."""

    def __init__(self, command, max_replicas, checkpoint_dir=None,
                 job_id="local/job", env=None, gpu_ids=None, standby=False,
                 preload=()):
        """``standby``: keep ``max_replicas`` warm interpreters
        (:mod:`adaptdl_b200.sched.standby`) so that a new generation skips
        the ~6.5 s of ``import torch`` / ``torch.optim``; needs a command of
        the form ``python script.py ...``. ``preload``: extra modules the
        standbys import ahead of time (e.g. ``torchvision``)."""
        self.command = list(command)
        self.preload = tuple(preload)
        self.standby = bool(standby) and self._standby_command() is not None
        self.pool = []               # idle warm interpreters
        self._pool_due = None        # when to top the pool up again
        self.max_replicas = max_replicas
        self.checkpoint_dir = checkpoint_dir or scratch_checkpoint_dir(
            "adaptdl-b200-ckpt-")
        self.job_id = job_id
        self.extra_env = dict(env or {})
        self.gpu_ids = list(gpu_ids) if gpu_ids is not None else None
        self.num_restarts = 0
        self.replicas = 0
        self.procs = []
        self.server = HintsServer()
        self.events = []             # (time, what, detail) for reports

    def _log(self, what, **detail):
        self.events.append((time.time(), what, detail))
        LOG.info("%s %s", what, detail)

    _STALE = ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR",
              "MASTER_PORT")

    def _job_env(self):
        """Variables every process of this job gets, whatever its
        generation."""
        return {
            "ADAPTDL_JOB_ID": self.job_id,
            "ADAPTDL_CHECKPOINT_PATH": self.checkpoint_dir,
            "ADAPTDL_MASTER_ADDR": "127.0.0.1",
            "ADAPTDL_NUM_NODES": "1",
            "ADAPTDL_SUPERVISOR_URL": "",
            "ADAPTDL_HINTS_URL": self.server.url,
        }

    def _replica_env(self, replicas, rank, port):
        """Variables of one replica of one generation."""
        return {
            "ADAPTDL_MASTER_PORT": str(port),
            "ADAPTDL_NUM_REPLICAS": str(replicas),
            "ADAPTDL_REPLICA_RANK": str(rank),
            "ADAPTDL_NUM_RESTARTS": str(self.num_restarts),
            "ADAPTDL_LOCAL_RANK": str(
                self.gpu_ids[rank] if self.gpu_ids else rank),
        }

    def _base_env(self):
        env = dict(os.environ)
        env.update(self.extra_env)
        env.update(self._job_env())
        for stale in self._STALE:
            env.pop(stale, None)
        return env

    def _standby_command(self):
        """``python -m adaptdl_b200.sched.standby script.py ...`` for a
        ``python script.py ...`` command, else ``None``."""
        cmd = self.command
        if len(cmd) >= 2 and cmd[1].endswith(".py") and \
                os.path.basename(cmd[0]).startswith("python"):
            head = [cmd[0], "-m", "adaptdl_b200.sched.standby"]
            if self.preload:
                head += ["--preload", ",".join(self.preload)]
            return head + ["--"] + cmd[1:]
        return None

    def fill_pool(self):
        """Top the pool of warm interpreters up to ``max_replicas``."""
        self._pool_due = None
        if not self.standby:
            return
        self.pool = [p for p in self.pool if p.poll() is None]
        env = self._base_env()
        # the repository root must be importable for ``-m``
        root = os.path.dirname(os.path.dirname(os.path.dirname(
            os.path.abspath(__file__))))
        env["PYTHONPATH"] = os.pathsep.join(
            [root] + [p for p in env.get("PYTHONPATH", "").split(os.pathsep)
                      if p])
        while len(self.pool) < self.max_replicas:
            self.pool.append(subprocess.Popen(
                self._standby_command(), env=env, stdin=subprocess.PIPE))

    def maintain(self):
        """Periodic housekeeping (called from :meth:`run`'s loop)."""
        if self._pool_due is not None and time.time() >= self._pool_due:
            self.fill_pool()

    def _release(self, proc, env):
        """Turn a warm interpreter into a replica."""
        order = {"env": env, "unset": list(self._STALE)}
        try:
            proc.stdin.write((json.dumps(order) + "\n").encode())
            proc.stdin.close()
            return True
        except (OSError, ValueError):
            return False

    def drain_pool(self):
        for p in self.pool:
            try:
                p.stdin.close()          # EOF = "not needed": exits 0
            except (OSError, ValueError):
                pass
        for p in self.pool:
            try:
                p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                p.kill()
                p.wait()
        self.pool = []

    def start(self, replicas, gpu_ids=None):
        """Start a generation with ``replicas`` processes; ``gpu_ids``
        (optional) are the device indices of this generation's replicas."""
        assert not self.procs
        if gpu_ids is not None:
            assert len(gpu_ids) >= replicas
            self.gpu_ids = list(gpu_ids)
        port = pick_unused_port()
        self.replicas = self.server.replicas = replicas
        self.pool = [p for p in self.pool if p.poll() is None]
        warm = 0
        for rank in range(replicas):
            mine = self._replica_env(replicas, rank, port)
            proc = None
            while self.pool and proc is None:
                candidate = self.pool.pop(0)
                if self._release(candidate, mine):
                    proc, warm = candidate, warm + 1
            if proc is None:
                env = self._base_env()
                env.update(mine)
                proc = subprocess.Popen(self.command, env=env)
            self.procs.append(proc)
        if self.standby:
            # refill once the new generation is through its own start-up
            # (the imports of 8 standbys would compete with it for the CPU)
            self._pool_due = time.time() + 5.0
        self._log("started", replicas=replicas, generation=self.num_restarts,
                  warm=warm)

    def poll(self):
        """``None`` while running; else ``"finished"`` / ``"preempted"`` /
        ``"failed"`` once every replica has exited."""
        # a replica that receives SIGTERM before its interpreter has installed
        # the handler (still importing torch) dies of the signal: it had
        # nothing to checkpoint yet, which is a preemption, not a failure
        # (Kubernetes reports the same case as exit code 143)
        codes = [EXIT_PREEMPTED if c == -signal.SIGTERM else c
                 for c in (p.poll() for p in self.procs)]
        if any(c is None for c in codes):
            if any(c not in (None, 0, EXIT_PREEMPTED) for c in codes):
                self._log("replica_failed", exit_codes=codes)
                self.kill()
                return "failed"
            return None
        self.procs = []
        if all(c == 0 for c in codes):
            return "finished"
        if all(c in (0, EXIT_PREEMPTED) for c in codes):
            return "preempted"
        self._log("replica_failed", exit_codes=codes)
        return "failed"

    def signal_stop(self):
        for p in self.procs:
            if p.poll() is None:
                p.send_signal(signal.SIGTERM)

    def kill(self):
        for p in self.procs:
            if p.poll() is None:
                p.kill()
        for p in self.procs:
            p.wait()
        self.procs = []
        self.drain_pool()

    def rescale(self, replicas, timeout=300.0, gpu_ids=None):
        """SIGTERM -> wait for the checkpoint exit -> start the next
        generation (``replicas == 0``: stay stopped). Returns the state
        (``"running"`` / ``"stopped"`` / ``"finished"`` / ``"failed"``)."""
        t0 = time.time()
        self.signal_stop()
        deadline = t0 + timeout
        state = None
        while state is None and time.time() < deadline:
            state = self.poll()
            if state is None:
                time.sleep(0.05)
        if state is None:
            self.kill()
            state = "failed"
        self._log("stopped", state=state, seconds=time.time() - t0)
        if state != "preempted":
            return state
        self.num_restarts += 1
        if replicas == 0:
            self.replicas = 0
            return "stopped"
        self.start(replicas, gpu_ids)
        self._log("rescaled", replicas=replicas,
                  seconds=time.time() - t0)
        return "running"

    def run(self, schedule=None, interval=30.0, adaptive=False,
            initial=None, stop_after=None):
        """Run to completion. ``schedule``: replica counts applied every
        ``interval`` seconds (the last one stays); ``adaptive``: follow the
        job's hints instead."""
        schedule = list(schedule or [])
        first = initial or (schedule.pop(0) if schedule else 1)
        self.start(min(first, self.max_replicas))
        next_change = time.time() + interval
        deadline = time.time() + stop_after if stop_after else None
        try:
            while True:
                if deadline is not None and time.time() >= deadline:
                    # preempt for good: checkpoint and leave
                    t0 = time.time()
                    self.signal_stop()
                    state = None
                    while state is None and time.time() - t0 < 300:
                        state = self.poll()
                        time.sleep(0.05)
                    self._log("stopped", state=state,
                              seconds=time.time() - t0)
                    return "stopped"
                self.maintain()
                state = self.poll()
                if state in ("finished", "failed"):
                    self._log(state)
                    return state
                if state == "preempted":          # exited on its own signal
                    self.num_restarts += 1
                    self.start(self.replicas)
                if time.time() >= next_change:
                    next_change = time.time() + interval
                    target = self.replicas
                    if schedule:
                        target = min(schedule.pop(0), self.max_replicas)
                    elif adaptive:
                        target = best_replicas(self.server.hints,
                                               self.max_replicas,
                                               self.replicas)
                    if target != self.replicas:
                        state = self.rescale(target)
                        if state != "running":
                            self._log(state)
                            return state
                time.sleep(0.1)
        finally:
            self.kill()
            self.server.close()


def main(argv=None):
    parser = argparse.ArgumentParser(
        description="run an elastic adaptdl_b200 job on this box")
    parser.add_argument("--gpus", type=int, default=None,
                        help="GPUs (max replicas); default: all visible")
    parser.add_argument("--schedule", default="",
                        help="comma-separated replica counts, e.g. 2,4,8,4")
    parser.add_argument("--interval", type=float, default=30.0)
    parser.add_argument("--adaptive", action="store_true",
                        help="choose the replica count from the job's hints")
    parser.add_argument("--checkpoint-dir", default=None)
    parser.add_argument("--stop-after", type=float, default=None,
                        help="preempt the job for good after this many "
                             "seconds (checkpoint + exit)")
    parser.add_argument("--report", default=None,
                        help="write the event log (JSON) here")
    parser.add_argument("--trace-rescale", action="store_true",
                        help="have every replica record its life-cycle "
                             "events (utils/rescale_trace.py) and add the "
                             "per-generation phase breakdown to the report")
    parser.add_argument("--standby", action="store_true",
                        help="keep a pool of warm interpreters (torch "
                             "already imported) for the next generation")
    parser.add_argument("--fast-exit", action="store_true",
                        help="replicas leave with os._exit once their "
                             "checkpoint is written (atexit handlers still "
                             "run; interpreter/torch teardown is skipped)")
    parser.add_argument("--preload", default="",
                        help="with --standby: extra modules to import ahead "
                             "of time, comma-separated")
    parser.add_argument("script", nargs=argparse.REMAINDER)
    args = parser.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    if not args.script:
        parser.error("missing training script")
    gpus = args.gpus
    if gpus is None:
        try:
            import torch
            gpus = max(torch.cuda.device_count(), 1)
        except Exception:  # noqa: BLE001
            gpus = 1
    command = args.script
    if command[0].endswith(".py"):
        command = [sys.executable] + command
    extra_env = {}
    trace_dir = None
    if args.trace_rescale:
        trace_dir = tempfile.mkdtemp(prefix="adaptdl-b200-rescale-trace-")
        extra_env["ADAPTDL_B200_RESCALE_TRACE"] = trace_dir
    if args.fast_exit:
        extra_env["ADAPTDL_B200_FAST_EXIT"] = "1"
    job = LocalElasticJob(command, gpus, args.checkpoint_dir, env=extra_env,
                          standby=args.standby,
                          preload=[m for m in args.preload.split(",") if m])
    schedule = [int(x) for x in args.schedule.split(",") if x]
    state = job.run(schedule, args.interval, args.adaptive,
                    stop_after=args.stop_after)
    if args.report:
        rows = [{"t": t, "event": what, **detail}
                for t, what, detail in job.events]
        if trace_dir:
            from adaptdl_b200.utils import rescale_trace
            traced = rescale_trace.collect(trace_dir)
            rows.append({"event": "rescale_phases_seconds",
                         "by_generation": rescale_trace.summarize(traced)})
            # launcher-side view: SIGTERM sent -> every replica exited ->
            # next generation's first optimizer step
            first_step = {}
            for row in traced:
                if row["event"] == "first_step_done":
                    gen = row["generation"]
                    first_step[gen] = max(first_step.get(gen, 0), row["t"])
            rows.append({"event": "first_step_done_at",
                         "by_generation": first_step})
        with open(args.report, "w") as f:
            json.dump(rows, f, indent=1)
    return 0 if state in ("finished", "stopped") else 1


if __name__ == "__main__":
    sys.exit(main())
