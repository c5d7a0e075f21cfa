"""Elastic worker group: the replicas of ONE Tune trial.

A trial's training function runs in ``len(allocation)`` worker processes (one
per replica, one per GPU) that form an ordinary adaptdl_b200 job among
themselves: ``ADAPTDL_*`` variables, the TCP control plane, a
``torch.distributed`` group. The trial object in the Tune driver only sees

* ``next_result()`` -- what rank 0 passed to ``report(...)``, with the job's
  scheduling hints attached so the trial scheduler can build a speedup model;
* ``checkpoint()`` -- asks the replicas to stop the AdaptDL way (the exit flag
  is OR-reduced by the data loader, every replica saves its registered
  ``State`` objects at the same iteration and exits with code 143) and returns
  rank 0's checkpoint directory as ``{relative path: bytes}``;
* a new ``WorkerGroup(..., checkpoint=that_object)`` on ANOTHER allocation
  resumes from it -- that is a rescale.

Where the workers run is a *spawner*: :class:`ProcessSpawner` (local
processes; used by the tests and by single-node runs without Ray) or
:class:`RayActorSpawner` (one actor per placement-group bundle). Reference
counterparts: ``ray/adaptdl_ray/tune/adaptdl_trainable.py:29-81`` and
``adaptdl_patch.py:29-63`` (workers exposing ``save_all_states`` /
``get_sched_hints``), ``adaptdl_trial.py:70-97`` (clone + restore in memory).
"""

import logging
import multiprocessing
import os
import queue
import shutil
import tempfile
import threading
import time
import traceback

from adaptdl_b200.ray.aws.utils import (checkpoint_obj_to_dir,
                                        serialize_checkpoint)
from adaptdl_b200.utils import pick_unused_port

LOG = logging.getLogger(__name__)

EXIT_PREEMPTED = 143
HINTS_PERIOD_S = 10.0            # re-fitting the model is not free

RESULT, DONE, PREEMPTED, FAILED = "result", "done", "preempted", "failed"


# ---------------------------------------------------------------------------
# inside a worker
# ---------------------------------------------------------------------------

class _Reporter(object):
    """``report(**metrics)`` handed to the training function."""

    def __init__(self, rank, results):
        self._rank = rank
        self._results = results
        self._hints = None
        self._hints_time = 0.0

    def _sched_hints(self):
        now = time.time()
        if now - self._hints_time >= HINTS_PERIOD_S:
            from adaptdl_b200.torch import _metrics
            try:
                if _metrics._get_sched_hints() is not None:
                    self._hints = _metrics._build_sched_hints()
            except Exception:  # noqa: BLE001 - telemetry only
                LOG.debug("no scheduling hints yet", exc_info=True)
            self._hints_time = now
        return self._hints

    def __call__(self, **metrics):
        if self._rank == 0:
            metrics["sched_hints"] = self._sched_hints()
            self._results.put((RESULT, metrics))


def run_worker(train_fn, config, rank, env_vars, checkpoint, results,
               wait_for_preempt=None):
    """Body of one replica. ``results``: anything with ``put``;
    ``wait_for_preempt``: blocking callable that returns when the driver
    wants the group to stop (``None``: the spawner sets the exit flag
    itself)."""
    os.environ.update(env_vars)
    ckpt_dir = tempfile.mkdtemp(prefix="adaptdl-tune-{}-".format(rank))
    os.environ["ADAPTDL_CHECKPOINT_PATH"] = ckpt_dir
    if checkpoint:
        checkpoint_obj_to_dir(ckpt_dir, checkpoint)
    from adaptdl_b200 import _signal
    _signal.set_exit_flag(False)
    if wait_for_preempt is not None:
        def watch():
            wait_for_preempt()
            _signal.set_exit_flag(True)
        threading.Thread(target=watch, daemon=True).start()
    status, payload = FAILED, None
    try:
        train_fn(config, _Reporter(rank, results))
        status = DONE
    except SystemExit as stop:
        status = PREEMPTED if stop.code == EXIT_PREEMPTED else FAILED
        payload = "exit code {}".format(stop.code)
    except BaseException:  # noqa: BLE001 - shipped to the driver
        payload = traceback.format_exc()
    finally:
        if status != FAILED:
            payload = serialize_checkpoint(ckpt_dir) if rank == 0 else None
        shutil.rmtree(ckpt_dir, ignore_errors=True)
        if rank == 0 or status == FAILED:
            results.put((status, payload))
    return status


def replica_env(rank, allocation, master_addr, master_port, generation,
                job_id):
    return {
        "ADAPTDL_JOB_ID": str(job_id),
        "ADAPTDL_MASTER_ADDR": master_addr,
        "ADAPTDL_MASTER_PORT": str(master_port),
        "ADAPTDL_REPLICA_RANK": str(rank),
        "ADAPTDL_NUM_REPLICAS": str(len(allocation)),
        "ADAPTDL_NUM_NODES": str(len(set(allocation))),
        "ADAPTDL_NUM_RESTARTS": str(generation),
        # replicas on one node take consecutive devices
        "ADAPTDL_LOCAL_RANK": str(list(allocation[:rank]).count(
            allocation[rank])),
    }


# ---------------------------------------------------------------------------
# spawners
# ---------------------------------------------------------------------------

def _process_entry(train_fn, config, rank, env_vars, checkpoint, results,
                   preempt):
    run_worker(train_fn, config, rank, env_vars, checkpoint, results,
               preempt.wait)


class ProcessSpawner(object):
    """Replicas are local processes (``spawn`` start method: the training
    function must be importable, i.e. defined at module level)."""

    def __init__(self, extra_env=None):
        self._ctx = multiprocessing.get_context("spawn")
        self._extra_env = dict(extra_env or {})

    def start(self, train_fn, config, allocation, checkpoint, generation,
              job_id):
        results = self._ctx.Queue()
        preempt = self._ctx.Event()
        port = pick_unused_port()
        procs = []
        for rank in range(len(allocation)):
            env_vars = replica_env(rank, allocation, "127.0.0.1", port,
                                   generation, job_id)
            env_vars.update(self._extra_env)
            proc = self._ctx.Process(
                target=_process_entry, daemon=True,
                args=(train_fn, config, rank, env_vars, checkpoint, results,
                      preempt))
            proc.start()
            procs.append(proc)
        return _ProcessHandle(procs, results, preempt)


class _ProcessHandle(object):

    def __init__(self, procs, results, preempt):
        self._procs, self.results, self._preempt = procs, results, preempt

    def preempt(self):
        self._preempt.set()

    def alive(self):
        return any(p.is_alive() for p in self._procs)

    def join(self, timeout):
        deadline = time.time() + timeout
        for proc in self._procs:
            proc.join(max(deadline - time.time(), 0.0))

    def kill(self):
        for proc in self._procs:
            if proc.is_alive():
                proc.terminate()
        for proc in self._procs:      # SIGTERM only sets the exit flag
            proc.join(2.0)
            if proc.is_alive():
                proc.kill()
                proc.join(5.0)


class RayActorSpawner(object):
    """Replicas are Ray actors, one per bundle of the trial's placement
    group (bundle 0 is the trial driver's own)."""

    def __init__(self, placement_group=None, resources_per_replica=None):
        self._pg = placement_group
        self._resources = resources_per_replica

    def start(self, train_fn, config, allocation, checkpoint, generation,
              job_id):
        from adaptdl_b200.ray import require_ray
        ray = require_ray()
        from ray.util.queue import Queue
        from ray.util.scheduling_strategies import \
            PlacementGroupSchedulingStrategy
        from adaptdl_b200.ray.config import default_device

        @ray.remote(max_concurrency=2)
        class Replica(object):
            def address(self):
                return ray.util.get_node_ip_address(), pick_unused_port("")

            def run(self, rank, env_vars, results):
                return run_worker(train_fn, config, rank, env_vars,
                                  checkpoint, results)

            def preempt(self):
                from adaptdl_b200 import _signal
                _signal.set_exit_flag(True)

        resources = dict(self._resources or {"CPU": 1, default_device(): 1})
        options = dict(num_cpus=resources.pop("CPU", 1),
                       num_gpus=resources.pop("GPU", 0), resources=resources)
        actors = []
        for rank in range(len(allocation)):
            if self._pg is not None:
                options["scheduling_strategy"] = \
                    PlacementGroupSchedulingStrategy(
                        placement_group=self._pg,
                        placement_group_bundle_index=rank + 1)
            actors.append(Replica.options(**options).remote())
        addr, port = ray.get(actors[0].address.remote())
        results = Queue()
        runs = [actor.run.remote(
            rank, replica_env(rank, allocation, addr, port, generation,
                              job_id), results)
            for rank, actor in enumerate(actors)]
        return _RayHandle(ray, actors, runs, results)


class _RayHandle(object):

    def __init__(self, ray, actors, runs, results):
        self._ray, self._actors, self._runs = ray, actors, runs
        self.results = results

    def preempt(self):
        for actor in self._actors:
            actor.preempt.remote()

    def alive(self):
        _, pending = self._ray.wait(self._runs, num_returns=len(self._runs),
                                    timeout=0)
        return bool(pending)

    def join(self, timeout):
        self._ray.wait(self._runs, num_returns=len(self._runs),
                       timeout=timeout)

    def kill(self):
        for actor in self._actors:
            self._ray.kill(actor)


# ---------------------------------------------------------------------------
# the group
# ---------------------------------------------------------------------------

class WorkerGroupError(RuntimeError):
    pass


class WorkerGroup(object):
    """The running replicas of one trial on one allocation.

    Arguments:
        train_fn: ``train_fn(config, report)``; an ordinary adaptdl_b200
            training script body (``init_process_group``,
            ``AdaptiveDataParallel``, ``AdaptiveDataLoader``,
            ``remaining_epochs_until``) that calls ``report(**metrics)``.
        allocation: one node name per replica.
        checkpoint: object returned by a previous group's
            :meth:`checkpoint` (``None``: fresh start).
        generation: restart counter (``ADAPTDL_NUM_RESTARTS``).
    """

    def __init__(self, train_fn, config, allocation, spawner=None,
                 checkpoint=None, generation=0, job_id="tune/trial"):
        if not allocation:
            raise ValueError("empty allocation")
        self.allocation = list(allocation)
        self.generation = generation
        self.finished = False
        self._final = None
        self._handle = (spawner or ProcessSpawner()).start(
            train_fn, config, self.allocation, checkpoint, generation,
            job_id)

    def _take(self, timeout):
        deadline = time.time() + timeout
        while True:
            try:
                return self._handle.results.get(
                    timeout=min(max(deadline - time.time(), 0.01), 1.0))
            except queue.Empty:
                pass
            except Exception as exc:  # ray.util.queue.Empty
                if type(exc).__name__ != "Empty":
                    raise
            if time.time() >= deadline:
                raise TimeoutError("no message from the replicas")
            if self._final is None and not self._handle.alive():
                # the replicas are gone: drain what they left, then give up
                try:
                    return self._handle.results.get(timeout=1.0)
                except Exception:  # noqa: BLE001
                    raise WorkerGroupError("the replicas exited without "
                                           "reporting") from None

    def _settle(self, kind, payload):
        if kind == FAILED:
            self._handle.kill()
            raise WorkerGroupError("a replica failed:\n{}".format(payload))
        self.finished = kind == DONE
        self._final = (kind, payload)

    def next_result(self, timeout=600.0):
        """The next ``report(...)`` of rank 0 as a dict, or ``None`` once
        the training function has returned."""
        if self._final is not None:
            return None
        kind, payload = self._take(timeout)
        if kind == RESULT:
            return payload
        self._settle(kind, payload)
        return None

    def checkpoint(self, timeout=600.0):
        """Stop the replicas at their next iteration boundary and return
        the checkpoint they wrote (results still in flight are dropped)."""
        if self._final is None:
            self._handle.preempt()
            deadline = time.time() + timeout
            while self._final is None:
                kind, payload = self._take(max(deadline - time.time(), 0.01))
                if kind != RESULT:
                    self._settle(kind, payload)
        self._handle.join(30.0)
        self._handle.kill()
        return self._final[1]

    def shutdown(self):
        self._handle.kill()
