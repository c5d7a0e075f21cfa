"""Elastic multi-process test harness.

``@elastic_multiprocessing`` runs the decorated function as a tiny elastic
job on localhost: it is started with ONE replica in a forked process with a
full ``ADAPTDL_*`` environment and a shared temporary checkpoint directory;
the function's **return value is the replica count of the next restart**
(``0``/``None`` ends the job). Every replica must return the same value and
exit cleanly. This is how multi-replica and rescale behaviour is tested
without a cluster (same idea as the reference's
``adaptdl/adaptdl/conftest.py:25-100``).
"""

import functools
import multiprocessing as mp
import os
import signal
import tempfile
import traceback

from adaptdl_b200.utils import pick_unused_port


def reset_global_state():
    """Forget every process-global of the framework (state registry, epoch /
    metrics / loader singletons, control plane). Forked test replicas call
    this so they do not inherit whatever the parent test process touched."""
    import sys
    from adaptdl_b200 import checkpoint, collective, _signal
    checkpoint._reset_registry_for_tests()
    if collective.is_initialized():
        collective._REDUCER = None      # parent's sockets are not ours
    _signal.set_exit_flag(False)
    for name in ("adaptdl_b200.torch.epoch", "adaptdl_b200.torch._metrics",
                 "adaptdl_b200.torch.data", "adaptdl_b200.torch.accumulator"):
        mod = sys.modules.get(name)
        if mod is not None:
            mod._reset_for_tests()


def _child(func, args, kwargs, environ, rank, queue):
    os.environ.update(environ)
    os.environ["ADAPTDL_REPLICA_RANK"] = str(rank)
    reset_global_state()
    ret, err = None, None
    try:
        ret = func(*args, **kwargs)
    except SystemExit as exc:      # exit(143) after a checkpoint
        err = ("exit", exc.code)
    except BaseException:          # noqa: BLE001
        err = ("error", traceback.format_exc())
    finally:
        queue.put((rank, ret, err))
        queue.close()
        queue.join_thread()
    os._exit(0)


def elastic_multiprocessing(func):
    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        ctx = mp.get_context("fork")
        num_restarts, num_replicas = 0, 1
        with tempfile.TemporaryDirectory() as tmpdir:
            while num_replicas:
                assert isinstance(num_replicas, int)
                environ = {
                    "ADAPTDL_CHECKPOINT_PATH": str(tmpdir),
                    "ADAPTDL_JOB_ID": "tmpjob",
                    "ADAPTDL_MASTER_ADDR": "127.0.0.1",
                    "ADAPTDL_MASTER_PORT": str(pick_unused_port()),
                    "ADAPTDL_NUM_REPLICAS": str(num_replicas),
                    "ADAPTDL_NUM_NODES": "1",
                    "ADAPTDL_NUM_RESTARTS": str(num_restarts),
                }
                queue = ctx.Queue()
                procs = [ctx.Process(target=_child,
                                     args=(func, args, kwargs, environ, rank,
                                           queue))
                         for rank in range(num_replicas)]
                for proc in procs:
                    proc.start()
                try:
                    results = {}
                    for _ in range(num_replicas):
                        rank, ret, err = queue.get(timeout=300)
                        if err is not None and err[0] == "error":
                            raise AssertionError(
                                "replica {} failed:\n{}".format(rank, err[1]))
                        assert err is None, \
                            "replica {} exited with {}".format(rank, err[1])
                        results[rank] = ret
                    for proc in procs:
                        proc.join(timeout=60)
                        assert proc.exitcode == 0
                    values = set(results.values())
                    assert len(values) == 1, \
                        "replicas disagree on next size: {}".format(results)
                    num_replicas = values.pop()
                finally:
                    for proc in procs:
                        if proc.is_alive():
                            os.kill(proc.pid, signal.SIGKILL)
                        proc.join()
                    queue.close()
                num_restarts += 1
    return wrapper
