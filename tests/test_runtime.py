"""L2 runtime: env, checkpoint registry, control-plane collectives."""
import os
import pickle

import pytest

from adaptdl_b200.utils.testing import elastic_multiprocessing


def test_env_defaults(monkeypatch):
    from adaptdl_b200 import env
    for key in list(os.environ):
        if key.startswith("ADAPTDL_"):
            monkeypatch.delenv(key)
    assert env.checkpoint_path() is None
    assert env.share_path() is None
    assert env.job_id() is None
    assert env.master_addr() == "0.0.0.0"
    assert env.master_port() == 0
    assert env.replica_rank() == 0
    assert env.num_replicas() == 1
    assert env.num_nodes() == 1
    assert env.num_restarts() == 0
    assert env.supervisor_url() is None
    assert env.from_ray() is False
    monkeypatch.setenv("ADAPTDL_NUM_REPLICAS", "4")
    assert env.num_nodes() == 4          # defaults to one node per replica
    monkeypatch.setenv("ADAPTDL_NUM_NODES", "2")
    monkeypatch.setenv("ADAPTDL_REPLICA_RANK", "3")
    assert env.num_nodes() == 2 and env.replica_rank() == 3
    assert env.local_rank() == 1


@elastic_multiprocessing
def test_duplicate_state():
    from adaptdl_b200.env import num_restarts
    from adaptdl_b200.checkpoint import State
    State("state_1")
    State("state_2")
    with pytest.raises(ValueError):
        State("state_1")
    return [2, 0][num_restarts()]


@elastic_multiprocessing
def test_save_load():
    from adaptdl_b200.checkpoint import State, save_all_states, load_state
    from adaptdl_b200.env import replica_rank, num_restarts, checkpoint_path

    class TestState(State):
        def __init__(self, name):
            super().__init__(name)
            self.synced = False

        def sync(self):
            self.synced = True

        def save(self, fileobj):
            assert replica_rank() == 0
            pickle.dump(self.value, fileobj)

        def load(self, fileobj):
            self.value = pickle.load(fileobj)

    state_1 = TestState("state_1")
    state_2 = TestState("state_2")
    if num_restarts() == 0:
        assert not load_state(state_1)
        state_1.value, state_2.value = 10, 20
        save_all_states()
        assert state_1.synced and state_2.synced
        assert sorted(os.listdir(checkpoint_path())) == ["checkpoint-0"]
        assert sorted(os.listdir(os.path.join(
            checkpoint_path(), "checkpoint-0"))) == ["state_1", "state_2"]
        return 2
    assert load_state(state_1) and load_state(state_2)
    if num_restarts() == 1:
        assert (state_1.value, state_2.value) == (10, 20)
        # a newer generation replaces the older one atomically
        from adaptdl_b200 import collective
        collective.initialize()
        collective.allreduce(0)   # barrier: everyone finished loading
        state_1.value = 11
        save_all_states()
        collective.allreduce(0)   # barrier: rank 0 finished publishing
        assert sorted(d for d in os.listdir(checkpoint_path())
                      if d.startswith("checkpoint-")) == ["checkpoint-1"]
        return 1
    assert state_1.value == 11
    return 0


@elastic_multiprocessing
def test_allreduce():
    from adaptdl_b200 import collective, env
    collective.initialize()
    assert collective.allreduce(1) == env.num_replicas()
    result = collective.allreduce({env.replica_rank()},
                                  reduce_fn=lambda a, b: a | b)
    assert result == set(range(env.num_replicas()))
    return [5, 0][env.num_restarts()]


@elastic_multiprocessing
def test_allreduce_async_out_of_order():
    from adaptdl_b200 import collective, env
    collective.initialize()
    f1 = collective.allreduce_async(1)
    f2 = collective.allreduce_async(2)
    f3 = collective.allreduce_async(3)
    n = env.num_replicas()
    assert f2.result() == 2 * n
    assert f1.result() == 1 * n
    assert f3.result() == 3 * n
    assert f2.result() == 2 * n        # idempotent
    return [5, 0][env.num_restarts()]


@elastic_multiprocessing
def test_broadcast_and_teardown():
    from adaptdl_b200 import collective, env
    with pytest.raises(RuntimeError):
        collective.broadcast(1)
    collective.initialize()
    with pytest.raises(RuntimeError):
        collective.initialize()
    assert collective.broadcast(env.replica_rank()) == 0
    big = collective.broadcast(list(range(200000)))
    assert len(big) == 200000
    collective.allreduce(0)
    collective.teardown()
    assert not collective.is_initialized()
    return [3, 0][env.num_restarts()]


def test_reducer_many_steps():
    """3 raw processes, Counter reduce, async loop (reference
    reducer_test)."""
    import collections
    import multiprocessing as mp
    from adaptdl_b200.utils import pick_unused_port
    port = pick_unused_port()

    def main(rank, size):
        from adaptdl_b200.reducer import Reducer
        reducer = Reducer(rank, size, "127.0.0.1", port)
        if rank == 0:
            batch_size = 28
            x = {"foo": 1}
        else:
            x = {"bar": 1}
            batch_size = 0
        assert reducer.broadcast(batch_size) == 28
        total = reducer.allreduce(collections.Counter(x))
        assert total["foo"] == 1 and total["bar"] == size - 1
        futures = [reducer.allreduce_async(i) for i in range(10)]
        for i, fut in reversed(list(enumerate(futures))):
            assert fut.result() == i * size
        reducer.close()

    ctx = mp.get_context("fork")
    procs = [ctx.Process(target=main, args=(r, 3)) for r in range(3)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0


def test_signal_flag():
    import signal
    from adaptdl_b200 import _signal
    _signal.install()
    assert not _signal.get_exit_flag()
    os.kill(os.getpid(), signal.SIGTERM)
    assert _signal.get_exit_flag()
    _signal.set_exit_flag(False)


def test_fd_exchange_between_processes():
    """The SCM_RIGHTS all-to-all used to share GPU memory handles, exercised
    with ordinary file descriptors: every rank sends an unlinked file's
    descriptor to every peer for two rounds; peers read the token through it."""
    import multiprocessing as mp
    import tempfile
    world = 3
    tmp = tempfile.mkdtemp()

    def main(rank):
        import time
        from adaptdl_b200.parallel.symm import FdExchange

        def gather(addr):
            with open(os.path.join(tmp, "addr{}".format(rank)), "wb") as f:
                f.write(addr.encode("utf-8", "surrogateescape"))
            out = []
            for r in range(world):
                path = os.path.join(tmp, "addr{}".format(r))
                while not os.path.exists(path) or \
                        os.path.getsize(path) == 0:
                    time.sleep(0.01)
                with open(path, "rb") as f:
                    out.append(f.read().decode("utf-8", "surrogateescape"))
            time.sleep(0.2)
            return out
        ex = FdExchange(rank, world, gather)
        for rnd in range(2):
            path = os.path.join(tmp, "tok{}-{}".format(rank, rnd))
            with open(path, "wb") as f:
                f.write("r{}-{}".format(rank, rnd).encode())
            mine = os.open(path, os.O_RDONLY)
            os.unlink(path)          # only reachable through the descriptor
            got = ex.exchange(mine)
            assert sorted(got) == [p for p in range(world) if p != rank]
            for src, fd in got.items():
                assert os.pread(fd, 64, 0) == \
                    "r{}-{}".format(src, rnd).encode()
                os.close(fd)
            os.close(mine)
        ex.close()

    ctx = mp.get_context("fork")
    procs = [ctx.Process(target=main, args=(r,)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0


def test_modules_import_standalone():
    """Each sub-package can be the FIRST thing a program imports (no import
    cycles hidden by importing ``adaptdl_b200.torch`` first)."""
    import subprocess
    import sys
    mods = ["adaptdl_b200.parallel.graph", "adaptdl_b200.parallel.engine",
            "adaptdl_b200.ops", "adaptdl_b200.models",
            "adaptdl_b200.torch.parallel", "adaptdl_b200.sched.local"]
    procs = [subprocess.Popen([sys.executable, "-c", "import " + m],
                              stderr=subprocess.PIPE) for m in mods]
    for m, proc in zip(mods, procs):
        _, err = proc.communicate(timeout=300)
        assert proc.returncode == 0, (m, err.decode()[-800:])


def test_print_exc_shows_tracebacks_of_hook_exceptions(capsys):
    from adaptdl_b200.utils import print_exc

    @print_exc
    def hook(x):
        raise KeyError("lost in autograd")

    with pytest.raises(KeyError):
        hook(1)
    assert "lost in autograd" in capsys.readouterr().err
    assert hook.__name__ == "hook"


def _fail_fast_worker(rank, port, out):
    from adaptdl_b200.reducer import Reducer
    reducer = Reducer(rank, 3, "127.0.0.1", port)
    assert reducer.allreduce(1) == 3                  # everybody is there
    if rank == 2:
        os._exit(7)                                   # dies without a word
    try:
        reducer.allreduce(1)                          # can never complete
        out.put((rank, "completed"))
    except ConnectionError as exc:
        out.put((rank, "ConnectionError"))
    except Exception as exc:  # noqa: BLE001
        out.put((rank, repr(exc)))


def test_a_replica_dying_mid_job_fails_the_others_fast():
    """The reference's reducer waits forever for a replica that is gone;
    here the survivors get a ConnectionError within moments."""
    import multiprocessing as mp
    import time
    from adaptdl_b200.utils import pick_unused_port
    ctx = mp.get_context("fork")
    out, port = ctx.Queue(), pick_unused_port()
    procs = [ctx.Process(target=_fail_fast_worker, args=(r, port, out))
             for r in range(3)]
    began = time.time()
    for proc in procs:
        proc.start()
    results = dict(out.get(timeout=60) for _ in range(2))
    assert results == {0: "ConnectionError", 1: "ConnectionError"}, results
    assert time.time() - began < 30
    for proc in procs:
        proc.join(10)


def test_resaving_a_generation_never_leaves_the_disk_without_a_checkpoint(
        tmp_path, monkeypatch):
    """Periodic saves inside one restart generation replace
    ``checkpoint-N``; a crash between the two renames must leave the previous
    copy loadable (the reference deletes first and renames second)."""
    from adaptdl_b200 import checkpoint
    checkpoint._reset_registry_for_tests()
    monkeypatch.setenv("ADAPTDL_CHECKPOINT_PATH", str(tmp_path))
    monkeypatch.setenv("ADAPTDL_NUM_RESTARTS", "0")

    class Counter(checkpoint.State):
        value = 0

        def save(self, f):
            pickle.dump(self.value, f)

        def load(self, f):
            self.value = pickle.load(f)
    state = Counter("counter")
    state.value = 1
    checkpoint.save_all_states()
    state.value = 2
    real_rename, calls = os.rename, []

    def crashing_rename(src, dst):
        calls.append((src, dst))
        if len(calls) == 2:                      # staging -> final
            raise OSError("power cut")
        return real_rename(src, dst)
    monkeypatch.setattr(os, "rename", crashing_rename)
    with pytest.raises(OSError):
        checkpoint.save_all_states()
    monkeypatch.setattr(os, "rename", real_rename)
    # "restart": the previous copy of generation 0 is still found and loads
    monkeypatch.setenv("ADAPTDL_NUM_RESTARTS", "1")
    state.value = None
    assert checkpoint.load_state(state) and state.value == 1
    # and a later successful save cleans everything up
    state.value = 3
    checkpoint.save_all_states()
    assert sorted(os.listdir(tmp_path)) == ["checkpoint-1"]
    checkpoint._reset_registry_for_tests()


def test_reducer_many_large_async_allreduces_do_not_deadlock():
    """Several outstanding multi-megabyte all-reduces per replica (a
    blocking server wedges here: it is stuck sending result k to a client
    that is itself stuck sending frame k+1)."""
    import threading
    import numpy as np
    from adaptdl_b200.reducer import Reducer
    from adaptdl_b200.utils.testing import pick_unused_port
    port = pick_unused_port()
    replicas, rounds, n = 2, 3, 3 * 1024 * 1024      # ~24 MB float64 each
    out = {}

    def worker(rank):
        red = Reducer(rank, replicas, "127.0.0.1", port)
        futures = [red.allreduce_async(np.full(n, float(rank + 1 + k)))
                   for k in range(rounds)]
        out[rank] = [float(f.result()[0]) for f in futures]
        red.broadcast(None)
        red.close()
    threads = [threading.Thread(target=worker, args=(r,), daemon=True)
               for r in range(replicas)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not any(t.is_alive() for t in threads), "control plane deadlocked"
    want = [float(sum(r + 1 + k for r in range(replicas)))
            for k in range(rounds)]
    assert out[0] == want and out[1] == want
