"""Epoch loop, data loader/sampler, accumulator and metrics across
restarts at changing replica counts (ideas from the reference's
torch/*_test.py, run on the local elastic harness)."""
import collections
import math
import os

import pytest
import torch
from torch.utils.data import TensorDataset

from adaptdl_b200.utils.testing import elastic_multiprocessing
from adaptdl_b200.torch.data import (ElasticSampler, AdaptiveDataLoader,
                                     current_dataloader)


@pytest.mark.parametrize("num_replicas", [1, 3, 5])
@pytest.mark.parametrize("dataset_size", [9, 15, 25])
def test_sampler_epoch(num_replicas, dataset_size):
    _sampler_epoch(num_replicas, dataset_size)


def _sampler_epoch(num_replicas, dataset_size, epoch=0, index=0,
                   shuffle=True):
    dataset = TensorDataset(torch.rand(dataset_size))
    sampler = ElasticSampler(dataset, shuffle=shuffle)
    sampler.num_replicas = num_replicas
    sampler.set_epoch(epoch, index)
    per_rank = []
    counts = collections.Counter()
    for rank in range(num_replicas):
        sampler.rank = rank
        per_rank.append(list(sampler))
        expect = math.ceil((dataset_size - index % dataset_size)
                           / num_replicas)
        assert len(sampler) == expect == len(per_rank[rank])
        assert list(sampler) == per_rank[rank]          # deterministic
        counts.update(per_rank[rank])
    assert len(counts) >= dataset_size - index % dataset_size
    assert all(0 <= key < dataset_size for key in counts)
    assert max(counts.values()) - min(counts.values()) <= 1
    return per_rank


@pytest.mark.parametrize("num_replicas", [1, 3, 5])
@pytest.mark.parametrize("dataset_size", [9, 15, 25])
def test_sampler_shuffle(num_replicas, dataset_size):
    e0 = _sampler_epoch(num_replicas, dataset_size, epoch=0)
    e1 = _sampler_epoch(num_replicas, dataset_size, epoch=1)
    assert e0 != e1
    e0 = _sampler_epoch(num_replicas, dataset_size, 0, shuffle=False)
    e1 = _sampler_epoch(num_replicas, dataset_size, 1, shuffle=False)
    assert e0 == e1


@pytest.mark.parametrize("num_replicas", [1, 3, 5])
@pytest.mark.parametrize("dataset_size", [9, 15, 25])
def test_sampler_index(num_replicas, dataset_size):
    index = dataset_size // 2
    per_rank = _sampler_epoch(num_replicas, dataset_size, index=index,
                                  shuffle=False)
    samples = sum(per_rank, [])
    assert all(idx in samples for idx in range(index, dataset_size))
    per_rank = _sampler_epoch(num_replicas, dataset_size,
                                  index=2 * dataset_size, shuffle=False)
    assert set(sum(per_rank, [])) == set(range(dataset_size))


def test_sampler_second_pass_reshuffles():
    dataset = TensorDataset(torch.rand(20))
    sampler = ElasticSampler(dataset, shuffle=True)
    sampler.set_epoch(3, 0)
    first = list(sampler)
    sampler.set_epoch(3, 20)       # same epoch, second pass over the data
    assert list(sampler) != first


@elastic_multiprocessing
def test_epoch():
    from adaptdl_b200 import checkpoint
    from adaptdl_b200.env import num_restarts
    from adaptdl_b200.torch.epoch import (remaining_epochs_until,
                                          current_epoch, finished_epochs)
    assert current_epoch() is None
    if num_restarts() == 0:
        assert finished_epochs() == 0
        expected = list(range(6))
    else:
        assert finished_epochs() == 5
        expected = list(range(5, 10))
    for idx, epoch in enumerate(remaining_epochs_until(10)):
        assert epoch == expected[idx] == current_epoch() == finished_epochs()
        with pytest.raises(RuntimeError):
            next(remaining_epochs_until(20))
        if num_restarts() == 0 and epoch == 5:
            checkpoint.save_all_states()
            return 5
    assert finished_epochs() == 10
    assert list(remaining_epochs_until(10)) == []
    assert list(remaining_epochs_until(12)) == [10, 11]
    return 0


@elastic_multiprocessing
def test_dataloader_restarts():
    from adaptdl_b200 import checkpoint, collective
    from adaptdl_b200.env import num_restarts, num_replicas
    collective.initialize()
    dataset_size, init_batch_size = 100, 10
    dataset = TensorDataset(torch.rand(dataset_size))
    dataloader = AdaptiveDataLoader(dataset, batch_size=init_batch_size)
    # 2 batches at 1 replica (20 samples), 5 batches at 4 replicas (local 3,
    # 60 samples), the remaining 20 samples at 2 replicas (2 batches).
    assert current_dataloader() is None
    for idx, batch in enumerate(dataloader):
        if num_restarts() == 0 and idx == 2:
            checkpoint.save_all_states()
            return 4
        if num_restarts() == 1 and idx == 5:
            checkpoint.save_all_states()
            return 2
        assert current_dataloader() is dataloader._elastic
        local_bsz = batch[0].size(0)
        assert dataloader.current_local_bsz == local_bsz
        assert local_bsz == math.ceil(init_batch_size / num_replicas())
        assert dataloader.current_batch_size == num_replicas() * local_bsz
    assert idx == 1
    assert dataloader.current_local_bsz is None
    return 0


@elastic_multiprocessing
def test_dataloader_break():
    from adaptdl_b200 import collective
    from adaptdl_b200.env import num_restarts
    if num_restarts() == 0:
        return 2
    collective.initialize()
    dataloader = AdaptiveDataLoader(TensorDataset(torch.rand(100)),
                                    batch_size=10)
    for idx, batch in enumerate(dataloader):
        assert current_dataloader() is dataloader._elastic
        if idx == 5:
            break
    assert current_dataloader() is None
    for idx, batch in enumerate(dataloader):
        pass
    assert idx == 9
    return 0


@elastic_multiprocessing
def test_dataloader_skipdone_replay():
    """A loop that finished before the checkpoint is skipped on replay."""
    from adaptdl_b200 import checkpoint, collective
    from adaptdl_b200.env import num_restarts
    from adaptdl_b200.torch.epoch import remaining_epochs_until
    collective.initialize()
    train = AdaptiveDataLoader(TensorDataset(torch.rand(40)), batch_size=10)
    valid = AdaptiveDataLoader(TensorDataset(torch.rand(20)), batch_size=10)
    seen = collections.Counter()
    for epoch in remaining_epochs_until(2):
        for batch in train:
            seen["train"] += 1
        for idx, batch in enumerate(valid):
            seen["valid"] += 1
            if num_restarts() == 0 and epoch == 0 and idx == 0:
                checkpoint.save_all_states()
                return 2
    if num_restarts() == 1:
        # epoch 0: train loop skipped (it finished before the checkpoint);
        # valid restarts from index 0 (the checkpoint was taken while its
        # first batch was still being processed): 2 batches per epoch.
        assert seen["train"] == 4 and seen["valid"] == 2 + 2
    return 0


def test_dataloader_rejects_samplers_and_bad_bounds():
    import os
    os.environ.pop("ADAPTDL_CHECKPOINT_PATH", None)
    ds = TensorDataset(torch.rand(10))
    with pytest.raises(ValueError):
        AdaptiveDataLoader(ds, batch_size=2,
                           sampler=torch.utils.data.SequentialSampler(ds))
    loader = AdaptiveDataLoader(ds, batch_size=4)
    with pytest.raises(ValueError):
        loader.autoscale_batch_size(2)
    with pytest.raises(ValueError):
        loader.autoscale_batch_size(8, local_bsz_bounds=(5, 8))
    with pytest.raises(ValueError):
        loader.autoscale_batch_size(8, local_bsz_bounds=(1, 3))
    loader.autoscale_batch_size(8, local_bsz_bounds=(1, 4),
                                gradient_accumulation=True)
    assert loader.training


@elastic_multiprocessing
def test_accumulator_restarts():
    from adaptdl_b200 import checkpoint, collective
    from adaptdl_b200.env import num_restarts, replica_rank
    from adaptdl_b200.torch.accumulator import Accumulator
    collective.initialize()
    accum = Accumulator()
    if num_restarts() == 0:
        accum["a"] += 15
    assert "a" not in accum
    with accum.synchronized():
        assert "a" in accum and accum["a"] == 15
    assert "a" not in accum
    if num_restarts() == 0:
        accum["a"] -= 5
        checkpoint.save_all_states()
        return 4
    if num_restarts() == 1:
        accum.update({"a": replica_rank(), "b": replica_rank()})
    assert len(accum) == 0
    with accum.synchronized():
        assert len(accum) == 2
        assert accum["a"] == 16 and accum["b"] == 6
    assert len(accum) == 0
    if num_restarts() == 1:
        checkpoint.save_all_states()
        return 2
    if num_restarts() == 2:
        accum -= {"b": 5, "c": 5}
    with accum.synchronized():
        assert accum["a"] == 16 and accum["b"] == -4 and accum["c"] == -10
        accum.clear()
    with accum.synchronized():
        assert not accum
    with pytest.raises(TypeError):
        accum["x"] = 3
    return 0


@pytest.mark.parametrize("num_replicas", [1, 3])
@elastic_multiprocessing
def test_profile(num_replicas):
    from adaptdl_b200 import checkpoint
    from adaptdl_b200.env import num_restarts
    from adaptdl_b200.torch._metrics import (
        profile_step_start, profile_sync_time, profile_step_commit,
        _metrics_state)
    if num_restarts() == 0:
        profile = _metrics_state().profile
        assert len(profile) == 0
        profile_step_start(1)            # never committed
        profile_sync_time(1.0)
        profile_step_start(2)
        profile_sync_time(1.0)
        profile_sync_time(2.0)
        profile_step_commit()
        key = (1, 1, 2)
        assert len(profile) == 1
        assert profile[key]["accum_count"] == 0
        assert profile[key]["optim_count"] == 1
        assert profile[key]["optim_sync_time"] == 3.0
        assert profile[key]["optim_step_time"] >= 0.0
        checkpoint.save_all_states()
        return num_replicas
    profile = _metrics_state().profile
    key = (1, 1, 2)
    assert len(profile) == 1 and profile[key]["optim_sync_time"] == 3.0
    profile_step_start(3)
    profile_sync_time(2.0)
    profile_sync_time(3.0)
    profile_step_commit()
    key = (1, num_replicas, 3)
    old = profile[key]["optim_step_time"]
    profile_step_start(3)
    profile_sync_time(3.0)
    profile_sync_time(4.0)
    profile_step_commit(step_time=0.5)     # device-timed override
    assert len(profile) == 2
    assert profile[key]["optim_count"] == 2
    assert profile[key]["optim_sync_time"] == 12.0
    assert profile[key]["optim_step_time"] == pytest.approx(old + 0.5)
    return 0


@elastic_multiprocessing
def test_profile_accumulation_and_fit():
    from adaptdl_b200 import checkpoint
    from adaptdl_b200.env import num_restarts
    from adaptdl_b200.torch import _metrics
    from adaptdl_b200.torch._metrics import (
        profile_step_start, profile_sync_time, profile_step_commit,
        _metrics_state, _fit_perf_params)
    if num_restarts() == 0:
        for bsz, sync in ((2, 4.0), (5, 6.0)):
            for _ in range(2):
                profile_step_start(bsz)
                profile_step_commit(accumulation_step=True, step_time=0.1)
            profile_step_start(bsz)
            profile_sync_time(sync)
            profile_step_commit(accumulation_step=False, step_time=0.2)
        profile = _metrics_state().profile
        assert len(profile) == 2
        assert profile[(1, 1, 2)]["accum_count"] == 2
        assert profile[(1, 1, 2)]["optim_count"] == 1
        profile_step_start(3)              # accumulation only, no optim yet
        profile_step_commit(accumulation_step=True)
        _fit_perf_params()                 # sync > step is clamped, no crash
        assert _metrics_state().perf_params is not None
        _metrics.set_batch_size(4, 64, (2, 8), True)
        _metrics.update_grad_params("k", 0.5, 0.25)
        hints = _metrics._build_sched_hints()
        assert hints["initBatchSize"] == 4 and hints["maxBatchSize"] == 64
        assert hints["gradParams"] == {"norm": 0.5, "var": 0.25}
        assert hints["maxProfiledReplicas"] == 1
        assert set(hints["perfParams"]) == {
            "alpha_c", "beta_c", "alpha_n", "beta_n", "alpha_r", "beta_r",
            "gamma"}
        assert _metrics.get_goodput_fn() is not None
        checkpoint.save_all_states()
        return 2
    state = _metrics_state()
    assert len(state.profile) == 3
    assert state.perf_params is not None and state.grad_params == (0.5, 0.25)
    assert state.local_bsz_bounds == (2, 8) and state.gradient_accumulation
    return 0


@elastic_multiprocessing
def test_bptt_iterator():
    """500 tokens, bsz 10, bptt 5: one (5x10) batch at 1 replica, restart,
    then (5x5) windows on 2 replicas until the stream is consumed."""
    from adaptdl_b200 import checkpoint, collective, env
    from adaptdl_b200.torch.iterator import AdaptiveBPTTIterator
    collective.initialize()
    tokens = torch.arange(500)
    it = AdaptiveBPTTIterator(tokens, batch_size=10, bptt_len=5)
    seen = 0
    for idx, batch in enumerate(it):
        if env.num_restarts() == 0 and idx == 1:
            assert batch.text.shape == (5, 10)
            assert torch.equal(batch.target[:-1], batch.text[1:])
            checkpoint.save_all_states()
            return 2
        if env.num_replicas() == 2:
            assert batch.text.shape in ((5, 5), (4, 5))
            assert torch.equal(batch.target[:-1], batch.text[1:])
        seen += 1
    if env.num_replicas() == 2:
        # fold is 100 rows x 5 cols; resumed at row ceil(5*100/50)=10;
        # rows 10..98 in windows of 5 shared by 2 replicas -> 9 steps each
        assert idx == 8 and seen == 9
    return 0


def test_periodic_report_does_not_stall_the_step_loop(monkeypatch):
    """The fit and the hints PUT run off-thread: a slow supervisor must not
    cost the training loop anything (the reference does both inline)."""
    import time
    from adaptdl_b200 import checkpoint
    from adaptdl_b200.torch import _metrics
    _metrics._reset_for_tests()
    stale = checkpoint._NAMES_TO_STATES.get("adaptdl-metrics")
    if stale is not None:          # left behind by an in-process test
        stale.unregister()
    posted = []

    def slow_post(hints, job):
        time.sleep(0.5)
        posted.append(hints)
    monkeypatch.setattr(_metrics, "post_sched_hints", slow_post)
    monkeypatch.setattr(_metrics, "REPORT_PERIOD_S", 0.0)
    monkeypatch.setattr(_metrics, "ASYNC_REPORT", True)
    _metrics.set_batch_size(32, 256, (8, 64), False)
    _metrics.update_grad_params("k", 1.0, 2.0)
    try:
        slowest = 0.0
        for step in range(6):
            _metrics.profile_step_start(16)
            began = time.time()
            _metrics.profile_step_commit(step_time=0.01)
            slowest = max(slowest, time.time() - began)
            time.sleep(0.01)
        assert slowest < 0.25, slowest          # never waited for the PUT
        _metrics.wait_for_report(10.0)
        assert posted and posted[0]["initBatchSize"] == 32
        assert posted[0]["maxProfiledReplicas"] == 1
        assert _metrics._metrics_state().perf_params is not None
        assert len(posted) <= 2                  # in-flight reports not doubled
        # the inline mode still works
        monkeypatch.setattr(_metrics, "ASYNC_REPORT", False)
        count = len(posted)
        _metrics.profile_step_start(16)
        _metrics.profile_step_commit(step_time=0.01)
        assert len(posted) == count + 1
    finally:
        _metrics.wait_for_report(10.0)
        _metrics._reset_for_tests()


def _beat_loop(iterations, signal_at, period, pause=0.001):
    """Run the preemption beat like a training loop does; returns what each
    replica saw: (iteration of the exit, consensus rounds started)."""
    import time
    from adaptdl_b200 import _signal, collective, env
    from adaptdl_b200.torch import data
    os.environ["ADAPTDL_HEARTBEAT_PERIOD"] = str(period)
    collective.initialize(env.master_addr(), env.master_port(),
                          env.replica_rank(), env.num_replicas())
    beat = data._PREEMPTION
    rounds, launch = [0], collective.allreduce_async

    def counting(*args, **kwargs):
        rounds[0] += 1
        return launch(*args, **kwargs)
    data.collective.allreduce_async = counting
    try:
        for iteration in range(iterations):
            if iteration == signal_at and env.replica_rank() == 1:
                _signal.set_exit_flag(True)          # only ONE replica
            try:
                beat.beat()
            except SystemExit as stop:
                assert stop.code == 143
                return iteration, rounds[0]
            time.sleep(pause)
        return None, rounds[0]
    finally:
        data.collective.allreduce_async = launch


@elastic_multiprocessing
def test_preemption_beat_every_iteration_like_the_reference():
    from adaptdl_b200 import env
    if env.num_restarts() == 0:
        return 2
    stopped_at, rounds = _beat_loop(40, signal_at=10, period=0)
    # flagged in round 10, resolved by everyone in iteration 11
    assert stopped_at == 11 and rounds == 11, (stopped_at, rounds)
    return 0


@elastic_multiprocessing
def test_preemption_beat_leaves_the_step_path_but_stays_in_lockstep():
    from adaptdl_b200 import collective, env
    if env.num_restarts() == 0:
        return 2
    stopped_at, rounds = _beat_loop(3000, signal_at=600, period=0.05)
    assert stopped_at is not None and 600 < stopped_at < 900, stopped_at
    # ~1 ms iterations, a round every ~50 ms: far fewer rounds than iterations
    assert rounds < stopped_at / 5, (rounds, stopped_at)
    # every replica left at the SAME iteration
    both = collective.allreduce([stopped_at], lambda a, b: a + b)
    assert both[0] == both[1], both
    return 0


@elastic_multiprocessing
def test_tensor_dataset_batches_are_fetched_whole():
    """In-memory TensorDatasets take the batched fast path; the batches are
    exactly what the per-sample path produces."""
    from adaptdl_b200.torch import data
    from adaptdl_b200.torch.data import AdaptiveDataLoader
    torch.manual_seed(0)
    x, y = torch.randn(50, 3, 4), torch.arange(50)

    def batches(flag):
        os.environ["ADAPTDL_B200_BATCHED_TENSOR_DATASET"] = flag
        loader = AdaptiveDataLoader(TensorDataset(x, y), batch_size=8,
                                    shuffle=True)
        assert isinstance(loader.dataset,
                          data._BatchedTensorDataset) == (flag == "1")
        sampler = loader.batch_sampler.sampler
        sampler.set_epoch(0)
        return [loader.collate_fn(
                    loader.dataset.__getitems__(idx) if flag == "1"
                    else [loader.dataset[i] for i in idx])
                for idx in loader.batch_sampler]
    fast, slow = batches("1"), batches("0")
    assert len(fast) == len(slow) == 7
    for a, b in zip(fast, slow):
        assert type(a) is type(b) is list and len(a) == 2
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # a custom collate_fn or a dataset subclass keeps the generic path
    loader = AdaptiveDataLoader(TensorDataset(x, y), batch_size=8,
                                collate_fn=lambda samples: samples)
    assert type(loader.dataset) is TensorDataset
    return 0


def test_sampler_partition_properties():
    """Property test: for any dataset size, replica count and resume index,
    the replicas' shares have equal length, only valid indices, and together
    cover exactly the rest of the pass (plus at most one pad each)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(n=st.integers(1, 60), replicas=st.integers(1, 12),
           passes=st.integers(0, 2), offset=st.integers(0, 59),
           shuffle=st.booleans(), epoch=st.integers(0, 3))
    def check(n, replicas, passes, offset, shuffle, epoch):
        index = passes * n + offset % n
        shares = []
        for rank in range(replicas):
            sampler = ElasticSampler(list(range(n)), shuffle=shuffle)
            sampler.num_replicas, sampler.rank = replicas, rank
            sampler.set_epoch(epoch, index)
            share = list(sampler)
            assert len(share) == len(sampler)
            assert all(0 <= i < n for i in share)
            shares.append(share)
        assert len({len(s) for s in shares}) == 1
        order = shares and ElasticSampler(list(range(n)), shuffle=shuffle)
        order.set_epoch(epoch, index)
        rest = order._order()[index % n:]
        real = [s[:len(rest[r::replicas])] for r, s in enumerate(shares)]
        merged = sorted(i for s in real for i in s)
        assert merged == sorted(rest)
    check()


@elastic_multiprocessing
def test_dataloader_visits_every_sample_for_arbitrary_sizes():
    """Property test (single replica): for any dataset size, batch size and
    drop_last, one epoch visits every sample once (or drops only a final
    partial batch), in batches no larger than asked for."""
    from hypothesis import given, settings, strategies as st
    import adaptdl_b200.torch as adl
    adl.init_process_group("gloo")
    seen_epochs = []

    @settings(max_examples=40, deadline=None)
    @given(n=st.integers(1, 70), batch=st.integers(1, 20),
           drop_last=st.booleans(), shuffle=st.booleans())
    def check(n, batch, drop_last, shuffle):
        data = TensorDataset(torch.arange(n))
        loader = AdaptiveDataLoader(data, batch_size=batch, shuffle=shuffle,
                                    drop_last=drop_last)
        epoch = len(seen_epochs)
        seen_epochs.append(epoch)
        for _ in adl.remaining_epochs_until(epoch + 1):
            got = []
            for (values,) in loader:
                assert 1 <= len(values) <= batch
                got.extend(values.tolist())
        if drop_last:
            assert len(got) == n // batch * batch
            assert len(set(got)) == len(got) and set(got) <= set(range(n))
        else:
            assert sorted(got) == list(range(n))
    check()
    return 0


@elastic_multiprocessing
def test_preemption_beat_relearns_its_pace_in_every_loop():
    import time
    from adaptdl_b200 import collective, env
    from adaptdl_b200.torch import data
    if env.num_restarts() == 0:
        return 2
    os.environ["ADAPTDL_HEARTBEAT_PERIOD"] = "0.05"
    collective.initialize(env.master_addr(), env.master_port(),
                          env.replica_rank(), env.num_replicas())
    beat = data._PREEMPTION
    for _ in range(400):                       # a fast loop: interval grows
        beat.beat()
        time.sleep(0.0005)
    assert 4 < beat.interval <= beat.MAX_INTERVAL
    beat.new_loop()                            # e.g. the validation loop
    beat.beat()
    assert beat.interval == 1 and beat.countdown == 0
    for _ in range(3):                         # slow iterations: stays at 1
        time.sleep(0.06)
        beat.beat()
        assert beat.interval == 1
    return 0


def test_accumulator_checkpoint_layout_and_delta_protocol(tmp_path):
    """The file an Accumulator writes is the reference's
    ``(history: {epoch: [snapshot...]}, results)`` pickle; in-loop sessions
    take no replay slot; deltas are immutable values."""
    import io
    import pickle
    from adaptdl_b200 import collective
    from adaptdl_b200.torch import accumulator as acc_mod
    collective.initialize()
    try:
        acc_mod._reset_for_tests()
        accum = acc_mod.Accumulator(seen=0)
        accum["seen"] += 2
        accum["loss"] += 1.5
        accum["loss"] -= 0.5
        delta = accum["loss"]
        assert (delta + 1).amount == 1 and delta.amount == 0   # immutable
        with pytest.raises(AttributeError):
            delta.amount = 7
        with pytest.raises(TypeError):
            accum["loss"] = accum["seen"] + accum["loss"]
        with accum.synchronized():
            assert dict(accum) == {"seen": 2, "loss": 1.0}
            with accum.synchronized():        # re-entrant
                accum["extra"] = "set inside"
        accum["seen"] += 1
        with accum.synchronized():
            assert accum["seen"] == 3 and accum["extra"] == "set inside"
        buf = io.BytesIO()
        accum._ledger.save(buf)
        history, results = pickle.loads(buf.getvalue())
        assert results == {"seen": 3, "loss": 1.0, "extra": "set inside"}
        assert list(history) == [None] and len(history[None]) == 2
        assert history[None][0]["seen"] == 2 and history[None][1]["seen"] == 3
        # a fresh accumulator loading that file replays both positions
        accum._ledger.unregister()
        acc_mod._reset_for_tests()
        again = acc_mod.Accumulator()
        again._ledger.load(io.BytesIO(buf.getvalue()))
        again["seen"] += 100                  # re-executed work: dropped
        with again.synchronized():
            assert again["seen"] == 2
        with again.synchronized():
            assert again["seen"] == 3
        again["seen"] += 1                    # new work after the replay
        with again.synchronized():
            assert again["seen"] == 4
        again._ledger.unregister()
    finally:
        collective.teardown()
        acc_mod._reset_for_tests()


def test_data_parallel_state_file_is_fast_and_still_plain_torch(tmp_path):
    """The model/optimizer state is written without zip CRCs and read back
    memory-mapped; the file is still something a plain ``torch.load`` (what
    the reference calls) understands, and file objects without a path fall
    back to ordinary reading."""
    import io
    import zipfile
    from adaptdl_b200.torch import parallel
    payload = ([{"w": torch.arange(12.0).reshape(3, 4)},
                {"state": {"gns": {"progress": 1.5}}}, None, None], 1.25, 0.5)
    path = tmp_path / "adaptdl-dataparallel"
    with open(path, "wb") as f:
        parallel._save_fast(payload, f)
    # an ordinary torch.save archive (CRC fields are simply left empty)
    with zipfile.ZipFile(path) as z:
        assert any(n.endswith("data.pkl") for n in z.namelist())
    plain = torch.load(str(path), weights_only=False)
    assert plain[1:] == (1.25, 0.5)
    assert torch.equal(plain[0][0]["w"], payload[0][0]["w"])
    with open(path, "rb") as f:
        mapped = parallel._load_fast(f)
    assert mapped[1:] == (1.25, 0.5)
    assert torch.equal(mapped[0][0]["w"], payload[0][0]["w"])
    assert mapped[0][1]["state"]["gns"]["progress"] == 1.5
    # an in-memory file object: no path to map
    buf = io.BytesIO(path.read_bytes())
    again = parallel._load_fast(buf)
    assert torch.equal(again[0][0]["w"], payload[0][0]["w"])
    # the process-wide CRC option is left as it was found
    get = getattr(torch.serialization, "get_crc32_options", None)
    if get is not None:
        assert get() is True


def test_checkpoint_fsync_can_be_switched_off(tmp_path, monkeypatch):
    from adaptdl_b200 import checkpoint
    synced = []
    monkeypatch.setattr(os, "fsync", lambda fd: synced.append(fd))

    class Tiny(checkpoint.State):
        def save(self, f):
            f.write(b"x")

        def load(self, f):
            pass
    state = Tiny("tiny-fsync-state")
    for flag, expect in ((None, 1), ("0", 0), ("1", 1)):
        if flag is None:
            monkeypatch.delenv("ADAPTDL_CHECKPOINT_FSYNC", raising=False)
        else:
            monkeypatch.setenv("ADAPTDL_CHECKPOINT_FSYNC", flag)
        del synced[:]
        checkpoint.save_state(state, str(tmp_path))
        assert len(synced) == expect
