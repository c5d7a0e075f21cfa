"""Elastic, adaptive-batch-size data loading.

* :class:`ElasticSampler` -- partitions a dataset over replicas with a
  deterministic per-(epoch, pass) shuffle and can resume at any global sample
  index after a restart at a *different* replica count.
* :class:`AdaptiveDataLoaderHelper` -- the per-iteration heartbeat shared by
  all loaders: exit-flag consensus -> checkpoint -> ``exit(143)``, step
  profiling, gradient-accumulation bookkeeping, goodput-driven choice of
  ``(local_bsz, accumulation_steps)``, loop-position tracking for replay.
* :class:`AdaptiveDataLoader` -- a ``torch.utils.data.DataLoader`` whose
  ``batch_size`` is the *global* target batch size.

Parity: reference ``adaptdl/adaptdl/torch/data.py:41-575`` (same public names,
semantics and checkpoint format). B200 additions: :class:`DevicePrefetcher`
(pinned-host -> HBM copies on a side stream, one batch ahead).
"""

import collections
import logging
import math
import os
import random
import sys
import time
from contextlib import contextmanager

import numpy as np
import torch
from torch.utils.data import DataLoader, Sampler

from adaptdl_b200 import checkpoint, collective, env
from adaptdl_b200._signal import get_exit_flag
from adaptdl_b200.torch import _metrics
from adaptdl_b200.torch.epoch import current_epoch

LOG = logging.getLogger(__name__)

EXIT_CODE_PREEMPTED = 143     # the contract with the job controller


def _leave_preempted():
    """The checkpoint is on disk: exit with the preemption code.

    ``ADAPTDL_B200_FAST_EXIT=1`` (set by ``sched.local --fast-exit``) skips
    the interpreter's and torch's own teardown -- module destructors, process
    group and CUDA context destruction: 0.7 s of a CPU replica's 1.2 s exit,
    more with a CUDA context -- after running the registered ``atexit``
    handlers and flushing the standard streams; the next generation cannot
    start before every replica of this one is gone."""
    from adaptdl_b200.utils import rescale_trace
    rescale_trace.mark("exiting")
    if os.environ.get("ADAPTDL_B200_FAST_EXIT") == "1":
        import atexit
        try:
            atexit._run_exitfuncs()
            logging.shutdown()
            sys.stdout.flush()
            sys.stderr.flush()
        finally:
            os._exit(EXIT_CODE_PREEMPTED)
    sys.exit(EXIT_CODE_PREEMPTED)


class _PreemptionBeat(object):
    """Exit-flag consensus of the replicas, off the step path.

    The reference OR-reduces the flag over its TCP reducer on EVERY training
    iteration (``torch/data.py:311-334``): a host-side barrier per step whose
    messages all pass through a Python thread on rank 0 -- irrelevant at its
    100 ms steps, a real cost at the 2 ms steps of a graph-replayed B200 job
    (8 replicas = 8000 messages/s on rank 0's interpreter). Here a round is
    started only every ``interval`` iterations, with ``interval`` chosen so
    that rounds are about ``ADAPTDL_HEARTBEAT_PERIOD`` seconds apart (default
    0.1; ``0`` = every iteration, as the reference). The interval must be the
    same on every replica, so it travels WITH the consensus: each round
    reduces ``(flag, rank 0's suggestion for the next interval)``, and every
    replica resolves round k and starts round k+1 at the same iteration. A
    signal is therefore acted upon within two rounds (~0.2 s) -- far inside
    any preemption grace period.

    One instance per process: rounds continue across data loaders (training
    and validation loops alternate), so short loops cannot starve it. The
    interval is only trusted while the pace is: every ``for batch in loader``
    loop starts over at one iteration per round (its batch size, model mode
    or dataset may differ from the previous loop's), and the interval never
    exceeds ``MAX_INTERVAL`` iterations, so that a sudden slow-down cannot
    push the reaction time towards a preemption grace period.
    """

    MAX_INTERVAL = 64

    def __init__(self):
        self.reset()

    def reset(self):
        self.pending = None          # future of the round in flight
        self.interval = 1            # iterations between round starts
        self.countdown = 0           # iterations until the next round start
        self.started = None          # host time the pending round started
        self.relearn = False

    @staticmethod
    def period():
        return float(os.environ.get("ADAPTDL_HEARTBEAT_PERIOD", "0.1"))

    def _suggestion(self, now):
        """Interval that would have put the rounds ``period`` apart at the
        pace of the iterations since the last round started."""
        period = self.period()
        if period <= 0 or self.started is None:
            return 1
        per_iteration = max(now - self.started, 1e-6) / self.interval
        return int(min(max(period / per_iteration, 1), self.MAX_INTERVAL))

    def new_loop(self):
        """A data-loader loop begins (same iteration on every replica): the
        next beat resolves the round in flight and the pace is re-learned."""
        self.countdown = 0
        self.interval = 1
        self.started = None
        self.relearn = True      # the round in flight carries the OLD pace

    def beat(self):
        """Call once per training iteration on every replica."""
        if self.countdown > 0:
            self.countdown -= 1
            return
        now = time.monotonic()
        suggestion = self._suggestion(now)
        if self.pending is not None:
            flagged, agreed = self.pending.result()
            if flagged:
                from adaptdl_b200.utils import rescale_trace
                rescale_trace.mark("exit_consensus")
                checkpoint.save_all_states()
                rescale_trace.mark("checkpoint_written")
                _leave_preempted()
            self.interval = 1 if self.relearn else agreed
        self.relearn = False
        # rank 0 is the left-most operand of the fold: its suggestion wins
        self.pending = collective.allreduce_async(
            (bool(get_exit_flag()), suggestion),
            lambda a, b: (a[0] or b[0], a[1]))
        self.started = now
        self.countdown = self.interval - 1


_PREEMPTION = _PreemptionBeat()


def _shuffle_seed(epoch, pass_index):
    # Compatibility point, not a design choice: the reference seeds the
    # permutation with CPython's hash of the (epoch, pass) tuple
    # (``torch/data.py:74``). Using the same number means a job preempted
    # mid-epoch under the reference and resumed here (or the other way round)
    # continues on exactly the samples that were left
    # (tests/test_reference_interop.py). Tuples of ints hash independently of
    # PYTHONHASHSEED.
    return hash((epoch, pass_index))


class ElasticSampler(Sampler):
    """Partition ``dataset`` across replicas; resumable at a global index.

    ``index`` counts samples consumed by *all* replicas in the current loop,
    so a checkpoint taken at N replicas resumes correctly at M replicas.

    Arguments:
        dataset: the dataset (only ``len()`` is used).
        shuffle (bool): shuffle deterministically per (epoch, pass).
    """

    def __init__(self, dataset, shuffle=True):
        self.dataset = dataset
        self.shuffle = shuffle
        self.num_replicas = env.num_replicas()
        self.rank = env.replica_rank()
        self.epoch = 0
        self.index = 0

    def _order(self):
        n = len(self.dataset)
        if not self.shuffle:
            return list(range(n))
        gen = torch.Generator()
        gen.manual_seed(_shuffle_seed(self.epoch or 0, self.index // n))
        return torch.randperm(n, generator=gen).tolist()

    def __iter__(self):
        """Indices of this replica's samples, from the set index onward."""
        order = self._order()
        base = self.index % len(self.dataset)
        mine = order[base + self.rank::self.num_replicas]
        if len(mine) < len(self):       # pad so all replicas are equal
            # (modulo: a dataset smaller than the replica count -- a tiny
            # validation set on many GPUs -- must not index past its end)
            mine.append(order[self.rank % len(order)])
        assert len(mine) == len(self)
        return iter(mine)

    def __len__(self):
        """Samples this replica still has to visit in the current pass."""
        base = self.index % len(self.dataset)
        return math.ceil((len(self.dataset) - base) / self.num_replicas)

    def set_epoch(self, epoch, index=0):
        """Select the epoch (shuffle order) and the global start index."""
        self.epoch = epoch
        self.index = index


def current_dataloader():
    """The :class:`AdaptiveDataLoaderHelper` currently being iterated."""
    return AdaptiveDataLoaderHelper._current


class _AdaptiveDataLoaderState(checkpoint.PickledFields):
    FIELDS = ("current_index", "end_index", "last_position")
    LAYOUT = "tuple"

    # Loaders are created in the same order on every replica; the name is
    # derived from (epoch at creation, creation ordinal within that epoch).
    init_count = collections.Counter()

    def __init__(self):
        if current_dataloader() is not None:
            raise RuntimeError("a data loader must be created outside of "
                               "data-loader loops (replicas create their "
                               "loaders in the same order)")
        epoch = current_epoch()
        counter = _AdaptiveDataLoaderState.init_count
        super().__init__("adaptdl-dataloader-epoch{}-{}".format(
            epoch, counter[epoch]))
        counter[epoch] += 1
        self.current_index = 0    # samples consumed in the current loop
        self.end_index = 0        # optional, for custom loaders
        self.last_position = {}   # epoch -> position of last finished loop
        self.current_local_bsz = 0
        self.accumulation_steps = 0


class AdaptiveDataLoaderHelper(object):
    """Fine-grained control over adaptive training loops; the building block
    for custom loaders (see :class:`AdaptiveDataLoaderMixin`).

    Arguments:
        batch_size (int): target *global* batch size.
    """

    _position = collections.Counter()  # epoch -> loops finished (all loaders)
    _training = None                   # the loader that feeds training
    _current = None                    # the loader being iterated

    SPEEDUP_THRESHOLD = 1.05

    def __init__(self, batch_size=1):
        self._max_batch_size = None
        self._local_bsz_bounds = None
        self._state = _AdaptiveDataLoaderState()
        checkpoint.load_state(self._state)
        self.batch_size = batch_size
        self._gradient_accumulation = False
        self._speedup_threshold = self.SPEEDUP_THRESHOLD
        self._accum_count = 0
        self._step_time_source = None   # optional device-side step timer

    # -- positions and batch sizes (views of the checkpointed record) ------

    def _is_current(self):
        return AdaptiveDataLoaderHelper._current is self

    #: samples consumed so far in the running loop, over all replicas; reads
    #: as ``None`` and ignores writes unless THIS loader is being iterated
    current_index = property(
        lambda self: self._state.current_index if self._is_current()
        else None,
        lambda self, value: setattr(self._state, "current_index", value)
        if self._is_current() else None)
    #: free slot for custom loaders (e.g. the end of a BPTT stream)
    end_index = property(
        lambda self: self._state.end_index,
        lambda self, value: setattr(self._state, "end_index", value))
    #: largest adaptive global batch size, ``None`` while adaptation is off
    max_batch_size = property(lambda self: self._max_batch_size)
    #: ``(min, max)`` micro-batch size per replica
    local_bsz_bounds = property(lambda self: self._local_bsz_bounds)
    #: micro-batch size per replica in use
    current_local_bsz = property(lambda self: self._state.current_local_bsz)
    #: additional micro-batches accumulated into every optimizer step
    accumulation_steps = property(
        lambda self: self._state.accumulation_steps)

    @property
    def current_batch_size(self):
        """Samples per optimizer step over the whole job."""
        micro_steps = self.accumulation_steps + 1
        return env.num_replicas() * micro_steps * self.current_local_bsz

    def is_accum_step(self):
        """This step's gradient is only accumulated (no sync, no update)."""
        return self._accum_count < self._state.accumulation_steps

    def is_optim_step(self):
        """This step ends with a gradient sync and an optimizer update."""
        return not self.is_accum_step()

    def train(self):
        """Mark this loader as the one feeding training (first wins)."""
        if AdaptiveDataLoaderHelper._training is None:
            AdaptiveDataLoaderHelper._training = self
        _metrics.set_batch_size(self.batch_size, self.max_batch_size,
                                self.local_bsz_bounds,
                                self._gradient_accumulation)

    @property
    def training(self):
        return self is AdaptiveDataLoaderHelper._training

    def autoscale_batch_size(self, max_batch_size, local_bsz_bounds=None,
                             gradient_accumulation=False):
        """Enable adaptive batch size.

        Arguments:
            max_batch_size (int): largest global batch size allowed.
            local_bsz_bounds (tuple): ``(min, max)`` per-replica batch size.
            gradient_accumulation (bool): allow accumulation to go beyond
                the per-replica maximum.

        Raises:
            ValueError: on inconsistent bounds.
        """
        if not isinstance(max_batch_size, int) \
                or max_batch_size < self.batch_size:
            raise ValueError("invalid max_batch_size")
        if local_bsz_bounds is not None:
            lo, hi = local_bsz_bounds
            if (lo is not None and lo > self.batch_size) or \
                    (hi is not None and hi < self.batch_size):
                raise ValueError("invalid local_bsz_bounds")
        self._max_batch_size = max_batch_size
        self._local_bsz_bounds = local_bsz_bounds
        self._gradient_accumulation = gradient_accumulation
        self.train()

    def _sync_local_bsz(self):
        """Decide ``(local_bsz, accumulation_steps)`` for the next pass and
        agree on rank 0's choice."""
        goodput_fn = _metrics.get_goodput_fn()
        state = self._state
        if self.max_batch_size is None or goodput_fn is None:
            state.current_local_bsz = math.ceil(
                self.batch_size / env.num_replicas())
            state.accumulation_steps = 0
        else:
            suggestion = goodput_fn.optimize(
                env.num_nodes(), env.num_replicas(),
                max_batch_size=self._max_batch_size,
                atomic_bsz_range=self._local_bsz_bounds,
                accumulation=self._gradient_accumulation)
            best_goodput, atomic_bsz, accum_steps = suggestion
            if not state.current_local_bsz:
                switch = True          # first decision: take the suggestion
            else:
                current = goodput_fn(env.num_nodes(), env.num_replicas(),
                                     self.current_local_bsz,
                                     self.accumulation_steps)
                switch = best_goodput / max(current, 1e-8) \
                    > self._speedup_threshold
            if switch:
                state.current_local_bsz = atomic_bsz
                state.accumulation_steps = accum_steps
        state.current_local_bsz, state.accumulation_steps = \
            collective.broadcast((state.current_local_bsz,
                                  state.accumulation_steps))
        return self.current_local_bsz

    # -- per-iteration heartbeat -------------------------------------------

    def set_step_time_source(self, fn):
        """Install a callable returning the last step's duration in seconds
        measured on the device (or ``None`` to fall back to wall-clock)."""
        self._step_time_source = fn

    @contextmanager
    def profile(self, commit):
        """Wrap every training iteration. Must be entered the same number
        of times on every replica.

        Takes part in the exit-flag consensus (:class:`_PreemptionBeat`): if
        any replica was signalled, *all* replicas checkpoint at the same
        iteration and exit with code 143.
        """
        _PREEMPTION.beat()
        _metrics.profile_step_start(self.current_local_bsz)
        yield
        if _metrics.device_timer() is not None:
            # steps are timed on the device and booked asynchronously
            _metrics.profile_step_commit_device(self.is_accum_step(), commit)
        elif commit:
            step_time = None
            if self._step_time_source is not None:
                step_time = self._step_time_source()
            _metrics.profile_step_commit(self.is_accum_step(),
                                         step_time=step_time)
        self._accum_count = (0 if self.is_optim_step()
                             else self._accum_count + 1)

    @contextmanager
    def context(self):
        """Wrap a whole loader iteration (one ``for batch in loader``)."""
        cls = AdaptiveDataLoaderHelper
        if cls._current is not None:
            raise RuntimeError("a data loader is already being iterated: "
                               "loops over adaptive loaders cannot nest")
        epoch = current_epoch()
        cls._current = self
        _PREEMPTION.new_loop()
        _metrics.device_timer_reset()     # no step interval across loops
        try:
            yield
        finally:
            # the loop (finished or abandoned) is over: rewind, and remember
            # its position so that a restarted job can skip it
            record = self._state
            record.current_index = record.end_index = 0
            record.last_position[epoch] = cls._position[epoch]
            cls._position[epoch] += 1
            cls._current = None

    def skipdone(self):
        """Call right after entering :meth:`context`: True if this loop had
        already completed before the last restart (and must be skipped)."""
        epoch = current_epoch()
        position = self._position[epoch]
        finished_before = self._state.last_position.get(epoch, -1)
        if position > finished_before:
            return False
        LOG.info("epoch %s: loop #%s of %s completed before the restart, "
                 "skipping it", epoch, position, type(self).__name__)
        self._position[epoch] += 1
        return True

    def to_tensorboard(self, writer, global_step, tag_prefix=""):
        """Write batch-size metrics to a TensorBoard ``SummaryWriter``."""
        prefix = tag_prefix.rstrip("/") + "/" if tag_prefix else ""
        for tag, value in (("Total_Batch_Size", self.current_batch_size),
                           ("Local_Batch_Size", self.current_local_bsz),
                           ("Accumulation_Steps", self.accumulation_steps)):
            writer.add_scalar(prefix + tag, value, global_step)


class AdaptiveDataLoaderMixin(object):
    """Give any custom loader elastic behaviour: it owns ``self._elastic``
    (an :class:`AdaptiveDataLoaderHelper`) and re-exports the user-facing
    knobs."""

    def __init__(self, batch_size):
        self._elastic = AdaptiveDataLoaderHelper(batch_size)

    def _iterating(self):
        return self._elastic._is_current()


def _forward_to_helper(cls):
    """Re-export the helper's user-facing methods and read-only views on the
    mix-in; the batch sizes only mean something while the loader runs."""
    def method(name):
        def call(self, *args, **kwargs):
            return getattr(self._elastic, name)(*args, **kwargs)
        call.__name__ = name
        call.__doc__ = getattr(AdaptiveDataLoaderHelper, name).__doc__
        return call

    def view(name, only_while_iterating):
        def read(self):
            if only_while_iterating and not self._iterating():
                return None
            return getattr(self._elastic, name)
        return property(read)
    for name in ("autoscale_batch_size", "to_tensorboard"):
        setattr(cls, name, method(name))
    for name, gated in (("current_local_bsz", True),
                        ("current_batch_size", True),
                        ("accumulation_steps", False), ("training", False)):
        setattr(cls, name, view(name, gated))
    return cls


_forward_to_helper(AdaptiveDataLoaderMixin)


class _SeededWorkerInit(object):
    """``worker_init_fn`` that first gives the (replica, worker) pair its own
    python / numpy / torch seed -- torch's per-worker base seed is the same
    on every replica -- and then runs the user's function, if any. A class
    (not a closure) so that it pickles for spawn-based workers."""

    def __init__(self, user_fn, num_workers):
        self.user_fn = user_fn
        self.stride = max(int(num_workers or 0), 1)

    def __call__(self, worker_id):
        seed = env.replica_rank() * self.stride + torch.initial_seed()
        random.seed(seed)
        np.random.seed(seed % (1 << 32))
        torch.manual_seed(seed)
        if self.user_fn is not None:
            return self.user_fn(worker_id)


class _BatchedTensorDataset(torch.utils.data.TensorDataset):
    """A ``TensorDataset`` that hands the loader a whole batch at once
    (``__getitems__`` = one ``index_select`` per tensor) instead of
    ``batch_size`` single-sample lookups followed by a ``stack``: for a
    128-sample batch that is the difference between ~0.5 ms and ~0.03 ms of
    interpreter time per step -- a quarter of a 2 ms training step."""

    def __getitems__(self, indices):
        index = torch.as_tensor(indices, dtype=torch.int64)
        return [tensor.index_select(0, index) for tensor in self.tensors]


def _already_batched(batch):
    return batch


def _batched_tensor_dataset(dataset, loader_kwargs):
    """The fast path applies to a plain in-memory ``TensorDataset`` collated
    the default way; the batches are the same objects the default path
    builds (a list with one stacked tensor per dataset tensor)."""
    if type(dataset) is not torch.utils.data.TensorDataset or \
            loader_kwargs.get("collate_fn") is not None or \
            os.environ.get("ADAPTDL_B200_BATCHED_TENSOR_DATASET", "1") == "0" \
            or any(t.is_sparse or t.device.type != "cpu"
                   for t in dataset.tensors):
        return dataset, loader_kwargs
    return _BatchedTensorDataset(*dataset.tensors), \
        dict(loader_kwargs, collate_fn=_already_batched)


class AdaptiveDataLoader(DataLoader, AdaptiveDataLoaderMixin):
    """Drop-in ``DataLoader`` with adaptive batch size and elasticity.

    Differences from ``torch.utils.data.DataLoader``:

    1. ``batch_size`` is the target *global* batch size over all replicas.
    2. Custom ``sampler`` / ``batch_sampler`` are not supported.
    3. Iterate only inside an epoch loop
       (:func:`adaptdl_b200.torch.remaining_epochs_until`); only one loader
       may be iterated at a time.

    Raises:
        ValueError: if ``sampler`` or ``batch_sampler`` is given.
    """

    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        for reserved in ("sampler", "batch_sampler"):
            if kwargs.get(reserved) is not None:
                raise ValueError(
                    "{!r} cannot be passed to AdaptiveDataLoader: it "
                    "partitions and re-partitions the dataset across "
                    "replicas with its own ElasticSampler".format(reserved))
        dataset, kwargs = _batched_tensor_dataset(dataset, kwargs)
        loader_kwargs = dict(
            kwargs, sampler=ElasticSampler(dataset, shuffle=shuffle),
            worker_init_fn=_SeededWorkerInit(kwargs.get("worker_init_fn"),
                                             kwargs.get("num_workers")))
        # the sampler shuffles; DataLoader itself must not
        DataLoader.__init__(self, dataset, batch_size, shuffle=False,
                            **loader_kwargs)
        AdaptiveDataLoaderMixin.__init__(self, batch_size)

    def _set_local_batch_size(self, local_bsz):
        # DataLoader forbids re-assigning batch_sampler, but mutating the
        # BatchSampler it built is allowed (and what the reference does).
        self.batch_sampler.batch_size = local_bsz

    def __iter__(self):
        """Yield batches until one epoch's worth of *progress* is made.

        Without adaptive batch size: one pass over this replica's 1/N share
        of the dataset. With it: until the scale-invariant progress equals
        one pass at the initial batch size, possibly more than one pass over
        the data. A checkpoint-restart may happen between any two batches.
        """
        epoch = current_epoch()
        num_replicas = env.num_replicas()
        elastic = self._elastic
        with elastic.context():
            if elastic.skipdone():
                return
            done = False
            while not done:
                self.sampler.set_epoch(epoch, index=elastic.current_index)
                self._set_local_batch_size(elastic._sync_local_bsz())
                for idx, batch in enumerate(super().__iter__()):
                    # the first batch of a pass is warm-up: never committed
                    with elastic.profile(self.training and idx >= 1):
                        yield batch
                        elastic.current_index += \
                            num_replicas * self.batch_sampler.batch_size
                        if elastic.max_batch_size is not None and \
                                _metrics.get_progress() >= \
                                len(self.dataset) * ((epoch or 0) + 1) \
                                / self.batch_size:
                            done = True
                            break
                if elastic.max_batch_size is None:
                    done = True
                # round up to the end of this pass over the dataset
                elastic.current_index -= \
                    elastic.current_index % -len(self.dataset)


class DevicePrefetcher(object):
    """Wrap a loader that yields (nested) CPU tensors: pin them, and copy the
    *next* batch to ``device`` on a side stream while the current one is
    being consumed (B200 addition; keeps the H2D copy off the step's critical
    path).

    The wrapped loader's elastic bookkeeping is untouched because this wraps
    the iterator, not the loader class.
    """

    def __init__(self, loader, device):
        self.loader = loader
        self.device = torch.device(device)
        self._stream = torch.cuda.Stream(self.device) \
            if self.device.type == "cuda" else None

    def __getattr__(self, name):
        return getattr(self.loader, name)

    def __len__(self):
        return len(self.loader)

    def _to_device(self, obj):
        if torch.is_tensor(obj):
            if self._stream is not None and not obj.is_pinned():
                obj = obj.pin_memory()
            return obj.to(self.device, non_blocking=True)
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._to_device(o) for o in obj)
        if isinstance(obj, dict):
            return {k: self._to_device(v) for k, v in obj.items()}
        return obj

    def _record(self, obj, stream):
        if torch.is_tensor(obj):
            obj.record_stream(stream)
        elif isinstance(obj, (list, tuple)):
            for o in obj:
                self._record(o, stream)
        elif isinstance(obj, dict):
            for o in obj.values():
                self._record(o, stream)

    def __iter__(self):
        if self._stream is None:
            for batch in self.loader:
                yield self._to_device(batch)
            return
        it = iter(self.loader)
        # NOTE: prefetching one batch ahead would advance the elastic loader's
        # profile()/current_index one step early; so only the H2D copy is
        # overlapped (issued on the side stream, waited by the consumer).
        for batch in it:
            with torch.cuda.stream(self._stream):
                dev = self._to_device(batch)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_stream(self._stream)
            self._record(dev, cur)
            yield dev


def _reset_for_tests():
    AdaptiveDataLoaderHelper._position = collections.Counter()
    AdaptiveDataLoaderHelper._training = None
    AdaptiveDataLoaderHelper._current = None
    _AdaptiveDataLoaderState.init_count = collections.Counter()
    _PREEMPTION.reset()
