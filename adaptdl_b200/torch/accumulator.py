"""Replica- and restart-consistent metric accumulation.

An :class:`Accumulator` looks like a ``dict`` with two modes:

* **accumulation mode** (default): ``accum[k] += v`` / ``accum.update(k=v)``
  record *local* additive updates, lazily summed over replicas; reads behave
  like an empty dict.
* **synchronized mode** (``with accum.synchronized():``): pending updates of
  all replicas are summed in, every replica sees the same contents and may
  use the object like a normal ``dict`` (writes must be identical everywhere).

Results of synchronisations that happened outside of data-loader loops are
remembered per epoch and *replayed* after a restart, so code that already ran
before the checkpoint observes the same values when the script is re-executed
(parity: reference ``torch/accumulator.py:27-312``).
"""

import collections
import collections.abc
import contextlib
import copy

from adaptdl_b200 import checkpoint, collective
from adaptdl_b200.torch.data import current_dataloader
from adaptdl_b200.torch.epoch import current_epoch

__all__ = ["Accumulator"]


def _dict_iadd(a, b):
    for k, v in b.items():
        a[k] = a[k] + v if k in a else v
    return a


class _AccumulatorState(checkpoint.PickledFields):
    """What an :class:`Accumulator` persists: the totals of the current epoch
    and, per epoch, the history of synchronised snapshots (replayed after a
    restart). Pending local updates are not persisted -- they are summed into
    the totals before every save."""

    FIELDS = ("results_history", "results")
    LAYOUT = "tuple"
    init_count = collections.Counter()   # epoch -> accumulators created

    def __init__(self, *args, **kwargs):
        if current_dataloader() is not None:
            raise RuntimeError("an Accumulator must be created outside of "
                               "data-loader loops (every replica has to "
                               "create the same accumulators in the same "
                               "order)")
        epoch = current_epoch()
        counter = _AccumulatorState.init_count
        super().__init__("adaptdl-accumulator-epoch{}-{}".format(
            epoch, counter[epoch]))
        counter[epoch] += 1
        self.updates = {}
        self.results = dict(*args, **kwargs)
        self.results_history = collections.defaultdict(list)

    def sync(self):
        """Sum the pending updates of all replicas into ``results``."""
        updates = collective.allreduce(self.updates, _dict_iadd)
        _dict_iadd(self.results, updates)
        self.updates.clear()


class _Pending(object):
    """What ``accum[key]`` returns in accumulation mode: captures the
    ``+ v`` / ``- v`` of an in-place update so ``__setitem__`` can record
    it."""

    __slots__ = ("accum", "key", "update")

    def __init__(self, accum, key):
        self.accum = accum
        self.key = key
        self.update = 0

    def __add__(self, update):
        if isinstance(update, _Pending):
            raise TypeError("invalid update type: {}".format(type(update)))
        self.update += update
        return self

    def __sub__(self, update):
        if isinstance(update, _Pending):
            raise TypeError("invalid update type: {}".format(type(update)))
        self.update -= update
        return self


_Value = _Pending   # reference name


class Accumulator(collections.abc.MutableMapping):
    """See the module docstring. Example::

        accum = Accumulator()
        for epoch in remaining_epochs_until(60):
            for batch in validloader:
                accum["loss_sum"] += loss_sum
                accum["total"] += len(batch)
            with accum.synchronized():
                print(accum["loss_sum"] / accum["total"])
                accum.clear()

    Arguments: same as ``dict``.
    """

    def __init__(self, *args, **kwargs):
        self._sync_count = collections.Counter()
        self._synchronized = None
        self._state = _AccumulatorState(*args, **kwargs)
        checkpoint.load_state(self._state)

    @contextlib.contextmanager
    def synchronized(self):
        """Enter synchronized mode. A distributed synchronisation point: all
        replicas must enter it at the same place."""
        if self._synchronized is not None:      # re-entrant
            yield self
            return
        epoch = current_epoch()
        history = self._state.results_history
        for key in list(history.keys()):        # finished epochs never replay
            if key is not None and epoch is not None and key < epoch:
                history.pop(key)
        ordinal = self._sync_count[epoch]
        self._sync_count[epoch] += 1
        saved = history[epoch]
        assert ordinal <= len(saved)
        if ordinal < len(saved):
            # this synchronisation already happened before the restart
            self._synchronized = saved[ordinal]
            self._state.updates.clear()
        else:
            self._state.sync()
            if current_dataloader() is None:
                # code inside loader loops is not replayed, so only
                # out-of-loop results need remembering
                saved.append(copy.deepcopy(self._state.results))
            self._synchronized = self._state.results
        try:
            yield self
        finally:
            self._synchronized = None

    def update(self, *args, **kwargs):
        """*Additively* apply key/update pairs (unlike ``dict.update``)."""
        for key, val in dict(*args, **kwargs).items():
            self[key] += val

    def subtract(self, *args, **kwargs):
        """Subtract key/update pairs."""
        for key, val in dict(*args, **kwargs).items():
            self[key] -= val

    def __iadd__(self, other):
        """``accum += {k: v}`` == ``accum.update({k: v})``."""
        self.update(other)
        return self

    def __isub__(self, other):
        """``accum -= {k: v}`` == ``accum.subtract({k: v})``."""
        self.subtract(other)
        return self

    def __getitem__(self, key):
        """Read access is meaningful in synchronized mode only; in
        accumulation mode this supports ``accum[key] += v``."""
        if self._synchronized is not None:
            return self._synchronized[key]
        return _Pending(self, key)

    def __setitem__(self, key, value):
        if self._synchronized is not None:
            self._synchronized[key] = value
            return
        # ``a[k] += v`` is  tmp = a[k]; tmp += v; a[k] = tmp  -- tmp is the
        # _Pending returned by __getitem__, carrying v.
        if not isinstance(value, _Pending):
            raise TypeError("invalid value type: {}".format(type(value)))
        if value.accum is not self:
            raise ValueError("incompatible {}".format(type(self).__name__))
        if key != value.key:
            raise ValueError("incompatible key: {}".format(value.key))
        updates = self._state.updates
        updates[key] = updates.get(key, 0) + value.update

    def _view(self):
        return self._synchronized if self._synchronized is not None else {}

    def __contains__(self, key):
        return key in self._view()

    def __delitem__(self, key):
        del self._view()[key]

    def __iter__(self):
        return iter(self._view())

    def __len__(self):
        return len(self._view())

    def __repr__(self):
        return repr(self._view())


def _reset_for_tests():
    _AccumulatorState.init_count = collections.Counter()
