"""Metric accumulation that is consistent across replicas and restarts.

``Accumulator`` is the user-facing object of the reference's
``adaptdl/adaptdl/torch/accumulator.py`` (same constructor, ``+=`` / ``-=``
/ ``update`` / ``subtract`` / ``synchronized()`` behaviour, same checkpoint
file contents); the machinery below is this repo's own:

* local updates are an append-only **delta log** of ``(key, amount)``
  entries -- ``accum[k] += v`` builds an immutable :class:`_Delta` through
  the ``__getitem__`` / ``__add__`` / ``__setitem__`` protocol and logs it;
* a **ledger** (the checkpointed :class:`~adaptdl_b200.checkpoint.State`)
  holds the job-wide totals and the snapshots that must be replayed after a
  restart, indexed by *loop position* ``(epoch, n-th synchronisation of that
  epoch outside data-loader loops)``;
* ``synchronized()`` opens a **session**: the compacted delta logs of all
  replicas are summed over the control plane into the totals (or, if this
  loop position was already passed before the last restart, the recorded
  snapshot is served instead and the local log is dropped), and while the
  session is open the accumulator is an ordinary ``dict`` view of the totals.

Outside a session reads see an empty mapping: local partial sums are never
observable, which is what makes results independent of the replica count.
"""

import collections
import collections.abc
import copy
import pickle

from adaptdl_b200 import checkpoint, collective
from adaptdl_b200.torch.data import current_dataloader
from adaptdl_b200.torch.epoch import current_epoch

__all__ = ["Accumulator"]


def _merge_sums(total, part):
    """``total[k] += part[k]`` (keys missing from ``total`` start at the
    addend itself, so any type with ``+`` works)."""
    for key, amount in part.items():
        total[key] = (total[key] + amount) if key in total else amount
    return total


class _Delta(object):
    """The pending effect of ``accum[key] <op>= amount``. Immutable: every
    ``+`` / ``-`` gives a new delta, so a stray reference can never change
    what was logged."""

    __slots__ = ("owner", "key", "amount")

    def __init__(self, owner, key, amount=0):
        object.__setattr__(self, "owner", owner)
        object.__setattr__(self, "key", key)
        object.__setattr__(self, "amount", amount)

    def __setattr__(self, name, value):
        raise AttributeError("_Delta is immutable")

    def _shifted(self, amount, sign):
        if isinstance(amount, _Delta):
            raise TypeError("an accumulator entry cannot be added to "
                            "another one; add plain values")
        return _Delta(self.owner, self.key,
                      self.amount + amount if sign > 0
                      else self.amount - amount)

    def __add__(self, amount):
        return self._shifted(amount, +1)

    __radd__ = __add__

    def __sub__(self, amount):
        return self._shifted(amount, -1)


class _Ledger(checkpoint.State):
    """Checkpointed part of an accumulator: the replica-independent totals
    and the snapshots to replay, keyed by loop position."""

    _created = collections.Counter()      # epoch -> ledgers created in it

    def __init__(self, initial):
        if current_dataloader() is not None:
            raise RuntimeError(
                "create Accumulators before entering a data-loader loop: "
                "all replicas must create the same ones in the same order")
        epoch = current_epoch()
        index = _Ledger._created[epoch]
        _Ledger._created[epoch] += 1
        # the name doubles as the file name inside a checkpoint
        super().__init__("adaptdl-accumulator-epoch{}-{}".format(epoch,
                                                                 index))
        self.totals = dict(initial)
        self.snapshots = {}               # (epoch, ordinal) -> dict
        self.log = []                     # local (key, amount) entries

    # -- local log ---------------------------------------------------------

    def record(self, key, amount):
        self.log.append((key, amount))

    def compact(self):
        """Collapse the log into one amount per key (and empty it)."""
        sums = {}
        for key, amount in self.log:
            sums[key] = (sums[key] + amount) if key in sums else amount
        self.log = []
        return sums

    def settle(self):
        """Sum every replica's log into the totals (a collective)."""
        combined = collective.allreduce(self.compact(), _merge_sums)
        _merge_sums(self.totals, combined)

    # -- checkpoint.State --------------------------------------------------

    def sync(self):
        self.settle()

    def save(self, fileobj):
        # on-disk layout of the reference: one pickle of
        # (history: {epoch: [snapshot, ...]}, results)
        history = collections.defaultdict(list)
        for (epoch, ordinal) in sorted(
                self.snapshots, key=lambda k: (k[0] is not None, k[0] or 0,
                                               k[1])):
            history[epoch].append(self.snapshots[(epoch, ordinal)])
        pickle.dump((history, self.totals), fileobj)

    def load(self, fileobj):
        history, self.totals = pickle.load(fileobj)
        self.snapshots = {}
        for epoch, rows in history.items():
            for ordinal, snap in enumerate(rows):
                self.snapshots[(epoch, ordinal)] = snap

    def forget_before(self, epoch):
        """Snapshots of finished epochs can never be asked for again."""
        if epoch is None:
            return
        for position in [p for p in self.snapshots
                         if p[0] is not None and p[0] < epoch]:
            del self.snapshots[position]


class _Session(object):
    """``with accum.synchronized():`` -- re-entrant; the outermost entry
    decides what the accumulator shows."""

    def __init__(self, accum):
        self.accum = accum

    def __enter__(self):
        accum = self.accum
        accum._depth += 1
        if accum._depth > 1:
            return accum
        ledger = accum._ledger
        if current_dataloader() is not None:
            # inside a data-loader loop: such code is skipped, never
            # re-executed, after a restart -- nothing to replay or record
            ledger.settle()
            accum._visible = ledger.totals
            return accum
        epoch = current_epoch()
        ledger.forget_before(epoch)
        position = (epoch, accum._visits[epoch])   # n-th out-of-loop session
        accum._visits[epoch] += 1
        if position in ledger.snapshots:
            # this loop position was passed before the last restart: show
            # what was shown then; whatever was logged on the way here is a
            # repetition of work already counted
            ledger.log = []
            accum._visible = ledger.snapshots[position]
        else:
            ledger.settle()
            ledger.snapshots[position] = copy.deepcopy(ledger.totals)
            accum._visible = ledger.totals
        return accum

    def __exit__(self, *exc):
        accum = self.accum
        accum._depth -= 1
        if accum._depth == 0:
            accum._visible = None
        return False


class Accumulator(collections.abc.MutableMapping):
    """A ``dict``-like metric accumulator. Example::

        accum = Accumulator()
        for epoch in remaining_epochs_until(60):
            for batch in validloader:
                accum["loss_sum"] += loss_sum
                accum["total"] += len(batch)
            with accum.synchronized():
                print(accum["loss_sum"] / accum["total"])
                accum.clear()

    In *accumulation mode* (the default) only additive updates are allowed
    (``+=``, ``-=``, :meth:`update`, :meth:`subtract`); they stay local
    until the next ``synchronized()`` block, inside which every replica sees
    the same job-wide sums and may treat the object as a plain ``dict``
    (writes there must be identical on every replica).

    Arguments: same as ``dict``.
    """

    def __init__(self, *args, **kwargs):
        self._visits = collections.Counter()   # epoch -> sessions opened
        self._depth = 0
        self._visible = None                   # dict shown inside a session
        self._ledger = _Ledger(dict(*args, **kwargs))
        checkpoint.load_state(self._ledger)

    def synchronized(self):
        """Context manager entering synchronized mode. A collective: every
        replica must open the session at the same point of the program."""
        return _Session(self)

    # -- additive updates ------------------------------------------------------

    def update(self, *args, **kwargs):
        """Add the given amounts key by key (NOT ``dict.update``)."""
        for key, amount in dict(*args, **kwargs).items():
            self[key] += amount

    def subtract(self, *args, **kwargs):
        """Subtract the given amounts key by key."""
        for key, amount in dict(*args, **kwargs).items():
            self[key] -= amount

    def __iadd__(self, other):
        self.update(other)
        return self

    def __isub__(self, other):
        self.subtract(other)
        return self

    # -- mapping protocol ------------------------------------------------------

    def __getitem__(self, key):
        if self._visible is not None:
            return self._visible[key]
        # accumulation mode: ``accum[key] += v`` evaluates to
        # ``accum[key] = accum[key] + v``; hand out a delta to carry ``v``
        return _Delta(self, key)

    def __setitem__(self, key, value):
        if self._visible is not None:
            self._visible[key] = value
            return
        if not isinstance(value, _Delta):
            raise TypeError(
                "outside synchronized() an Accumulator only takes additive "
                "updates (accum[k] += v), not assignment of {}".format(
                    type(value).__name__))
        if value.owner is not self or value.key != key:
            raise ValueError("the update was built from a different "
                             "accumulator entry ({!r})".format(value.key))
        self._ledger.record(key, value.amount)

    def _shown(self):
        return self._visible if self._visible is not None else {}

    def __delitem__(self, key):
        del self._shown()[key]

    def __contains__(self, key):
        return key in self._shown()

    def __iter__(self):
        return iter(self._shown())

    def __len__(self):
        return len(self._shown())

    def __repr__(self):
        return "Accumulator({!r})".format(self._shown())


def _reset_for_tests():
    _Ledger._created = collections.Counter()
