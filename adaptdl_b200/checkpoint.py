"""Checkpoint-restart state registry.

Elasticity is checkpoint -> exit(143) -> restart at a new world size; every
piece of state that must survive is a named :class:`State`. On-disk layout is
the reference's (``adaptdl/adaptdl/checkpoint.py:33-206``)::

    $ADAPTDL_CHECKPOINT_PATH/
        _checkpoint/               staging dir while rank 0 writes
        checkpoint-<num_restarts>/ exactly one survives a successful save
            <state name>           one file per State

Rank 0 writes, every rank reads, so the path must be shared storage. The
staging directory is renamed into place atomically and older generations are
removed only afterwards, so a crash mid-save leaves the previous checkpoint
intact.
"""

import logging
import os
import pickle
import shutil
import weakref

from adaptdl_b200 import env
from adaptdl_b200.utils.trace import traced

LOG = logging.getLogger(__name__)

CKPT_DIR_PREFIX = "checkpoint-"
_STAGING = "_checkpoint"


class _Registry:
    """name <-> State, insertion ordered (save order = creation order)."""

    def __init__(self):
        self.by_name = {}

    def add(self, name, state):
        if name in self.by_name:
            raise ValueError("State '{}' already exists".format(name))
        self.by_name[name] = state

    def name_of(self, state):
        return state._name

    def states(self):
        return list(self.by_name.values())

    def discard(self, state):
        if self.by_name.get(state._name) is state:
            del self.by_name[state._name]

    def clear(self):
        self.by_name.clear()


_REGISTRY = _Registry()


class State(object):
    """A named piece of state which can be saved/loaded as part of a
    checkpoint and synchronised across replicas before saving. Subclass and
    override :meth:`save`, :meth:`load`, and optionally :meth:`sync`.

    Arguments:
        name (str): unique name; also the file name inside the checkpoint.

    Raises:
        ValueError: if another live State already has this name.
    """

    def __init__(self, name):
        self._name = str(name)
        _REGISTRY.add(self._name, self)

    @property
    def name(self):
        return self._name

    def save(self, fileobj):
        """Write this state to a binary file object (rank 0 only)."""

    def load(self, fileobj):
        """Restore this state from a binary file object (all ranks)."""

    def sync(self):
        """Make the state consistent across replicas (called on ALL ranks
        before :meth:`save`)."""

    def unregister(self):
        """Remove from the registry (extension: lets long-lived processes
        and tests drop states; the reference leaks them)."""
        _REGISTRY.discard(self)


class PickledFields(State):
    """A :class:`State` whose payload is a fixed list of attributes, pickled
    in one of the three layouts the checkpoint format uses:

    ``"value"``     a single attribute, pickled bare;
    ``"tuple"``     one pickle holding the tuple of attributes;
    ``"sequence"``  one pickle per attribute, back to back.
    """

    FIELDS = ()
    LAYOUT = "tuple"

    def save(self, fileobj):
        values = tuple(getattr(self, name) for name in self.FIELDS)
        if self.LAYOUT == "value":
            pickle.dump(values[0], fileobj)
        elif self.LAYOUT == "tuple":
            pickle.dump(values, fileobj)
        else:
            for value in values:
                pickle.dump(value, fileobj)

    def load(self, fileobj):
        if self.LAYOUT == "value":
            values = (pickle.load(fileobj),)
        elif self.LAYOUT == "tuple":
            values = pickle.load(fileobj)
        else:
            values = [pickle.load(fileobj) for _ in self.FIELDS]
        for name, value in zip(self.FIELDS, values):
            setattr(self, name, value)


_REPLACED = ".replaced"
# Every state file is forced to stable storage before the checkpoint is
# published: the atomic rename protects against a dying process, fsync against
# the NODE dying within seconds of a checkpoint (rename before data reaches
# the disk can leave empty files behind). It costs the device's write
# bandwidth -- 0.6 s per 1.3 GB on this container's disk, several seconds on
# a slow network volume -- on the path where a preempted job races its
# termination grace period; ``ADAPTDL_CHECKPOINT_FSYNC=0`` trades that
# protection for the time (the reference never syncs).


def _fsync_wanted():
    return os.environ.get("ADAPTDL_CHECKPOINT_FSYNC", "1") != "0"


def _staging_dir(root):
    path = os.path.join(root, _STAGING)
    os.makedirs(path, exist_ok=True)
    return path


def _generation_dirs(root):
    out = {}
    try:
        names = os.listdir(root)
    except FileNotFoundError:
        return out
    replaced = {}
    for name in names:
        if name.startswith(CKPT_DIR_PREFIX):
            stem = name[len(CKPT_DIR_PREFIX):]
            target = out
            if stem.endswith(_REPLACED):     # see save_all_states
                stem, target = stem[:-len(_REPLACED)], replaced
            try:
                target[int(stem)] = os.path.join(root, name)
            except ValueError:
                continue
    # a generation whose re-save died between its two renames survives as
    # its previous copy
    for gen, path in replaced.items():
        out.setdefault(gen, path)
    return out


def _resolve_checkpoint_dir(for_save):
    if env.from_ray():
        try:
            from adaptdl_b200.ray import tune_checkpoint_dir
            return tune_checkpoint_dir(for_save)
        except ImportError:
            pass
    return env.checkpoint_path()


@traced("checkpoint_save")
def save_all_states():
    """Checkpoint every registered :class:`State`: ``sync()`` on all
    replicas, ``save()`` on rank 0 into a staging dir which is then renamed
    to ``checkpoint-<num_restarts>``; older generations are deleted.

    Returns the checkpoint root on rank 0 (``None`` elsewhere).
    """
    root = _resolve_checkpoint_dir(for_save=True)
    for state in _REGISTRY.states():
        save_state(state, root)
    if env.replica_rank() != 0 or root is None:
        return None
    final = os.path.join(root, CKPT_DIR_PREFIX + str(env.num_restarts()))
    staging = _staging_dir(root)
    if os.path.isdir(final):       # re-save within the same generation:
        # move the old copy aside instead of deleting it first, so that a
        # crash at any point leaves one complete checkpoint on disk
        aside = final + _REPLACED
        shutil.rmtree(aside, ignore_errors=True)
        os.rename(final, aside)
    os.rename(staging, final)      # atomic publish
    for path in _generation_dirs(root).values():
        if path != final:
            shutil.rmtree(path, ignore_errors=True)
    shutil.rmtree(final + _REPLACED, ignore_errors=True)
    return root


def save_state(state, checkpoint_dir, sync=True):
    """Save one state into the staging dir of ``checkpoint_dir``.

    ``state.sync()`` runs on every replica (it may contain collectives),
    ``state.save()`` only on rank 0.
    """
    if sync:
        state.sync()
    if env.replica_rank() == 0 and checkpoint_dir is not None:
        path = os.path.join(_staging_dir(checkpoint_dir),
                            _REGISTRY.name_of(state))
        with open(path, "wb") as f:
            state.save(f)
            f.flush()
            if _fsync_wanted():
                os.fsync(f.fileno())


def latest_checkpoint_dir(root=None):
    """Path of the newest ``checkpoint-N`` directory or ``None``."""
    root = root or _resolve_checkpoint_dir(for_save=False)
    if root is None:
        return None
    gens = _generation_dirs(root)
    if not gens:
        return None
    newest = max(gens)
    if newest != env.num_restarts() - 1:
        LOG.warning("no checkpoint from the previous restart; loading "
                    "generation %d", newest)
    return gens[newest]


def load_state(state):
    """Load ``state`` from the newest checkpoint, if it was saved there.

    Returns ``True`` iff ``State.load`` was invoked.
    """
    ckpt = latest_checkpoint_dir()
    if ckpt is None:
        return False
    path = os.path.join(ckpt, _REGISTRY.name_of(state))
    if not os.path.isfile(path):
        # normal for states created after the checkpoint was written (e.g. a
        # new epoch's Accumulator): they simply start fresh
        LOG.debug("state file %s not found", path)
        return False
    with open(path, "rb") as f:
        state.load(f)
    return True


def _reset_registry_for_tests():
    _REGISTRY.clear()


# Back-compat aliases for code that pokes at the reference's module globals.
_NAMES_TO_STATES = _REGISTRY.by_name
_ = weakref  # (reserved: weak registry mode)
