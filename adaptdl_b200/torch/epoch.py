"""Restart-safe epoch loops.

Elastic jobs restart by re-executing the training script from the top, so
epoch loops must *skip* what already finished. ``remaining_epochs_until(n)``
yields only the epochs that still have to run; the count of finished epochs
is part of every checkpoint (parity: reference ``torch/epoch.py:96-178``).

The contract for user code is idempotency: everything executed between
restarts outside of epoch/dataloader loops must be safe to replay::

    for epoch in remaining_epochs_until(30):     # epochs 0..29
        for batch in loader: ...
    for epoch in remaining_epochs_until(60):     # epochs 30..59
        ...
"""

import logging

from adaptdl_b200 import checkpoint

LOG = logging.getLogger(__name__)

__all__ = ["remaining_epochs_until", "current_epoch", "finished_epochs"]


class _EpochState(checkpoint.PickledFields):
    """Epochs completed so far (checkpointed) and the epoch of the loop that
    is running now (not checkpointed: a restart re-enters the loop)."""

    FIELDS = ("finished_epochs",)
    LAYOUT = "value"

    def __init__(self):
        super().__init__(".adaptdl-epoch")
        self.current_epoch = None
        self.finished_epochs = 0


_EPOCH_STATE = None


def _epoch_state():
    global _EPOCH_STATE
    if _EPOCH_STATE is None:
        _EPOCH_STATE = _EpochState()
        checkpoint.load_state(_EPOCH_STATE)
    return _EPOCH_STATE


def remaining_epochs_until(epoch):
    """Iterate over the epochs in ``[finished_epochs(), epoch)``.

    Raises:
        RuntimeError: when nested inside another epoch loop.
    """
    state = _epoch_state()
    if state.current_epoch is not None:
        raise RuntimeError("overlapping epoch loops detected")
    if state.finished_epochs < epoch:
        LOG.info("starting at epoch %s", state.finished_epochs)
    else:
        LOG.info("skipping all epochs up to %s", epoch)
    while state.finished_epochs < epoch:
        state.current_epoch = state.finished_epochs
        try:
            yield state.current_epoch
        finally:
            # Runs on normal completion, ``break`` and exceptions alike.
            state.finished_epochs += 1
            state.current_epoch = None


def current_epoch():
    """Epoch being iterated by :func:`remaining_epochs_until`, else None."""
    return _epoch_state().current_epoch


def finished_epochs():
    """Number of epochs completed so far (across restarts)."""
    return _epoch_state().finished_epochs


def _reset_for_tests():
    global _EPOCH_STATE
    _EPOCH_STATE = None
