"""Graceful-preemption signal flag.

SIGTERM / SIGINT set a process-global flag; the data loader OR-reduces it
across replicas every step so that all replicas checkpoint at the same
iteration and exit with code 143 (parity: reference ``_signal.py:25-42``).
A second SIGINT falls through to the previous handler (force quit).
"""

import logging
import signal
import threading

LOG = logging.getLogger(__name__)

_EXIT_FLAG = False
_PREV_SIGINT = signal.getsignal(signal.SIGINT)


def get_exit_flag():
    return _EXIT_FLAG


def set_exit_flag(value=True):
    """Programmatic preemption request (used by launchers and tests)."""
    global _EXIT_FLAG
    _EXIT_FLAG = bool(value)


def _on_signal(signum, frame):
    set_exit_flag(True)
    try:
        from adaptdl_b200.utils import rescale_trace
        rescale_trace.mark("signal_received")
    except Exception:  # noqa: BLE001 - never fail inside a signal handler
        pass
    if signum == signal.SIGINT:
        LOG.info("SIGINT: finishing this step then checkpointing; "
                 "send it again to force exit")
        signal.signal(signal.SIGINT, _PREV_SIGINT)


def install():
    """Install the handlers (only possible from the main thread)."""
    if threading.current_thread() is threading.main_thread():
        signal.signal(signal.SIGTERM, _on_signal)
        signal.signal(signal.SIGINT, _on_signal)


install()
