"""Step profiler fed by on-device timestamps.

The reference times every training iteration on the host
(``adaptdl/adaptdl/torch/_metrics.py:43-59``: ``time.time()`` around the loop
body) and the gradient synchronisation with a blocking CUDA event
(``adaptdl/adaptdl/torch/parallel.py:103-146``). With the device engine the
host never waits for the GPU inside a step -- it may run several steps ahead
of it -- so host clocks measure launch rate, not step time. Here the finalize
of every optimizer step (``csrc/adl_kernels.cu``, ``finalize_body``) stamps
``%globaltimer`` and publishes, through the statistics mailbox:

* ``step_ns``   the interval since the previous step mark, max over ranks;
* ``sync_ns``   end of the local backward -> gradients reduced and statistics
  exchanged, max over ranks;
* ``accum_ns`` / ``accum_count``  the accumulation micro-steps since the
  previous optimizer step (``adl_step_mark``).

:class:`DeviceStepTimer` pairs those records with the profile keys
``(num_nodes, num_replicas, atomic_bsz)`` of the iterations that produced
them and hands completed ones to ``_metrics`` -- asynchronously: a record is
booked when the device has published it, normally one or two iterations
after the host issued the step.
"""

import collections

import numpy as np

from adaptdl_b200._native import (MB_SEQ, MB_SYNC_NS, MB_STEP_NS,
                                  MB_ACCUM_NS, MB_ACCUM_COUNT)

__all__ = ["DeviceStepTimer", "StepRecord"]

StepRecord = collections.namedtuple(
    "StepRecord", ["key", "step_time", "sync_time", "accum_time",
                   "accum_count"])


class DeviceStepTimer(object):

    def __init__(self, reducer):
        self.reducer = reducer
        self._pending = collections.OrderedDict()   # step index -> key
        self._noted = reducer._steps - 1
        self.booked = 0
        self.dropped = 0

    def active(self):
        engine = getattr(self.reducer, "engine", None)
        return engine is not None and engine.enabled

    def reset(self):
        """The next step's interval is not a training iteration (start of a
        loop, evaluation or a checkpoint in between): do not measure it."""
        self.reducer.reset_step_clock()

    def note(self, key):
        """The iteration that just ended issued optimizer step(s); ``key`` is
        its profile key, or ``None`` if it must not be booked (warm-up)."""
        latest = self.reducer._steps - 1
        for step in range(self._noted + 1, latest + 1):
            self._pending[step] = key
        self._noted = max(self._noted, latest)

    def drain(self, wait=False):
        """Records the device has published so far, oldest first."""
        out = []
        red = self.reducer
        if wait and self._pending:
            import torch
            torch.cuda.synchronize(red.device)
        while self._pending:
            step, key = next(iter(self._pending.items()))
            arr = red.peek_slot(step)
            if arr is None:
                break                       # not published yet
            del self._pending[step]
            if arr is False:                # overwritten: the host fell
                self.dropped += 1           # a whole ring behind
                continue
            hdr = np.array(arr[:MB_ACCUM_COUNT + 1])
            if int(hdr[MB_SEQ]) != step + 1:
                self.dropped += 1
                continue
            step_s = float(hdr[MB_STEP_NS]) * 1e-9
            if key is None or step_s <= 0.0:
                continue                    # warm-up / right after a reset
            self.booked += 1
            out.append(StepRecord(key, step_s,
                                  float(hdr[MB_SYNC_NS]) * 1e-9,
                                  float(hdr[MB_ACCUM_NS]) * 1e-9,
                                  int(hdr[MB_ACCUM_COUNT])))
        return out
