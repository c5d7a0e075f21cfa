"""Stand-in for the HIPS `autograd` package (not installable offline). The
reference differentiates its perf-model fitting objective with it
(adaptdl/goodput.py); a finite difference is enough for the baseline (the fit runs on rank 0
every >= 30 s, off the measured path). FORWARD differences: every parameter
of that objective has a lower bound (1e-8, or 1 for gamma) that the optimiser
sits on for long stretches, and a central difference would evaluate the
objective below it (log / power of negative numbers -> NaN gradients)."""
import numpy as _np

from . import numpy  # noqa: F401


def grad(fn):
    def gradient(x, *args):
        x = _np.asarray(x, dtype=float)
        out = _np.zeros_like(x)
        base = fn(x, *args)
        for i in range(x.size):
            h = 1e-7 * max(1.0, abs(x[i]))
            e = _np.zeros_like(x)
            e[i] = h
            out[i] = (fn(x + e, *args) - base) / h
        return out
    return gradient
