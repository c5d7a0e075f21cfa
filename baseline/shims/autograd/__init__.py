"""Stand-in for the HIPS `autograd` package (not installable offline). The
reference differentiates its perf-model fitting objective with it
(adaptdl/goodput.py); a central finite difference is enough for the baseline
(the fit runs on rank 0 every >= 30 s, off the measured path)."""
import numpy as _np

from . import numpy  # noqa: F401


def grad(fn):
    def gradient(x, *args):
        x = _np.asarray(x, dtype=float)
        out = _np.zeros_like(x)
        for i in range(x.size):
            h = 1e-6 * max(1.0, abs(x[i]))
            e = _np.zeros_like(x)
            e[i] = h
            out[i] = (fn(x + e, *args) - fn(x - e, *args)) / (2 * h)
        return out
    return gradient
