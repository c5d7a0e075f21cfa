from numpy import *  # noqa: F401,F403
import numpy as _np

inf = _np.inf
