import numpy as np


def crossover_mask(X, M):
    """Two parents [2, n_matings, ...]: the children swap the entries where
    the mask is set."""
    out = np.copy(X)
    out[0][M] = X[1][M]
    out[1][M] = X[0][M]
    return out
