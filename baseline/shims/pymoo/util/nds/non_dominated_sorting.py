import numpy as np


def fronts_of(F):
    n = len(F)
    le = np.all(F[:, None, :] <= F[None, :, :], axis=2)
    lt = np.any(F[:, None, :] < F[None, :, :], axis=2)
    dominates = le & lt
    counts = dominates.sum(axis=0)
    left = np.ones(n, dtype=bool)
    fronts = []
    while left.any():
        front = np.flatnonzero(left & (counts == 0))
        if front.size == 0:
            front = np.flatnonzero(left)
        fronts.append(front)
        left[front] = False
        counts = counts - dominates[front].sum(axis=0)
    return fronts


class NonDominatedSorting(object):
    def do(self, F, only_non_dominated_front=False, **kwargs):
        fronts = fronts_of(np.asarray(F, dtype=float))
        return fronts[0] if only_non_dominated_front else fronts
