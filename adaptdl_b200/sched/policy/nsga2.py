"""A small NSGA-II (Deb et al. 2002) for integer-matrix genomes.

The reference leans on ``pymoo`` for this; the algorithm is simple enough to
own: fast non-dominated sorting, crowding distance, binary tournament on
(rank, crowding), elitist (mu + lambda) survival, duplicate elimination. The
problem object supplies the domain-specific operators:

    problem.evaluate(X)        -> F   [pop, n_obj]   (minimised)
    problem.crossover(A, B)    -> children (same shape as concat(A, B))
    problem.mutate(X)          -> X'
    problem.repair(X)          -> feasible X
"""

import numpy as np


def non_dominated_fronts(F):
    """List of index arrays: front 0 is the Pareto set of ``F`` (rows are
    points, all objectives minimised), front 1 the Pareto set of the rest,
    and so on."""
    n = F.shape[0]
    le = np.all(F[:, None, :] <= F[None, :, :], axis=2)
    lt = np.any(F[:, None, :] < F[None, :, :], axis=2)
    dominates = le & lt                       # [i, j]: i dominates j
    dominated_by = dominates.sum(axis=0)      # how many dominate j
    fronts = []
    remaining = np.ones(n, dtype=bool)
    counts = dominated_by.copy()
    while remaining.any():
        front = np.flatnonzero(remaining & (counts == 0))
        if front.size == 0:                   # numerical safety
            front = np.flatnonzero(remaining)
        fronts.append(front)
        remaining[front] = False
        counts = counts - dominates[front].sum(axis=0)
    return fronts


def crowding_distance(F):
    """Crowding distance of every row of ``F`` (within one front)."""
    n, m = F.shape
    dist = np.zeros(n)
    if n <= 2:
        dist[:] = np.inf
        return dist
    for k in range(m):
        order = np.argsort(F[:, k], kind="stable")
        col = F[order, k]
        span = col[-1] - col[0]
        dist[order[0]] = dist[order[-1]] = np.inf
        if span > 0:
            dist[order[1:-1]] += (col[2:] - col[:-2]) / span
    return dist


def _rank_and_crowding(F):
    rank = np.zeros(F.shape[0], dtype=int)
    crowd = np.zeros(F.shape[0])
    for r, front in enumerate(non_dominated_fronts(F)):
        rank[front] = r
        crowd[front] = crowding_distance(F[front])
    return rank, crowd


_HASH_SEED = np.random.default_rng(0x5EED)


def _row_keys(X):
    """One 64-bit key per row: a random linear form over the integers mod
    2**64 (rows are ~100 KB at cluster scale, so hashing their bytes in
    Python dominated the generation loop). Different rows collide with
    probability ~2**-64."""
    global _HASH_VECTOR
    width = X.shape[1]
    if _HASH_VECTOR is None or len(_HASH_VECTOR) < width:
        _HASH_VECTOR = _HASH_SEED.integers(
            1, np.iinfo(np.int64).max, size=max(width, 1024), dtype=np.int64)
    with np.errstate(over="ignore"):
        return X.astype(np.int64, copy=False) @ _HASH_VECTOR[:width]


_HASH_VECTOR = None


def _unique_rows(X, against=None):
    """Boolean mask of rows of ``X`` that are new (not in ``against`` and
    not repeated earlier in ``X``)."""
    seen = set(_row_keys(against).tolist()) if against is not None \
        and len(against) else set()
    keep = np.zeros(len(X), dtype=bool)
    for i, key in enumerate(_row_keys(X).tolist()):
        if key not in seen:
            seen.add(key)
            keep[i] = True
    return keep


def minimize(problem, initial, pop_size=100, n_gen=100, rng=None):
    """Run NSGA-II from the ``initial`` population (any number of genomes,
    shape ``[k, ...]``). Returns ``(X, F)`` of the final population."""
    rng = np.random.default_rng() if rng is None else rng
    shape = initial.shape[1:]
    X = problem.repair(np.array(initial, copy=True))
    X = X.reshape(len(X), -1)
    X = X[_unique_rows(X)]
    F = problem.evaluate(X.reshape(-1, *shape))
    for _ in range(n_gen):
        rank, crowd = _rank_and_crowding(F)
        n = len(X)
        # binary tournaments -> 2 parents per mating
        n_matings = (pop_size + 1) // 2

        def tournament(count):
            a = rng.integers(n, size=count)
            b = rng.integers(n, size=count)
            better = (rank[a] < rank[b]) | (
                (rank[a] == rank[b]) & (crowd[a] >= crowd[b]))
            return np.where(better, a, b)
        pa, pb = tournament(n_matings), tournament(n_matings)
        children = problem.crossover(X[pa].reshape(-1, *shape),
                                     X[pb].reshape(-1, *shape), rng)
        children = problem.mutate(children, rng)
        children = problem.repair(children).reshape(len(children), -1)
        children = children[_unique_rows(children, against=X)]
        if len(children):
            Fc = problem.evaluate(children.reshape(-1, *shape))
            X = np.concatenate([X, children])
            F = np.concatenate([F, Fc])
        if len(X) > pop_size:                 # elitist survival
            keep = []
            for front in non_dominated_fronts(F):
                if len(keep) + len(front) <= pop_size:
                    keep.extend(front.tolist())
                else:
                    crowd_f = crowding_distance(F[front])
                    order = np.argsort(-crowd_f, kind="stable")
                    keep.extend(front[order[:pop_size - len(keep)]].tolist())
                    break
            keep = np.array(keep)
            X, F = X[keep], F[keep]
    return X.reshape(-1, *shape), F
