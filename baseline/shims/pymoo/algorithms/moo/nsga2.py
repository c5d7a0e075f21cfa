import numpy as np

from pymoo.core.population import Population
from pymoo.util.nds.non_dominated_sorting import fronts_of


def crowding(F):
    n, m = F.shape
    dist = np.zeros(n)
    if n <= 2:
        return np.full(n, np.inf)
    for k in range(m):
        order = np.argsort(F[:, k], kind="stable")
        col = F[order, k]
        dist[order[0]] = dist[order[-1]] = np.inf
        if col[-1] > col[0]:
            dist[order[1:-1]] += (col[2:] - col[:-2]) / (col[-1] - col[0])
    return dist


def rank_and_crowding(F):
    rank = np.zeros(len(F), dtype=int)
    crowd = np.zeros(len(F))
    for r, front in enumerate(fronts_of(F)):
        rank[front] = r
        crowd[front] = crowding(F[front])
    return rank, crowd


def unique_rows(X, against=None):
    seen = set() if against is None else {row.tobytes() for row in against}
    keep = np.zeros(len(X), dtype=bool)
    for i, row in enumerate(X):
        key = row.tobytes()
        if key not in seen:
            seen.add(key)
            keep[i] = True
    return keep


class NSGA2(object):
    def __init__(self, pop_size=100, sampling=None, crossover=None,
                 mutation=None, repair=None, eliminate_duplicates=True,
                 n_offsprings=None, **kwargs):
        self.pop_size = pop_size
        self.sampling = sampling
        self.crossover, self.mutation, self.repair = \
            crossover, mutation, repair
        self.eliminate_duplicates = eliminate_duplicates
        self.n_offsprings = n_offsprings or pop_size

    # -- pieces ---------------------------------------------------------

    def _repair(self, problem, X):
        if self.repair is None:
            return X
        return np.asarray(self.repair.do(problem, Population(X)).get("X"))

    def _mating(self, problem, X, F, count):
        rank, crowd = rank_and_crowding(F)
        n = len(X)

        def tournament(k):
            a = np.random.randint(n, size=k)
            b = np.random.randint(n, size=k)
            better = (rank[a] < rank[b]) | (
                (rank[a] == rank[b]) & (crowd[a] >= crowd[b]))
            return np.where(better, a, b)
        n_matings = -(-count // self.crossover.n_offsprings)
        parents = np.stack([X[tournament(n_matings)],
                            X[tournament(n_matings)]])
        children = self.crossover.do(problem, parents)
        children = self.mutation.do(problem, children)
        return self._repair(problem, children.reshape(len(children), -1))

    def _survive(self, X, F):
        if len(X) <= self.pop_size:
            return X, F
        keep = []
        for front in fronts_of(F):
            if len(keep) + len(front) <= self.pop_size:
                keep.extend(front.tolist())
            else:
                order = np.argsort(-crowding(F[front]), kind="stable")
                keep.extend(front[order[:self.pop_size - len(keep)]].tolist())
                break
        keep = np.array(keep)
        return X[keep], F[keep]

    # -- driver ---------------------------------------------------------

    def run(self, problem, n_gen):
        X = np.asarray(self.sampling)
        X = self._repair(problem, X.reshape(len(X), -1).copy())
        if self.eliminate_duplicates:
            X = X[unique_rows(X)]
        F = np.asarray(problem.evaluate(X), dtype=float)
        X, F = self._survive(X, F)
        # the first "generation" of pymoo is the evaluation of the start
        # population
        for _ in range(max(n_gen - 1, 0)):
            off = np.zeros((0, X.shape[1]), dtype=X.dtype)
            for _attempt in range(10):
                fresh = self._mating(problem, X, F,
                                     self.n_offsprings - len(off))
                if self.eliminate_duplicates:
                    fresh = fresh[unique_rows(
                        fresh, against=np.concatenate([X, off]))]
                off = np.concatenate([off, fresh])[:self.n_offsprings]
                if len(off) >= self.n_offsprings:
                    break
            if len(off):
                F_off = np.asarray(problem.evaluate(off), dtype=float)
                X = np.concatenate([X, off])
                F = np.concatenate([F, F_off])
            X, F = self._survive(X, F)
        return X, F
