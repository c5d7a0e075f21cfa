import numpy as np


class Crossover(object):
    def __init__(self, n_parents, n_offsprings, prob=0.9, **kwargs):
        self.n_parents, self.n_offsprings, self.prob = \
            n_parents, n_offsprings, prob

    def do(self, problem, parents, **kwargs):
        """``parents``: [n_parents, n_matings, n_var] -> offspring
        [n_offsprings * n_matings, n_var]; a mating is crossed with
        probability ``prob``, else its parents pass through."""
        children = self._do(problem, parents.copy(), **kwargs)
        keep = np.random.random(parents.shape[1]) >= self.prob
        if self.n_offsprings == self.n_parents:
            children[:, keep] = parents[:, keep]
        return children.reshape(-1, parents.shape[-1])

    def _do(self, problem, X, **kwargs):
        raise NotImplementedError
