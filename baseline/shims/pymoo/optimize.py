from pymoo.util.nds.non_dominated_sorting import fronts_of


class Result(object):
    """``X`` / ``F``: the non-dominated set of the final population (what
    pymoo returns for a multi-objective run); ``pop_X`` / ``pop_F``: all of
    it."""


def minimize(problem, algorithm, termination=("n_gen", 100), seed=None,
             **kwargs):
    assert termination[0] == "n_gen"
    if seed is not None:
        import numpy as np
        np.random.seed(seed)
    X, F = algorithm.run(problem, int(termination[1]))
    front = fronts_of(F)[0]
    result = Result()
    result.X, result.F = X[front], F[front]
    result.pop_X, result.pop_F = X, F
    return result
