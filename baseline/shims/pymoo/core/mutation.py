class Mutation(object):
    def do(self, problem, X, **kwargs):
        return self._do(problem, X.copy(), **kwargs)

    def _do(self, problem, X, **kwargs):
        raise NotImplementedError
