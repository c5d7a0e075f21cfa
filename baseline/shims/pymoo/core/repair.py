class Repair(object):
    def do(self, problem, pop, **kwargs):
        return self._do(problem, pop, **kwargs)

    def _do(self, problem, pop, **kwargs):
        return pop
