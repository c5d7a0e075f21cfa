class Problem(object):
    def __init__(self, n_var=-1, n_obj=1, n_constr=0, xl=None, xu=None,
                 type_var=None, **kwargs):
        self.n_var, self.n_obj, self.n_constr = n_var, n_obj, n_constr
        self.xl, self.xu, self.type_var = xl, xu, type_var

    def evaluate(self, X, *args, **kwargs):
        out = {}
        self._evaluate(X, out, *args, **kwargs)
        return out["F"]

    def _evaluate(self, X, out, *args, **kwargs):
        raise NotImplementedError
