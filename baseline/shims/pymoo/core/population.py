import numpy as np


class Population(object):
    """The little of pymoo's Population the reference's Repair touches:
    ``pop.get("X")`` and ``pop.new("X", array)``."""

    def __init__(self, X):
        self.X = np.asarray(X)

    def get(self, key):
        assert key == "X"
        return self.X

    def new(self, key, value):
        assert key == "X"
        return Population(value)

    def __len__(self):
        return len(self.X)
