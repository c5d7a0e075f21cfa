"""Launch accounting for the hand-written compute kernels (``bench.py``
reports how many of this repo's kernels ran inside the timed region). A CUDA
graph replay re-runs the launches recorded at capture time, so
``parallel/graph.py`` adds them back per replay."""

_LAUNCHES = [0]


def add(n=1):
    _LAUNCHES[0] += n


def total():
    return _LAUNCHES[0]
