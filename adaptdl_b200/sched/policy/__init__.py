"""Scheduling policy: job/node descriptions, speedup functions derived from
the goodput model, a dependency-free NSGA-II, and the Pollux co-adaptive
allocation policy built on it."""

from .utils import JobInfo, NodeInfo  # noqa: F401
from .speedup import SpeedupFunction  # noqa: F401
from .pollux import PolluxPolicy  # noqa: F401
