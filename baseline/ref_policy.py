"""Loader of the reference's UNMODIFIED scheduling policy
(``baseline/_ref/_ref_sched/adaptdl_sched/policy``, copied there by
``baseline/install_reference.sh``) on top of the pymoo stand-in of
``baseline/shims/pymoo``. Used by ``tools/sched_sim.py --policy reference``
and ``tools/policy_bench.py --search reference``; never by the product."""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SCHED = os.path.join(HERE, "_ref", "_ref_sched")


def available():
    return os.path.isfile(os.path.join(SCHED, "adaptdl_sched", "policy",
                                       "pollux.py"))


def load():
    """``(PolluxPolicy, JobInfo, NodeInfo, SpeedupFunction)`` of the
    reference."""
    import numpy as np
    if not hasattr(np, "int"):           # the reference predates numpy 1.24
        np.int, np.float = int, float
    for path in (os.path.join(HERE, "shims"), SCHED):
        while path in sys.path:          # must come BEFORE the repository root
            sys.path.remove(path)
        sys.path.insert(0, path)
    # the repository root ships an ``adaptdl_sched`` ALIAS package (this
    # framework under the reference's name): make sure the real one is found
    for name in [n for n in sys.modules
                 if n == "adaptdl_sched" or n.startswith("adaptdl_sched.")]:
        del sys.modules[name]
    finders = [f for f in sys.meta_path
               if type(f).__name__ == "_AliasFinder"
               and getattr(f, "name", None) == "adaptdl_sched"]
    for finder in finders:
        sys.meta_path.remove(finder)
    try:
        pollux = importlib.import_module("adaptdl_sched.policy.pollux")
        utils = importlib.import_module("adaptdl_sched.policy.utils")
        speedup = importlib.import_module("adaptdl_sched.policy.speedup")
    finally:
        sys.meta_path[:0] = finders
    assert SCHED in pollux.__file__, pollux.__file__
    return (pollux.PolluxPolicy, utils.JobInfo, utils.NodeInfo,
            speedup.SpeedupFunction)
