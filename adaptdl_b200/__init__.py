"""adaptdl_b200 -- a Blackwell-native elastic data-parallel training engine.

Same capabilities and public API as petuum/adaptdl (``adaptdl.torch.*``,
``adaptdl.env``, ``adaptdl.checkpoint``, ``adaptdl.collective``,
``adaptdl.goodput``, ``adaptdl.sched_hints``), re-designed for 8xB200:
gradient buckets live in NVLink-mapped symmetric memory and are reduced by a
hand-written sm_100a kernel that fuses the all-reduce with the gradient
scale/cast and the gradient-noise-scale statistics.

Like the reference (``adaptdl/adaptdl/__init__.py`` is empty), importing the
top-level package pulls in nothing heavy; import the sub-modules explicitly::

    import adaptdl_b200.torch as adl
"""

__version__ = "0.1.0"
