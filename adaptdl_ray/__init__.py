"""``import adaptdl_ray`` -> ``adaptdl_b200.ray`` (see adaptdl_b200/compat.py): scripts
written for petuum/adaptdl run on this framework without edits."""
from adaptdl_b200.compat import alias as _alias

_alias("adaptdl_ray", "adaptdl_b200.ray")
