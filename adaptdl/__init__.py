"""``import adaptdl`` -> ``adaptdl_b200`` (see adaptdl_b200/compat.py): scripts
written for petuum/adaptdl run on this framework without edits."""
from adaptdl_b200.compat import alias as _alias

_alias("adaptdl", "adaptdl_b200")
