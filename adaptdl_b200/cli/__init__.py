"""Command-line front-end: ``adaptdl-b200 submit | ls | logs | cp |
tensorboard {create,delete,list,proxy}`` (reference: ``cli/bin/adaptdl`` and
``cli/adaptdl_cli``). Talks to the cluster through ``kubectl`` / ``docker``
subprocesses only, so it has no Python dependencies beyond the standard
library + PyYAML."""
