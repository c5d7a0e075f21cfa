from . import sdk  # noqa: F401
