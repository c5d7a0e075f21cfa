"""``adaptdl.conftest`` of the reference ships the ``elastic_multiprocessing``
test decorator inside the package (``adaptdl/adaptdl/conftest.py:25-100``) and
user test-suites import it from there; here it lives in
:mod:`adaptdl_b200.utils.testing` and this module keeps the import path."""

from adaptdl_b200.utils.testing import (  # noqa: F401
    elastic_multiprocessing, reset_global_state)

__all__ = ["elastic_multiprocessing", "reset_global_state"]
