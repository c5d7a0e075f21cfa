"""Test-only stand-in for the *legacy* torchtext API (<= 0.8: ``data.Field``,
``data.Example``, ``data.Dataset``, ``data.utils.get_tokenizer``,
``datasets.WikiText2.splits``), which is
not installable here and which the reference's own ``data_test.py`` builds its
BPTT fixture from. Only what that fixture touches; used by
``tests/test_reference_suite.py`` / ``tests/test_api_surface.py`` and nowhere on a
product path."""
from . import data, datasets  # noqa: F401
