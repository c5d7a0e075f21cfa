"""The reference's OWN unit tests, unmodified, run against this framework.

petuum/adaptdl keeps its tests next to the code (``adaptdl/adaptdl/*_test.py``,
``adaptdl/adaptdl/torch/*_test.py``, ``sched/adaptdl_sched/**/*_test.py``).
They are written against ``import adaptdl`` / ``import adaptdl_sched``; the
alias packages in the repository root resolve those names to
``adaptdl_b200``, so the files can be executed as they are: that is the
strongest statement of "a user of the reference can switch" this repository
can make on a CPU box.

How: the test files are copied at run time from ``/root/reference`` (read-only,
nothing of it is kept in this repository) into a flat temporary directory --
inside the reference tree pytest would import the reference's own package
through the ``__init__.py`` chain -- and a pytest subprocess runs them with
``PYTHONPATH = <this repo> : baseline/shims : baseline/shims_test``
(stand-ins for the third-party modules that are not installable here:
``portpicker``, legacy ``torchtext.data``).

``validator_test.py`` needs the ``aiohttp_client`` fixture (aiohttp's own
pytest plugin, loaded with ``-p``) and a ``kubernetes_asyncio`` module for its
``ApiException`` (``tests/fixtures/fake_k8s``).

Not run, and why:

* ``ray/adaptdl_ray/**`` -- they start a real Ray cluster (``ray.init``,
  ``tune.run``) or replace private members of the reference's controller;
  Ray is not installable here. ``tests/test_ray*.py`` drive the same classes
  against recorded API fixtures.

The suite is skipped where the reference checkout does not exist (e.g. on the
GPU box).
"""

import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"

pytestmark = pytest.mark.skipif(
    not os.path.isdir(os.path.join(REFERENCE, "adaptdl", "adaptdl")),
    reason="the reference checkout is not on this machine")

# Python 3.12 compatibility of the reference's test *code* (not of what it
# tests): non_preemptible_test.py calls random.sample() on a dict view, which
# Python >= 3.11 refuses.
CONFTEST = '''
import random
_sample = random.sample


def _sample_any(population, k, **kwargs):
    if not hasattr(population, "__getitem__"):
        population = list(population)
    return _sample(population, k, **kwargs)


random.sample = _sample_any
'''

GROUPS = {
    # name: (directories of the reference, excluded files, minimum passes)
    "runtime_and_trainer": (
        ["adaptdl/adaptdl", "adaptdl/adaptdl/torch"], (), 750),
    "scheduler": (
        ["sched/adaptdl_sched", "sched/adaptdl_sched/policy"], (), 30),
}


def _run(tmp_path, group):
    dirs, excluded, at_least = GROUPS[group]
    work = tmp_path / group
    work.mkdir()
    names = []
    for rel in dirs:
        src = os.path.join(REFERENCE, rel)
        for name in sorted(os.listdir(src)):
            if name.endswith("_test.py") and name not in excluded:
                shutil.copy(os.path.join(src, name), str(work / name))
                names.append(name)
    assert names, "no reference tests found under {}".format(dirs)
    (work / "conftest.py").write_text(CONFTEST)
    env = dict(os.environ)
    for key in [k for k in env if k.startswith("ADAPTDL_")]:
        del env[key]
    env["PYTHONPATH"] = os.pathsep.join([
        ROOT, os.path.join(ROOT, "baseline", "shims"),
        os.path.join(ROOT, "baseline", "shims_test"),
        os.path.join(ROOT, "tests", "fixtures", "fake_k8s")])
    cmd = [sys.executable, "-m", "pytest", "-q", "--no-header",
           "-p", "no:cacheprovider", "-p", "aiohttp.pytest_plugin",
           "-W", "ignore", "--rootdir", str(work)]
    try:
        import xdist  # noqa: F401
        cmd += ["-n", str(min(4, os.cpu_count() or 1))]
    except ImportError:
        pass
    proc = subprocess.run(cmd + names, cwd=str(work), env=env, timeout=1500,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True)
    tail = proc.stdout[-6000:]
    # the modules under test must be this framework's, not the reference's
    check = subprocess.run(
        [sys.executable, "-c",
         "import adaptdl.checkpoint as m, adaptdl_sched.policy.pollux as p;"
         "print(m.__file__); print(p.__file__)"],
        cwd=str(work), env=env, stdout=subprocess.PIPE, text=True, timeout=300)
    for line in check.stdout.split():
        assert line.startswith(os.path.join(ROOT, "adaptdl_b200")), line
    summary = proc.stdout.strip().splitlines()[-1] if proc.stdout else ""
    passed = re.search(r"(\d+) passed", summary)
    assert proc.returncode == 0, tail
    assert not re.search(r"\d+ (failed|error)", summary), tail
    assert passed and int(passed.group(1)) >= at_least, tail


def test_reference_runtime_and_trainer_tests_pass_unmodified(tmp_path):
    """checkpoint, collective, reducer, goodput (+fit), _metrics, accumulator,
    data (sampler, loader, BPTT iterator), epoch, gradient noise scale,
    parallel, scaling rules: 12 files, ~760 cases."""
    _run(tmp_path, "runtime_and_trainer")


def test_reference_scheduler_tests_pass_unmodified(tmp_path):
    """Pollux policy (incl. non-preemptible jobs), speedup function, resource
    arithmetic, admission webhook."""
    _run(tmp_path, "scheduler")
