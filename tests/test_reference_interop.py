"""On-disk compatibility with the reference, both directions.

The same job (``tests/interop_job.py``, written against ``import adaptdl``)
is preempted under one implementation and resumed under the other: the
unmodified reference package from ``baseline/_ref`` and this framework's
``adaptdl`` alias. The resumed run must start from exactly the saved model,
optimizer, scheduler, epoch, data-loader position and accumulator state, and
finish where an uninterrupted run of the first implementation finishes.
"""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
JOB = os.path.join(ROOT, "tests", "interop_job.py")
EPOCHS = 60            # 720 tiny steps: long enough to be preempted mid-way

pytestmark = pytest.mark.skipif(
    not os.path.isdir(os.path.join(REF, "adaptdl")),
    reason="reference package not installed (baseline/install_reference.sh)")


def _run(impl, ckpt, restarts, *job_args):
    env = {k: v for k, v in os.environ.items()
           if not k.startswith("ADAPTDL_") and k != "PYTHONPATH"}
    if impl == "reference":
        path = [REF, os.path.join(ROOT, "baseline", "shims")]
        # the reference calls torch.load() with its 2021 default; torch >=
        # 2.6 refuses the numpy arrays of optimizer.state["gns"] otherwise
        # (the reference cannot even resume its OWN checkpoints without it)
        env["TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD"] = "1"
    else:
        path = [ROOT]                       # adaptdl/ alias package
    env.update(PYTHONPATH=os.pathsep.join(path), OMP_NUM_THREADS="1",
               CUDA_VISIBLE_DEVICES="", ADAPTDL_CHECKPOINT_PATH=str(ckpt),
               ADAPTDL_NUM_RESTARTS=str(restarts),
               ADAPTDL_MASTER_ADDR="127.0.0.1")
    proc = subprocess.run([sys.executable, JOB] + list(job_args), env=env,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=300, cwd=str(ckpt))
    rows = [json.loads(line[len("STATE "):])
            for line in proc.stdout.splitlines() if line.startswith("STATE ")]
    return proc, rows


def _close(a, b, tol=1e-5):
    return all(abs(x - y) <= tol * max(1.0, abs(x), abs(y))
               for x, y in zip(a, b))


@pytest.mark.parametrize("first,second", [("reference", "own"),
                                          ("own", "reference")])
def test_checkpoint_written_by_one_is_resumed_by_the_other(tmp_path, first,
                                                           second):
    whole = tmp_path / "whole"
    split = tmp_path / "split"
    whole.mkdir()
    split.mkdir()
    # uninterrupted run of the first implementation: the trajectory to match
    # (in the background while the preempted one runs)
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(1)
    whole_run = pool.submit(_run, first, whole, 0, "--epochs", str(EPOCHS))
    # same job, preempted early in epoch 1 (12 steps per epoch). The
    # reference checkpoints at the very next iteration; this framework agrees
    # on the iteration through its ~0.1 s consensus beat, i.e. some (tiny,
    # sub-millisecond) steps later -- the test reads where it stopped.
    proc, a_rows = _run(first, split, 0, "--epochs", str(EPOCHS),
                        "--stop-after-steps", "17")
    assert proc.returncode == 143, (proc.returncode, proc.stderr[-3000:])
    saved = [r for r in a_rows if r["tag"] == "step"][-1]
    at = saved["epoch"]
    assert 1 <= at < EPOCHS - 1, saved
    done_in_epoch = len([r for r in a_rows
                         if r["tag"] == "step" and r["epoch"] == at])
    assert any(n.startswith("checkpoint-") for n in os.listdir(str(split)))

    # ... and resumed by the OTHER implementation
    proc, b_rows = _run(second, split, 1, "--epochs", str(EPOCHS))
    assert proc.returncode == 0, proc.stderr[-3000:]
    whole_proc, ref_rows = whole_run.result()
    pool.shutdown()
    assert whole_proc.returncode == 0, whole_proc.stderr[-3000:]
    expected = "adaptdl" if first == "reference" else "adaptdl_b200"
    assert ref_rows[0]["impl"] == expected, ref_rows[0]
    assert (REF in ref_rows[0]["file"]) == (first == "reference")
    start = b_rows[0]
    assert start["impl"] != a_rows[0]["impl"]
    assert _close(start["params"], saved["params"]), (start, saved)
    assert _close([start["momentum"]], [saved["momentum"]])
    assert start["lr"] == pytest.approx(saved["lr"])
    assert start["sched_epoch"] == saved["sched_epoch"] == at
    assert start["finished_epochs"] == at
    assert start["has_gns_state"]

    # the data loader continues where the first incarnation stopped, and the
    # epoch it finishes reports the whole epoch's accumulated statistics
    resumed_steps = [r for r in b_rows
                     if r["tag"] == "step" and r["epoch"] == at]
    assert len(resumed_steps) == 12 - done_in_epoch
    if resumed_steps:
        same_step_in_whole = [r for r in ref_rows if r["tag"] == "step"
                              and r["step"] == saved["step"] + 1][0]
        assert resumed_steps[0]["first_feature"] == pytest.approx(
            same_step_in_whole["first_feature"])
    ends = {r["epoch"]: r for r in b_rows if r["tag"] == "epoch_end"}
    whole_ends = {r["epoch"]: r for r in ref_rows if r["tag"] == "epoch_end"}
    assert sorted(ends) == list(range(at, EPOCHS))
    assert ends[at]["batches"] == whole_ends[at]["batches"] == 12
    assert ends[at]["loss_sum"] == pytest.approx(whole_ends[at]["loss_sum"],
                                                 rel=1e-4)

    # and the two-implementation run ends where the single one does
    final, want = b_rows[-1], ref_rows[-1]
    assert final["tag"] == want["tag"] == "done"
    assert _close(final["params"], want["params"], tol=1e-4), (final, want)
    assert final["finished_epochs"] == want["finished_epochs"] == EPOCHS
