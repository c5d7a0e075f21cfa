"""Every example workload runs end to end (tiny settings, CPU, one replica):
the scripts are the reference's workloads, so a user switching over starts
from them."""
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EXAMPLES = {
    "linear_regression": ["examples/linear_regression/main.py", "--epochs",
                          "2", "--size", "512", "--autoscale-bsz"],
    "pytorch-cifar": ["examples/pytorch-cifar/main.py", "--model", "LeNet",
                      "--bs", "64", "--epochs", "1", "--synthetic",
                      "--synthetic-size", "256", "--autoscale-bsz"],
    "ncf": ["examples/NCF/main.py", "--epochs", "1", "--positives", "2000",
            "--autoscale-bsz"],
    "transformer": ["examples/transformer/transformer.py", "--epochs", "1",
                    "--tokens", "20000", "--emsize", "32", "--nhid", "32",
                    "--autoscale-bsz"],
    "dcgan": ["examples/dcgan/dcgan.py", "--epochs", "1", "--images", "128",
              "--autoscale-bsz"],
    "bert-mlm": ["examples/BERT/mlm_task_adaptdl.py", "--epochs", "1",
                 "--emsize", "32", "--nhid", "64", "--nlayers", "2",
                 "--nhead", "2", "--tokens", "20000", "--ntoken", "1000",
                 "--batch_size", "8"],
}


@pytest.mark.parametrize("name", sorted(EXAMPLES))
@pytest.mark.timeout(600)
def test_example_runs(name):
    with tempfile.TemporaryDirectory() as ckpt:
        env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="",
                   OMP_NUM_THREADS="2", ADAPTDL_CHECKPOINT_PATH=ckpt)
        for stale in ("RANK", "WORLD_SIZE", "LOCAL_RANK",
                      "ADAPTDL_NUM_REPLICAS", "ADAPTDL_REPLICA_RANK"):
            env.pop(stale, None)
        command = [sys.executable, os.path.join(ROOT, EXAMPLES[name][0])] \
            + EXAMPLES[name][1:]
        done = subprocess.run(command, env=env, cwd=ROOT, timeout=550,
                              stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True)
        assert done.returncode == 0, done.stdout[-3000:]
