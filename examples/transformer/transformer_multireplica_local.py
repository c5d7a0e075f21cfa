#!/usr/bin/env python
"""The transformer example as a 2-replica gloo job on this machine, with tiny
dimensions -- a smoke test of multi-replica training without GPUs (reference:
examples/transformer/transformer_multireplica_local.py)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from adaptdl_b200.utils import pick_unused_port  # noqa: E402


def main():
    port = pick_unused_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, ADAPTDL_MASTER_ADDR="127.0.0.1",
                   ADAPTDL_MASTER_PORT=str(port), ADAPTDL_NUM_REPLICAS="2",
                   ADAPTDL_REPLICA_RANK=str(rank), ADAPTDL_NUM_NODES="1",
                   CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen(
            [sys.executable, os.path.join(HERE, "transformer.py"),
             "--epochs", "1", "--tokens", "20000", "--emsize", "32",
             "--nhid", "32", "--nlayers", "1", "--bptt", "16"], env=env))
    codes = [p.wait() for p in procs]
    sys.exit(max(codes))


if __name__ == "__main__":
    main()
