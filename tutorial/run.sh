#!/bin/bash
# Run the final tutorial script as a standalone elastic job on this machine
# (2 replicas, then a rescale to 4 after 30 s).
cd "$(dirname "$0")/.." && python -m adaptdl_b200.sched.local --gpus 4 \
    --schedule 2,4 --interval 30 tutorial/mnist_step_5.py --epochs 3
