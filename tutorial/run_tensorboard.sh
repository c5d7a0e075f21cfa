#!/bin/bash
# The tutorial with TensorBoard logging, as a standalone elastic job on this
# machine; watch it with `tensorboard --logdir /tmp/adaptdl-tutorial-tb`.
export ADAPTDL_TENSORBOARD_LOGDIR=${ADAPTDL_TENSORBOARD_LOGDIR:-/tmp/adaptdl-tutorial-tb}
cd "$(dirname "$0")/.." && python -m adaptdl_b200.sched.local --gpus 2 \
    tutorial/mnist_tensorboard.py --epochs 3
