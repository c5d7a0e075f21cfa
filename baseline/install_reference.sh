#!/bin/bash
# Offline install of the UNMODIFIED reference trainer package into
# baseline/_ref (git-ignored, travels to the GPU box with the snapshot) plus
# the reference's own example models used by `bench.py --impl reference`.
# /root/reference has no top-level setup.py (four separate packages), so the
# trainer package directory is installed, from a /tmp copy (the source tree
# is read-only and setuptools writes egg-info next to setup.py); --no-deps
# because autograd/portpicker/semver/redis are not in the wheelhouse (tiny
# stand-ins live in baseline/shims/).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
rm -rf /tmp/refsrc "$HERE/_ref"
cp -r /root/reference/adaptdl /tmp/refsrc
python -m pip install --no-index --no-build-isolation \
    --find-links /opt/wheelhouse --no-deps --target "$HERE/_ref" /tmp/refsrc
mkdir -p "$HERE/_ref/_ref_examples"
cp -r /root/reference/examples/pytorch-cifar/models "$HERE/_ref/_ref_examples/cifar_models"
cp /root/reference/examples/BERT/model.py "$HERE/_ref/_ref_examples/bert_model.py"
cp /root/reference/examples/NCF/model.py "$HERE/_ref/_ref_examples/ncf_model.py"
# the reference's scheduling policy (sched/adaptdl_sched/policy: pollux.py,
# speedup.py, utils.py), unmodified, for tools/sched_sim.py --policy reference
# and tools/policy_bench.py --search reference (needs the pymoo stand-in in
# baseline/shims/)
mkdir -p "$HERE/_ref/_ref_sched/adaptdl_sched/policy"
touch "$HERE/_ref/_ref_sched/adaptdl_sched/__init__.py"
for f in __init__.py pollux.py speedup.py utils.py; do
    cp "/root/reference/sched/adaptdl_sched/policy/$f" "$HERE/_ref/_ref_sched/adaptdl_sched/policy/$f"
done
echo "reference installed into $HERE/_ref"
