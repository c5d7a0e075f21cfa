"""Cluster / single-box workload suite.

The reference ships its end-to-end tests as shell heredocs piped to the CLI
(``tests/long-workload/*.sh`` x12, ``tests/short-workload/*.sh`` x2,
``tests/testworkload.sh``, ``tests/test-localmode2.sh``). Here the same
workloads are DATA: one table, three ways to use it --

    python tests/workloads/workloads.py list
    python tests/workloads/workloads.py yaml resnet18-cifar10-elastic       # AdaptDLJob manifest
    python tests/workloads/workloads.py submit resnet18-cifar10-elastic     # through adaptdl_b200.cli
    python tests/workloads/workloads.py local transformer-wikitext2-elastic \
        --gpus 8 --schedule 2,4,8,4 --interval 30                           # no Kubernetes: sched.local
    python tests/workloads/workloads.py soak --jobs 6                       # keep 6 random jobs alive on the cluster
                                                                            # (reference tests/testworkload.sh)
    python tests/workloads/workloads.py standalone3                         # three independent single-replica jobs on this
                                                                            # box (reference tests/test-localmode2.sh)

``tests/test_cli.py::test_workload_specs_validate`` runs every manifest
through the CLI's job preparation and the scheduler's validator on CPU.
"""

import argparse
import copy
import os
import subprocess
import sys

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__))))
API_VERSION = "adaptdl.petuum.com/v1"
IMAGE_ROOT = "/opt/adaptdl_b200"      # WORKDIR of deploy/docker/Dockerfile.trainer

# name -> (suite, script, args, spec overrides)
WORKLOADS = {
    "resnet18-cifar10-elastic": (
        "long", "examples/pytorch-cifar/main.py",
        ["--model=ResNet18", "--bs=128", "--lr=0.1", "--epochs=60",
         "--autoscale-bsz"], {}),
    "resnet18-cifar10-elastic-min-replicas": (
        "long", "examples/pytorch-cifar/main.py",
        ["--model=ResNet18", "--bs=128", "--lr=0.1", "--epochs=60",
         "--autoscale-bsz"], {"minReplicas": 2, "maxReplicas": 4}),
    "resnet18-cifar10-inelastic": (
        "long", "examples/pytorch-cifar/main.py",
        ["--model=ResNet18", "--bs=128", "--lr=0.1", "--epochs=60"],
        {"minReplicas": 2, "maxReplicas": 2}),
    "resnet18-cifar10-mixed-precision": (
        "long", "examples/pytorch-cifar/main.py",
        ["--model=ResNet18", "--bs=128", "--lr=0.1", "--epochs=60",
         "--autoscale-bsz", "--mixed-precision"], {}),
    # one pod per node: all replicas of a node in one container, so the
    # peer-memory reducer is used (docs/commandline.md)
    "resnet18-cifar10-elastic-node-pods": (
        "long", "examples/pytorch-cifar/main.py",
        ["--model=ResNet18", "--bs=128", "--lr=0.1", "--epochs=60",
         "--autoscale-bsz"], {"podPerNode": True, "maxReplicas": 8}),
    "densenet121-cifar10": (
        "long", "examples/pytorch-cifar/main.py",
        ["--model=DenseNet121", "--bs=128", "--lr=0.1", "--epochs=60",
         "--autoscale-bsz"], {}),
    "bert": (
        "long", "examples/BERT/mlm_task_adaptdl.py",
        ["--epochs=1", "--batch_size=32", "--bptt=128", "--autoscale-bsz"],
        {}),
    "dcgan": (
        "long", "examples/dcgan/dcgan.py",
        ["--epochs=5", "--autoscale-bsz"], {}),
    "ncf": (
        "long", "examples/NCF/main.py",
        ["--epochs=20", "--autoscale-bsz"], {}),
    "ncf-accumulation": (
        "long", "examples/NCF/main.py",
        ["--epochs=20", "--autoscale-bsz", "--gradient-accumulation"], {}),
    "transformer-wikitext2": (
        "long", "examples/transformer/transformer.py",
        ["--epochs=3", "--bs=20", "--lr=5.0"], {}),
    "transformer-wikitext2-elastic": (
        "long", "examples/transformer/transformer.py",
        ["--epochs=3", "--bs=20", "--lr=5.0", "--autoscale-bsz"], {}),
    "lr-elastic-cpu": (
        "long", "examples/linear_regression/main.py",
        ["--epochs=90", "--autoscale-bsz"], {"cpu": True}),
    "resnet18-cifar10-short": (
        "short", "examples/pytorch-cifar/main.py",
        ["--model=ResNet18", "--bs=128", "--lr=0.1", "--epochs=2",
         "--autoscale-bsz", "--synthetic"], {}),
    "densenet121-cifar10-short": (
        "short", "examples/pytorch-cifar/main.py",
        ["--model=DenseNet121", "--bs=128", "--lr=0.1", "--epochs=2",
         "--autoscale-bsz", "--synthetic"], {}),
}


def manifest(name, image_root=IMAGE_ROOT):
    """The AdaptDLJob object for workload ``name``."""
    suite, script, args, overrides = WORKLOADS[name]
    overrides = copy.deepcopy(overrides)
    cpu = overrides.pop("cpu", False)
    launcher = ["-m", "adaptdl_b200.launch"] \
        if overrides.get("podPerNode") else []
    container = {
        "name": "main",
        "command": ["python3"] + launcher
        + [os.path.join(image_root, script)] + list(args),
        "env": [{"name": "PYTHONUNBUFFERED", "value": "true"}],
    }
    if cpu:
        container["resources"] = {"limits": {"cpu": 1}}
    else:
        container["resources"] = {"limits": {"nvidia.com/gpu": 1}}
    spec = {"template": {"spec": {"containers": [container]}}}
    spec.update(overrides)
    return {"apiVersion": API_VERSION, "kind": "AdaptDLJob",
            "metadata": {"generateName": name + "-"}, "spec": spec}


def local_command(name):
    _, script, args, _ = WORKLOADS[name]
    return [os.path.join(ROOT, script)] + list(args)


def active_jobs():
    """Names of AdaptDLJobs that have not reached a final phase."""
    import json
    out = subprocess.run(["kubectl", "get", "adaptdljobs", "-o", "json"],
                         stdout=subprocess.PIPE, check=True).stdout
    items = json.loads(out).get("items", [])
    return [item["metadata"]["name"] for item in items
            if (item.get("status") or {}).get("phase")
            not in ("Succeeded", "Failed")]


def soak(args):
    """Keep ``--jobs`` randomly chosen workloads of a suite alive: whenever
    fewer are active, submit more (scheduler soak test)."""
    import random
    import time
    names = [n for n, w in WORKLOADS.items() if w[0] == args.suite]
    rounds = 0
    while args.rounds == 0 or rounds < args.rounds:
        missing = args.jobs - len(active_jobs())
        for _ in range(max(missing, 0)):
            name = random.choice(names)
            print("submitting", name, flush=True)
            main(["submit", name])
        rounds += 1
        time.sleep(args.period)
    return 0


def standalone(names):
    """Independent single-replica jobs side by side on this machine, each with
    its own checkpoint directory (standalone mode, no scheduler)."""
    import tempfile
    procs = []
    for i, name in enumerate(names):
        env = dict(os.environ, PYTHONPATH=ROOT,
                   ADAPTDL_CHECKPOINT_PATH=tempfile.mkdtemp(
                       prefix="adaptdl-b200-standalone-"),
                   ADAPTDL_JOB_ID="standalone/{}-{}".format(name, i))
        cmd = [sys.executable] + local_command(name)
        procs.append(subprocess.Popen(cmd, env=env))
    codes = [p.wait() for p in procs]
    print("exit codes:", codes)
    return 0 if all(c == 0 for c in codes) else 1


def main(argv=None):
    parser = argparse.ArgumentParser()
    sub = parser.add_subparsers(dest="verb", required=True)
    sub.add_parser("list")
    p = sub.add_parser("yaml")
    p.add_argument("name")
    p = sub.add_parser("submit")
    p.add_argument("name")
    p.add_argument("--tensorboard", default=None)
    p = sub.add_parser("local")
    p.add_argument("name")
    p.add_argument("--gpus", type=int, default=None)
    p.add_argument("--schedule", default="")
    p.add_argument("--interval", type=float, default=30.0)
    p.add_argument("--adaptive", action="store_true")
    p = sub.add_parser("soak")
    p.add_argument("--jobs", type=int, default=4)
    p.add_argument("--suite", default="short")
    p.add_argument("--period", type=float, default=60.0)
    p.add_argument("--rounds", type=int, default=0, help="0 = forever")
    p = sub.add_parser("standalone3")
    p.add_argument("--names", default="lr-elastic-cpu,lr-elastic-cpu,"
                                      "lr-elastic-cpu")
    args = parser.parse_args(argv)
    if args.verb == "soak":
        return soak(args)
    if args.verb == "standalone3":
        return standalone(args.names.split(","))
    if args.verb == "list":
        for name, (suite, script, wargs, over) in sorted(WORKLOADS.items()):
            print("{:6s} {:40s} {} {}".format(suite, name, script,
                                              " ".join(wargs)))
        return 0
    if args.verb == "yaml":
        print(yaml.safe_dump(manifest(args.name), sort_keys=False))
        return 0
    if args.verb == "submit":
        cmd = [sys.executable, "-m", "adaptdl_b200.cli", "submit", ROOT,
               "-d", os.path.join(ROOT, "deploy", "docker", "Dockerfile.trainer"), "-f", "-",
               "--checkpoint-storage-size", "1Gi"]
        if args.tensorboard:
            cmd += ["--tensorboard", args.tensorboard]
        return subprocess.run(cmd, input=yaml.safe_dump(
            manifest(args.name)).encode()).returncode
    cmd = [sys.executable, "-m", "adaptdl_b200.sched.local"]
    if args.gpus:
        cmd += ["--gpus", str(args.gpus)]
    if args.schedule:
        cmd += ["--schedule", args.schedule]
    if args.adaptive:
        cmd += ["--adaptive"]
    cmd += ["--interval", str(args.interval)] + local_command(args.name)
    return subprocess.run(cmd, env=dict(os.environ, PYTHONPATH=ROOT)).returncode


if __name__ == "__main__":
    sys.exit(main())
