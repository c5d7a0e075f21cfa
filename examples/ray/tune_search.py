"""Hyper-parameter search over elastic trials.

With Ray:      python examples/ray/tune_search.py --samples 8
               (``--baseline`` uses Tune's FIFO scheduler with fixed-size
               trials instead of AdaptDLScheduler, for comparison)
Without Ray:   python examples/ray/tune_search.py --local --samples 3
               (runs the same trials one after the other through the
               Ray-free ``ElasticTrial``, rescaling each 1 -> 2 replicas
               half-way, to show the mechanism)

Every trial trains a small CNN on synthetic 1x28x28 images with the usual
adaptdl_b200 loop; the learning rate and momentum are the searched
parameters (reference: ``ray/adaptdl_ray/examples/hyperopt_example.py``,
``hyperopt_example_baseline.py``, ``tune_proposal.py``).
"""
import argparse
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__)))))


def train_fn(config, report):
    import torch
    import torch.nn.functional as F
    import adaptdl_b200.torch as adl
    use_cuda = torch.cuda.is_available()
    adl.init_process_group("nccl" if use_cuda else "gloo")
    device = torch.device("cuda" if use_cuda else "cpu")
    torch.manual_seed(0)
    model = torch.nn.Sequential(
        torch.nn.Conv2d(1, 8, 3), torch.nn.ReLU(), torch.nn.MaxPool2d(3),
        torch.nn.Flatten(), torch.nn.Linear(8 * 8 * 8, 10)).to(device)
    optimizer = torch.optim.SGD(model.parameters(), lr=config["lr"],
                                momentum=config["momentum"])
    net = adl.AdaptiveDataParallel(model, optimizer)
    images = torch.randn(2048, 1, 28, 28)
    labels = (images.mean(dim=(1, 2, 3)) * 40).long().clamp(-5, 4) + 5
    loader = adl.AdaptiveDataLoader(
        torch.utils.data.TensorDataset(images, labels), batch_size=64,
        shuffle=True, drop_last=True)
    loader.autoscale_batch_size(512, local_bsz_bounds=(16, 128))
    stats = adl.Accumulator()
    for epoch in adl.remaining_epochs_until(config["epochs"]):
        for x, y in loader:
            x, y = x.to(device), y.to(device)
            optimizer.zero_grad()
            out = net(x)
            loss = F.cross_entropy(out, y)
            loss.backward()
            optimizer.step()
            stats["loss"] += loss.item() * len(y)
            stats["hit"] += int((out.argmax(1) == y).sum())
            stats["n"] += len(y)
        with stats.synchronized():
            report(epoch=epoch, mean_loss=stats["loss"] / stats["n"],
                   mean_accuracy=stats["hit"] / stats["n"],
                   batch_size=loader.current_batch_size)
            stats.clear()


def sample_config(rng, epochs):
    return {"lr": 10 ** rng.uniform(-3, -0.5),
            "momentum": rng.uniform(0.1, 0.95), "epochs": epochs}


def run_local(args):
    from adaptdl_b200.ray.tune.trainable import ElasticTrial
    from adaptdl_b200.ray.tune.workers import ProcessSpawner
    root = os.path.dirname(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))))
    spawner = ProcessSpawner(extra_env={
        "PYTHONPATH": os.pathsep.join([root, os.path.join(root, "examples",
                                                          "ray")]),
        "OMP_NUM_THREADS": "1"})
    rng = random.Random(0)
    best = None
    for index in range(args.samples):
        config = sample_config(rng, args.epochs)
        trial = ElasticTrial(train_fn, config, ["local"], spawner,
                             job_id="search/{}".format(index))
        result = last = trial.step()
        while not result["done"]:
            last = result
            if last["epoch"] == args.epochs // 2 - 1 and \
                    last["num_replicas"] == 1:
                state = trial.save()                  # rescale 1 -> 2
                trial = ElasticTrial(train_fn, config, ["local", "local"],
                                     spawner,
                                     job_id="search/{}".format(index))
                trial.restore(state)
            result = trial.step()
        trial.stop()
        print("trial {}: lr={:.4f} momentum={:.2f} -> accuracy {:.3f} "
              "(finished on {} replicas)".format(
                  index, config["lr"], config["momentum"],
                  last["mean_accuracy"], last["num_replicas"]))
        if best is None or last["mean_accuracy"] > best[0]:
            best = (last["mean_accuracy"], config)
    print("best:", best)


def run_tune(args):
    import ray
    from ray import tune
    from adaptdl_b200.ray.tune import (AdaptDLScheduler,
                                       AdaptDLTrainableCreator)
    ray.init(address=args.address)
    trainable = AdaptDLTrainableCreator(train_fn,
                                        num_workers=args.workers)
    space = {"lr": tune.loguniform(1e-3, 0.3),
             "momentum": tune.uniform(0.1, 0.95), "epochs": args.epochs}
    analysis = tune.run(
        trainable, config=space, num_samples=args.samples,
        metric="mean_accuracy", mode="max",
        scheduler=None if args.baseline else AdaptDLScheduler())
    print("best config:", analysis.best_config)


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("--samples", type=int, default=4)
    parser.add_argument("--epochs", type=int, default=6)
    parser.add_argument("--workers", type=int, default=1,
                        help="initial replicas per trial (Tune mode)")
    parser.add_argument("--address", default=None, help="Ray cluster address")
    parser.add_argument("--baseline", action="store_true")
    parser.add_argument("--local", action="store_true")
    args = parser.parse_args()
    (run_local if args.local else run_tune)(args)
