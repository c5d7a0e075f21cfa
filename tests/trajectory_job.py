"""Deterministic multi-replica job against the reference import names; prints
the adaptive-training quantities after every step (``tests/
test_reference_trajectory.py`` runs it under both implementations)."""

import argparse
import json
import sys


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--steps", type=int, default=30)
    parser.add_argument("--rule", default="adascale")
    parser.add_argument("--optimizer", default="sgd")
    parser.add_argument("--batch-size", type=int, default=64)
    parser.add_argument("--accumulation", action="store_true")
    parser.add_argument("--shuffle", action="store_true")
    parser.add_argument("--autoscale", action="store_true")
    args = parser.parse_args()

    import numpy as np
    if not hasattr(np, "int"):
        np.int, np.float = int, float
    import torch
    import adaptdl
    import adaptdl.env
    import adaptdl.torch as adl
    from adaptdl.torch import scaling_rules

    torch.manual_seed(4321)
    torch.set_num_threads(1)
    adl.init_process_group("gloo")
    rank = adaptdl.env.replica_rank()
    features = torch.randn(4096, 12)
    weights = torch.linspace(-1.0, 1.0, 12).unsqueeze(1)
    targets = features @ weights + 0.3 * torch.randn(4096, 1)
    dataset = torch.utils.data.TensorDataset(features, targets)
    loader = adl.AdaptiveDataLoader(dataset, batch_size=args.batch_size,
                                    shuffle=args.shuffle,
                                    drop_last=True)
    if args.autoscale:
        loader.autoscale_batch_size(
            512, local_bsz_bounds=(16, 128),
            gradient_accumulation=args.accumulation)
    model = torch.nn.Sequential(torch.nn.Linear(12, 24), torch.nn.ReLU(),
                                torch.nn.Linear(24, 1))
    if args.optimizer == "sgd":
        optimizer = torch.optim.SGD(model.parameters(), lr=0.01,
                                    momentum=0.9, weight_decay=1e-4)
    else:
        optimizer = torch.optim.AdamW(model.parameters(), lr=0.01)
    rule = {"adascale": scaling_rules.AdaScale,
            "adamscale": scaling_rules.AdamScale,
            "sqrt": scaling_rules.SqrtScale,
            "linear": scaling_rules.LinearScale,
            "default": lambda: None}[args.rule]()
    net = adl.AdaptiveDataParallel(model, optimizer, scaling_rule=rule)

    steps = 0
    stats = adl.Accumulator()
    for epoch in adl.remaining_epochs_until(1000):
        for x, y in loader:
            optimizer.zero_grad()
            loss = torch.nn.functional.mse_loss(net(x), y)
            loss.backward()
            optimizer.step()
            stats["loss_sum"] += float(loss) * len(x)
            stats["samples"] += len(x)
            stats.update({"batches": 1})
            is_step = not args.accumulation or \
                not loader._elastic.is_accum_step()
            if not is_step:
                continue
            steps += 1
            if rank == 0:
                print("STATE " + json.dumps({
                    "step": steps, "impl": adaptdl.__name__,
                    "loss": float(loss),
                    "bsz": int(loader.current_batch_size),
                    "local_bsz": int(loader.current_local_bsz),
                    "accum": int(loader.accumulation_steps),
                    "gain": float(net.gain),
                    "sqr_avg": float(net.gns.sqr_avg()),
                    "var_avg": float(net.gns.var_avg()),
                    "progress": float(net.gns.get_progress()),
                    "lr": [float(g["lr"]) for g in optimizer.param_groups],
                    "params": [float(p.detach().double().sum())
                               for p in model.parameters()],
                    "first": float(x[0, 0])}), flush=True)
            if steps >= args.steps:
                return 0
        with stats.synchronized():
            if rank == 0:
                print("EPOCH " + json.dumps(dict(stats, epoch=epoch)),
                      flush=True)
            stats.clear()
    return 0


if __name__ == "__main__":
    sys.exit(main())
