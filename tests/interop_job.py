"""A tiny training job written against the REFERENCE import names (``import
adaptdl``), used by ``test_reference_interop.py``: run under the unmodified
reference package and under this framework's alias packages, it must be able
to resume the other one's checkpoint.

Every line printed with the ``STATE`` prefix is a JSON snapshot.
"""

import argparse
import json
import os
import signal
import sys


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--epochs", type=int, default=3)
    parser.add_argument("--stop-after-steps", type=int, default=0,
                        help="send SIGTERM to ourselves after this many "
                             "optimizer steps of this incarnation")
    args = parser.parse_args()

    import numpy as np
    if not hasattr(np, "int"):           # the reference predates numpy 1.24
        np.int, np.float = int, float
    import torch
    import adaptdl
    import adaptdl.torch as adl

    torch.manual_seed(1234)
    adl.init_process_group("gloo")
    features = torch.randn(192, 8)
    targets = features @ torch.arange(8.0).unsqueeze(1) + 0.5
    dataset = torch.utils.data.TensorDataset(features, targets)
    loader = adl.AdaptiveDataLoader(dataset, batch_size=16, shuffle=True,
                                    drop_last=True)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(),
                                torch.nn.Linear(16, 1))
    optimizer = torch.optim.SGD(model.parameters(), lr=0.02, momentum=0.9)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, 1, gamma=0.9)
    net = adl.AdaptiveDataParallel(model, optimizer, scheduler)
    stats = adl.Accumulator()

    def snapshot(tag, **extra):
        momentum = sum(float(s["momentum_buffer"].double().sum())
                       for s in optimizer.state.values()
                       if isinstance(s, dict) and "momentum_buffer" in s)
        row = {"tag": tag, "impl": adaptdl.__name__,
               "file": os.path.dirname(adaptdl.__file__),
               "params": [float(p.detach().double().sum())
                          for p in model.parameters()],
               "momentum": momentum,
               "lr": optimizer.param_groups[0]["lr"],
               "sched_epoch": scheduler.last_epoch,
               "finished_epochs": adl.finished_epochs(),
               "has_gns_state": "gns" in optimizer.state}
        row.update(extra)
        print("STATE " + json.dumps(row), flush=True)

    snapshot("start")
    steps = 0
    for epoch in adl.remaining_epochs_until(args.epochs):
        for x, y in loader:
            optimizer.zero_grad()
            loss = torch.nn.functional.mse_loss(net(x), y)
            loss.backward()
            optimizer.step()
            steps += 1
            stats["loss_sum"] += float(loss)
            stats["batches"] += 1
            snapshot("step", epoch=epoch, step=steps,
                     first_feature=float(x[0, 0]))
            if steps == args.stop_after_steps:
                os.kill(os.getpid(), signal.SIGTERM)
        scheduler.step()
        with stats.synchronized():
            snapshot("epoch_end", epoch=epoch,
                     batches=stats["batches"],
                     loss_sum=stats["loss_sum"])
            stats.clear()
    snapshot("done")


if __name__ == "__main__":
    sys.exit(main())
