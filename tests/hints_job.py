"""2-replica job against the reference import names that, after a few dozen
steps, fits its performance model and reports scheduling hints to the
supervisor named by ``ADAPTDL_SUPERVISOR_URL`` (``tests/
test_reference_hints.py`` runs it under both implementations)."""

import sys


def main():
    import numpy as np
    if not hasattr(np, "int"):
        np.int, np.float = int, float
    import torch
    import adaptdl
    import adaptdl.env
    import adaptdl.torch as adl
    import adaptdl.torch._metrics as metrics

    torch.manual_seed(99)
    torch.set_num_threads(1)
    adl.init_process_group("gloo")       # master found through /discover
    features = torch.randn(2048, 12)
    targets = features.sum(dim=1, keepdim=True)
    loader = adl.AdaptiveDataLoader(
        torch.utils.data.TensorDataset(features, targets), batch_size=64,
        shuffle=True, drop_last=True)
    loader.autoscale_batch_size(1024, local_bsz_bounds=(16, 256),
                                gradient_accumulation=True)
    model = torch.nn.Linear(12, 1)
    optimizer = torch.optim.SGD(model.parameters(), lr=0.01)
    net = adl.AdaptiveDataParallel(model, optimizer)
    steps = 0
    for epoch in adl.remaining_epochs_until(2):
        for x, y in loader:
            optimizer.zero_grad()
            torch.nn.functional.mse_loss(net(x), y).backward()
            optimizer.step()
            steps += 1
    if adaptdl.env.replica_rank() == 0:
        # what the trainer does on its own every 30 s
        metrics._fit_perf_params()
        metrics._report_sched_hints()
        wait = getattr(metrics, "wait_for_report", None)
        if wait is not None:
            wait(30)
        print("GRAD {} {}".format(float(net.gns.sqr_avg()),
                                  float(net.gns.var_avg())), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
