"""Training function for the elastic worker-group tests (module level: the
workers are spawned processes and import it)."""
import time


def train_fn(config, report):
    import torch
    import adaptdl_b200.torch as adl
    from adaptdl_b200 import env
    adl.init_process_group("gloo")
    torch.manual_seed(0)
    model = torch.nn.Linear(4, 1)
    optimizer = torch.optim.SGD(model.parameters(), lr=config["lr"])
    net = adl.AdaptiveDataParallel(model, optimizer)
    x = torch.randn(256, 4)
    y = x @ torch.tensor([[1.0], [-2.0], [0.5], [3.0]])
    loader = adl.AdaptiveDataLoader(
        torch.utils.data.TensorDataset(x, y), batch_size=32, drop_last=True)
    for epoch in adl.remaining_epochs_until(config["epochs"]):
        for xb, yb in loader:
            optimizer.zero_grad()
            loss = torch.nn.functional.mse_loss(net(xb), yb)
            loss.backward()
            optimizer.step()
        time.sleep(config.get("pause", 0.0))
        report(epoch=epoch, loss=float(loss), replicas=env.num_replicas(),
               restarts=env.num_restarts())


def failing_fn(config, report):
    raise ValueError("boom")
