#!/usr/bin/env python
"""Elastic job used by ``tools/rescale_bench.py``: a small transformer
language model (WikiText-2 shaped synthetic token stream, BASELINE config 4)
written against the API the reference and this framework share
(``import adaptdl.torch``), so the SAME file runs under the unmodified
reference package and under this framework (whose ``adaptdl`` package is an
alias of ``adaptdl_b200``). It trains until it is preempted; the framework it
runs under does the checkpoint + exit(143) and the resume.

Life-cycle marks (wall clock) go to ``$RESCALE_MARKS/<generation>-<rank>.jsonl``.
"""
import json
import os
import time

_T0 = time.time()
_MARKS = os.environ.get("RESCALE_MARKS")
_GEN = int(os.environ.get("ADAPTDL_NUM_RESTARTS", "0") or 0)
_RANK = int(os.environ.get("ADAPTDL_REPLICA_RANK", "0") or 0)


def mark(event, **fields):
    if not _MARKS:
        return
    row = {"event": event, "t": time.time(), "generation": _GEN,
           "rank": _RANK}
    row.update(fields)
    with open(os.path.join(_MARKS, "{}-{}.jsonl".format(_GEN, _RANK)),
              "a") as f:
        f.write(json.dumps(row) + "\n")


mark("script_start")

import numpy as np  # noqa: E402

if not hasattr(np, "int"):           # the reference predates numpy 1.24
    np.int, np.float = int, float

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

mark("torch_imported")

import adaptdl.torch as adl  # noqa: E402

VOCAB, BPTT, DIM, HEADS, HIDDEN, LAYERS = 2000, 35, 64, 2, 64, 2
SAMPLES, BATCH = 40000, 32


class TinyLM(nn.Module):
    def __init__(self):
        super().__init__()
        self.embed = nn.Embedding(VOCAB, DIM)
        layer = nn.TransformerEncoderLayer(DIM, HEADS, HIDDEN, dropout=0.0,
                                           batch_first=True)
        self.encoder = nn.TransformerEncoder(layer, LAYERS)
        self.decoder = nn.Linear(DIM, VOCAB)

    def forward(self, tokens):
        return self.decoder(self.encoder(self.embed(tokens)))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    on_gpu = os.environ.get("RESCALE_DEVICE", "cpu") == "cuda" and \
        torch.cuda.is_available()
    device = torch.device("cpu")
    if on_gpu:
        device = torch.device("cuda", int(os.environ.get(
            "ADAPTDL_LOCAL_RANK", os.environ.get("ADAPTDL_REPLICA_RANK", 0))))
        torch.cuda.set_device(device)
    adl.init_process_group("nccl" if on_gpu else "gloo")
    mark("process_group", impl=adl.__name__, file=adl.__file__)
    g = torch.Generator().manual_seed(1)
    stream = torch.randint(0, VOCAB, (SAMPLES, BPTT + 1), generator=g)
    dataset = torch.utils.data.TensorDataset(stream[:, :-1], stream[:, 1:])
    loader = adl.AdaptiveDataLoader(dataset, batch_size=BATCH, shuffle=True,
                                    drop_last=True)
    model = TinyLM().to(device)
    optimizer = torch.optim.SGD(model.parameters(), lr=0.5)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, 1, gamma=0.95)
    net = adl.AdaptiveDataParallel(model, optimizer, scheduler)
    mark("model_ready", resumed_epoch=adl.current_epoch()
         if hasattr(adl, "current_epoch") else None)
    steps = tokens = 0
    for epoch in adl.remaining_epochs_until(10 ** 6):
        for x, y in loader:
            x, y = x.to(device), y.to(device)
            optimizer.zero_grad()
            loss = nn.functional.cross_entropy(
                net(x).reshape(-1, VOCAB), y.reshape(-1))
            loss.backward()
            torch.nn.utils.clip_grad_norm_(net.parameters(), 0.5)
            optimizer.step()
            steps += 1
            tokens += x.numel()          # this replica's share of the batch
            if steps == 1:
                mark("first_step", epoch=epoch, loss=loss.item())
            elif steps % 10 == 0:
                mark("progress", steps=steps, tokens=tokens)
        scheduler.step()


if __name__ == "__main__":
    main()
