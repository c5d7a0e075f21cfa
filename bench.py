#!/usr/bin/env python
"""Headline benchmark: ResNet-18 (CIFAR shape, synthetic data) adaptive
data-parallel training throughput in samples/sec, device-timed, max over
ranks -- BASELINE.json config "pytorch-cifar ResNet-18 adaptive-batch".

    python bench.py --gpus 1 --steps 50 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...    # unmodified petuum/adaptdl

Both arms run the *same user program* (the reference's
examples/pytorch-cifar/main.py training loop: SGD m=0.9 wd=5e-4 with one param
group per tensor, MultiStepLR, AdaptiveDataParallel + AdaptiveDataLoader with
autoscale_batch_size, bf16 autocast, channels-last) through each framework's
public API; only the framework under the loop differs.

Weak scaling: the per-GPU batch is fixed (default 128), global batch =
128 x N. Two kinds of timed region, each W warm-up + EXACTLY K timed steps
bracketed by barrier + synchronize, CUDA events on the launching stream, max
over ranks:

  e2e    every step copies its batch host(pinned)->device and reads the loss
         back device->host (4 B, async into pinned memory);
  value  the same loop with the batch already resident on the device.

K steps of a 2 ms step are 40 ms of device work -- too little to resolve a few
per cent. So each kind of region is repeated as R separately bracketed
windows of K steps until at least 0.5 s of device time has been measured
(R is agreed across ranks after the first window); the reported numbers are
totals over the windows (samples / time), `ms_per_step` their mean, and the
per-window times are listed under "windows".
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))

PUBLISHED_BASELINE = None      # the reference publishes no number


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--local-bsz", type=int, default=None)
    ap.add_argument("--workload", default="resnet18",
                    choices=["resnet18", "ncf", "bert", "linreg"],
                    help="resnet18 = the headline config (default); linreg "
                         "= BASELINE config 1, the CPU / gloo plumbing "
                         "config (run it with --device cpu); ncf = "
                         "small-model/latency path; bert = BERT-base MLM "
                         "bf16 (reference arm: the same model in stock "
                         "PyTorch modules, baseline/models/bert_plain.py, "
                         "under the reference's AdaptiveDataParallel)")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--reducer", default="auto")
    ap.add_argument("--bucket-cap-mb", type=float, default=None,
                    help="own arm: gradient bucket cap (default: the "
                         "reducer's own, 25 MB)")
    ap.add_argument("--param-dtype", default="bf16", choices=["bf16", "fp32"],
                    help="own arm: store conv/linear weights in bf16 with "
                         "fp32 masters inside the fused optimizer (default) "
                         "or keep fp32 weights + autocast casts")
    ap.add_argument("--no-graph", action="store_true",
                    help="own arm: eager step instead of the CUDA-graph step")
    ap.add_argument("--min-timed-ms", type=float, default=500.0,
                    help="repeat the K-step window until this much device "
                         "time has been measured per region (0 = one window)")
    ap.add_argument("--max-windows", type=int, default=40)
    ap.add_argument("--no-fp32-variant", action="store_true",
                    help="own arm: skip the extra fp32-parameter / "
                         "fp32-gradient measurement")
    return ap.parse_args()


def setup_env(args):
    # torchrun pins OMP_NUM_THREADS=1 for N>1; do the same for the plain
    # `python bench.py` (N=1) launch so every point of the scaling curve -- and
    # both arms -- run with the same host threading
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(1)))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ["ADAPTDL_NUM_REPLICAS"] = str(world)
    os.environ["ADAPTDL_REPLICA_RANK"] = str(rank)
    os.environ["ADAPTDL_NUM_NODES"] = "1"
    os.environ["ADAPTDL_MASTER_ADDR"] = os.environ.get("MASTER_ADDR",
                                                       "127.0.0.1")
    base_port = int(os.environ.get("MASTER_PORT", "29400"))
    os.environ["ADAPTDL_MASTER_PORT"] = str(base_port + 1)
    os.environ.pop("ADAPTDL_CHECKPOINT_PATH", None)
    # both frameworks rendezvous over their own tcp:// store (port agreed on
    # their control plane); torchrun's agent-store flag would stop rank 0
    # from hosting it
    os.environ.pop("TORCHELASTIC_USE_AGENT_STORE", None)
    return rank, world, local_rank


class SyntheticCIFAR(object):
    """CIFAR-10-shaped random data held in pinned host memory, fetched a
    whole batch at a time (``__getitems__``) so the Python data path is not
    what is being measured. Plain user code: used by BOTH arms."""

    def __init__(self, size, pin):
        import torch
        g = torch.Generator().manual_seed(1234)
        self.x = torch.randn(size, 3, 32, 32, generator=g)
        self.y = torch.randint(0, 10, (size,), generator=g)
        if pin:
            self.x, self.y = self.x.pin_memory(), self.y.pin_memory()
        self.pin = pin

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], self.y[i]

    def __getitems__(self, idx):
        import torch
        idx = torch.as_tensor(idx)
        x = torch.empty((len(idx),) + tuple(self.x.shape[1:]),
                        dtype=self.x.dtype, pin_memory=self.pin)
        y = torch.empty((len(idx),), dtype=self.y.dtype, pin_memory=self.pin)
        torch.index_select(self.x, 0, idx, out=x)
        torch.index_select(self.y, 0, idx, out=y)
        return x, y


def identity_collate(batch):
    return batch


class ClockSampler(object):
    """SM clock and throttle reasons of one GPU, sampled ONLY while a timed
    window is open (``begin`` ... ``end``). NVML in-process (every 25 ms,
    plus one read from the timing loop itself in the middle of each window;
    a window of 20 ResNet steps is 40 ms, far too short for an
    ``nvidia-smi`` start-up); if ``pynvml`` is unusable, one long-running
    ``nvidia-smi -lms 50`` whose lines are kept while a window is open."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,"
             "clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* bits
    REASON_BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown",
                   0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                   0x80: "hw_power_brake_slowdown"}
    # sparse on purpose: the queries go through the driver of the GPU that
    # is being timed, from a thread of the process that launches its work
    PERIOD_S = 0.025

    def __init__(self, enabled, gpu_index, uuid=None):
        import threading
        self.rows = []               # (sm_mhz, sm_max_mhz, reasons)
        self.source = None
        self._open = False
        self._halt = threading.Event()
        self._thread = None
        self._proc = None
        self._read_once = None
        self._rows_at_begin = 0
        if not enabled:
            return
        target = self._nvml(gpu_index, uuid) or self._smi(gpu_index)
        if target is not None:
            self._thread = threading.Thread(target=target, daemon=True)
            self._thread.start()

    def _nvml(self, gpu_index, uuid):
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = None
            if uuid:
                try:
                    handle = pynvml.nvmlDeviceGetHandleByUUID(
                        uuid if isinstance(uuid, bytes) else uuid.encode())
                except Exception:  # noqa: BLE001
                    handle = None
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(
                handle, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(
                pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            reasons_fn(handle)       # fail here, not in the thread
        except Exception:  # noqa: BLE001
            return None
        self.source = "nvml"

        def read_once():
            try:
                sm = float(pynvml.nvmlDeviceGetClockInfo(
                    handle, pynvml.NVML_CLOCK_SM))
                mask = int(reasons_fn(handle))
            except Exception:  # noqa: BLE001
                return None
            return (sm, sm_max, frozenset(
                name for bit, name in self.REASON_BITS.items()
                if mask & bit))
        self._read_once = read_once

        def loop():
            while not self._halt.wait(self.PERIOD_S):
                if not self._open:
                    continue
                row = read_once()
                if row is not None and self._open:
                    self.rows.append(row)
        return loop

    def _smi(self, gpu_index):
        try:
            self._proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self._proc = None
            return None
        self.source = "nvidia-smi"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]

        def loop():
            for line in self._proc.stdout:
                if self._halt.is_set():
                    break
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 8 or not self._open:
                    continue
                try:
                    sm, mx = float(parts[1]), float(parts[2])
                except ValueError:
                    continue
                self.rows.append((sm, mx, frozenset(
                    name for name, val in zip(names, parts[4:8])
                    if val.lower().startswith("active"))))
        return loop

    def begin(self):
        self._rows_at_begin = len(self.rows)
        self._open = True

    def poke(self):
        """Called by the timing loop in the middle of a window: if the
        sampling thread has not been scheduled since ``begin`` (a window can
        be as short as 40 ms), read the clocks once from this thread."""
        if self._open and self._read_once is not None and \
                len(self.rows) == self._rows_at_begin:
            row = self._read_once()
            if row is not None:
                self.rows.append(row)

    def end(self):
        self._open = False

    def close(self):
        """Stop sampling; the summary of every sample taken inside a window
        (``None`` without samples)."""
        self._open = False
        self._halt.set()
        if self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self._proc.kill()
        if self._thread is not None:
            self._thread.join(timeout=5)
        if not self.rows:
            return None
        reasons = set()
        for _, _, r in self.rows:
            reasons |= r
        return {"sm_mhz": statistics.median(r[0] for r in self.rows),
                "sm_min_mhz": min(r[0] for r in self.rows),
                "sm_max_mhz": max(r[1] for r in self.rows),
                "reasons": sorted(reasons), "samples": len(self.rows),
                "source": self.source}


class SyntheticNCF(object):
    """MovieLens-1M shaped implicit feedback (6040 users x 3706 items)."""
    USERS, ITEMS = 6040, 3706

    def __init__(self, size, pin):
        import torch
        g = torch.Generator().manual_seed(1234)
        self.u = torch.randint(0, self.USERS, (size,), generator=g)
        self.i = torch.randint(0, self.ITEMS, (size,), generator=g)
        self.y = torch.randint(0, 2, (size,), generator=g).float()
        if pin:
            self.u, self.i, self.y = (t.pin_memory()
                                      for t in (self.u, self.i, self.y))

    def __len__(self):
        return self.u.shape[0]

    def __getitem__(self, k):
        return self.u[k], self.i[k], self.y[k]

    def __getitems__(self, idx):
        import torch
        idx = torch.as_tensor(idx)
        return self.u[idx], self.i[idx], self.y[idx]


class SyntheticMLM(object):
    """Token sequences for masked-LM pre-training (BERT vocabulary size)."""
    NTOKEN, SEQ = 28996, 128

    def __init__(self, size, pin):
        import torch
        g = torch.Generator().manual_seed(1234)
        self.x = torch.randint(2, self.NTOKEN, (size, self.SEQ), generator=g)
        self.y = self.x.clone()
        mask = torch.rand(size, self.SEQ, generator=g) < 0.15
        self.y[~mask] = -100
        self.x[mask] = 1
        if pin:
            self.x, self.y = self.x.pin_memory(), self.y.pin_memory()

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, k):
        return self.x[k], self.y[k]

    def __getitems__(self, idx):
        import torch
        idx = torch.as_tensor(idx)
        return self.x[idx], self.y[idx]


class SyntheticPoly(object):
    """The reference's ``examples/linear_regression`` problem (BASELINE
    config 1, the CPU / gloo plumbing config): features [x, x^2, x^3, x^4] of
    a normal x, target a fixed linear function of them plus noise."""
    DEGREE = 4

    def __init__(self, size, pin):
        import torch
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(size, generator=g).unsqueeze(1)
        self.x = torch.cat([x ** i for i in range(1, self.DEGREE + 1)], 1)
        w = torch.randn(self.DEGREE, 1, generator=g) * 5
        self.y = self.x.mm(w) + 0.25 * torch.randn(size, 1, generator=g)
        if pin:
            self.x, self.y = self.x.pin_memory(), self.y.pin_memory()

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, k):
        return self.x[k], self.y[k]

    def __getitems__(self, idx):
        import torch
        idx = torch.as_tensor(idx)
        return self.x[idx], self.y[idx]


class Workload(object):
    """One benchmark configuration: dataset, model, optimizer, loss."""

    def __init__(self, name, own):
        self.name, self.own = name, own
        self.channels_last = name == "resnet18"
        self.default_local_bsz = {"resnet18": 128, "ncf": 256,
                                  "bert": 32, "linreg": 64}[name]
        self.unit = {"resnet18": "samples/s", "ncf": "samples/s",
                     "bert": "sequences/s", "linreg": "samples/s"}[name]

    def dataset(self, size, pin):
        return {"resnet18": SyntheticCIFAR, "ncf": SyntheticNCF,
                "bert": SyntheticMLM, "linreg": SyntheticPoly}[self.name](
                    size, pin)

    def model(self):
        if self.name == "resnet18":
            if self.own:
                from adaptdl_b200.models import resnet18
                return resnet18()
            from cifar_models.resnet import ResNet18
            return ResNet18()
        if self.name == "linreg":
            import torch
            return torch.nn.Linear(SyntheticPoly.DEGREE, 1)
        if self.name == "ncf":
            if self.own:
                from adaptdl_b200.models import NCF
            else:
                from ncf_model import NCF
            return NCF(SyntheticNCF.USERS, SyntheticNCF.ITEMS, 32, 3, 0.0,
                       "NeuMF-end")
        if self.own:
            from adaptdl_b200.models import bert_base_mlm
        else:
            # the reference's examples/BERT/model.py needs torchtext.nn and
            # a pre-2.0 nn.TransformerEncoder: BASELINE.md section 2 defines
            # the baseline as the same model (stock PyTorch modules) under
            # the reference's unmodified AdaptiveDataParallel
            from models.bert_plain import bert_base_mlm
        return bert_base_mlm(SyntheticMLM.NTOKEN, max_len=SyntheticMLM.SEQ)

    def optimizer(self, model):
        import torch
        if self.name == "linreg":
            opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9,
                                  weight_decay=5e-4)
            return opt, torch.optim.lr_scheduler.MultiStepLR(opt, [30, 45],
                                                             0.1)
        if self.name == "resnet18":
            opt = torch.optim.SGD([{"params": [p]}
                                   for p in model.parameters()],
                                  lr=0.1, momentum=0.9, weight_decay=5e-4)
            sched = torch.optim.lr_scheduler.MultiStepLR(opt, [30, 45], 0.1)
        elif self.name == "ncf":
            opt = torch.optim.Adam(model.parameters(), lr=1e-3)
            sched = None
        else:
            opt = torch.optim.AdamW(model.parameters(), lr=1e-4,
                                    weight_decay=0.01)
            sched = None
        return opt, sched

    def loss_fn(self):
        import torch
        if self.name == "resnet18":
            ce = torch.nn.CrossEntropyLoss()
            return lambda net, x, y: ce(net(x), y)
        if self.name == "linreg":
            return lambda net, x, y: torch.nn.functional.smooth_l1_loss(
                net(x), y)
        if self.name == "ncf":
            bce = torch.nn.BCEWithLogitsLoss()
            return lambda net, u, i, y: bce(net(u, i), y)
        ce = torch.nn.CrossEntropyLoss(ignore_index=-100)
        return lambda net, x, y: ce(
            net(x).view(-1, SyntheticMLM.NTOKEN), y.view(-1))

    def l2_note(self):
        if self.name == "linreg":
            return ("framework-overhead config (5 parameters): nothing to "
                    "flush; a fresh batch every step")
        if self.name == "ncf":
            return ("latency-bound config: parameters + optimizer state "
                    "(26 MB) fit the 126 MB L2, no flush between steps; a "
                    "fresh batch every e2e step")
        return ("working set > L2 (the step's activations + weights + "
                "optimizer state exceed 126 MB) and a fresh batch every "
                "e2e step")

    def describe(self):
        return {
            "resnet18": ("ResNet-18 (pytorch-cifar, 3x32x32, 10 classes, "
                         "random init)",
                         "SGD m=0.9 wd=5e-4, one param group per tensor "
                         "(62 GNS groups), AdaScale LR"),
            "ncf": ("NeuMF-end NCF (MovieLens-1M shape, 1.6 M params, "
                    "random init)", "Adam lr=1e-3, AdamScale LR"),
            "linreg": ("linear regression on 4 polynomial features "
                       "(examples/linear_regression)",
                       "SGD m=0.9 wd=5e-4, MultiStepLR, AdaScale LR"),
            "bert": ("BERT-base MLM (768/3072/12L/12H, seq 128, untied "
                     "head, random init)", "AdamW lr=1e-4, AdamScale LR"),
        }[self.name]


def warmup_steps(args):
    """Untimed steps actually run before each timed region: at least the
    requested ``--warmup``, and never fewer than 6 so that cuDNN autotuning
    and the own arm's CUDA-graph capture (3 eager steps, then the capturing
    step) are over before the clock starts -- same for both arms."""
    return max(args.warmup, 6)


def build_program(args, adl, device, world, workload, param_dtype=None,
                  name=None):
    """The user program (identical for both arms)."""
    import torch
    param_dtype = param_dtype or args.param_dtype
    local_bsz = args.local_bsz or workload.default_local_bsz
    global_bsz = local_bsz * world
    # one pass over the dataset = one timed window (W warm-up + K steps)
    total_steps = warmup_steps(args) + args.steps + 2
    dataset = workload.dataset(global_bsz * total_steps,
                               pin=device.type == "cuda")
    loader = adl.AdaptiveDataLoader(dataset, batch_size=global_bsz,
                                    shuffle=True, drop_last=True,
                                    collate_fn=identity_collate)
    loader.autoscale_batch_size(32 * global_bsz,
                                local_bsz_bounds=(min(32, local_bsz), 1024),
                                gradient_accumulation=False)
    model = workload.model().to(device)
    if device.type == "cuda" and workload.channels_last:
        model = model.to(memory_format=torch.channels_last)
    if workload.own and device.type == "cuda" and param_dtype == "bf16" \
            and not args.no_graph:
        adl.mixed_precision_params(model)
    optimizer, scheduler = workload.optimizer(model)
    kwargs = {}
    if name is not None:
        kwargs["name"] = name
    if workload.own and args.reducer != "auto":
        kwargs["reducer"] = args.reducer
    if workload.own and args.bucket_cap_mb:
        kwargs["bucket_cap_mb"] = args.bucket_cap_mb
    if workload.name == "ncf":
        kwargs["find_unused_parameters"] = True
    net = adl.AdaptiveDataParallel(model, optimizer, scheduler, **kwargs)
    return dataset, loader, net, optimizer, global_bsz


class Region(object):
    """Timed windows of one kind ("e2e" or "device") for one program."""

    def __init__(self, name):
        self.name = name
        self.ms = []            # device ms per window (this rank)
        self.wall_ms = []
        self.samples = 0        # local samples inside timed windows
        self.target = None      # windows to run (agreed after the first)
        self.launches = 0

    def done(self):
        return self.target is not None and len(self.ms) >= self.target


def run_program(args, adl, device, rank, world, local_rank, workload,
                param_dtype, regions, name=None, sample_clocks=True):
    """Build the user program once and run the requested regions on it.
    Returns a dict of per-region totals plus bookkeeping."""
    import math
    import torch
    import torch.distributed as dist
    own = workload.own
    dataset, loader, net, optimizer, global_bsz = build_program(
        args, adl, device, world, workload, param_dtype, name)
    loss_fn = workload.loss_fn()
    cl = device.type == "cuda" and workload.channels_last
    W, K = warmup_steps(args), args.steps
    autocast = device.type == "cuda"
    loss_host = torch.zeros(W + K + 8, dtype=torch.float32)
    if device.type == "cuda":
        loss_host = loss_host.pin_memory()
    if own:
        from adaptdl_b200.ops import launch_count as ops_launch_count

    def barrier():
        if world > 1:
            dist.barrier()
        if device.type == "cuda":
            torch.cuda.synchronize()

    trainer = None
    if own:
        # the framework's own step API: whole-step CUDA graph on top of the
        # device-resident estimator + fused optimizer (eager with --no-graph)
        trainer = adl.GraphedTrainStep(
            net, optimizer, loss_fn,
            autocast_dtype=torch.bfloat16 if autocast else None,
            enabled=not args.no_graph, channels_last=cl)

    def train_step(batch, slot, read_back):
        if trainer is not None:
            loss = trainer(*batch)
        else:
            optimizer.zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16,
                                enabled=autocast):
                loss = loss_fn(net, *batch)
            loss.backward()
            optimizer.step()
        if read_back:
            loss_host[slot].copy_(loss.detach(), non_blocking=True)
        return loss

    def on_device(t, blocking):
        t = t.to(device, non_blocking=not blocking)
        if cl and t.dim() == 4:
            t = t.contiguous(memory_format=torch.channels_last)
        return t

    def launches_now():
        return net.reducer.launches + ops_launch_count() if own else 0

    todo = [Region(r) for r in regions]
    gpu_uuid = None
    if device.type == "cuda":
        try:
            gpu_uuid = "GPU-" + str(
                torch.cuda.get_device_properties(device).uuid)
        except Exception:  # noqa: BLE001 - older torch: index it is
            gpu_uuid = None
    sampler = ClockSampler(
        sample_clocks and rank == 0 and device.type == "cuda", local_rank,
        gpu_uuid)
    resident = None
    h2d_bytes = 0
    epochs = adl.remaining_epochs_until(10 ** 6)
    for _ in epochs:
        region = next((r for r in todo if not r.done()), None)
        if region is None:
            epochs.close()       # leave the epoch loop cleanly
            break
        # the first window of a region gets the full warm-up (cuDNN
        # autotuning, graph capture); later ones the recipe's minimum
        w = W if not region.ms else 3
        t_wall = ev0 = ev1 = None
        n0 = 0
        for step, batch in enumerate(loader):
            if step == w:
                barrier()
                if region.name == "device":
                    sampler.begin()
                if device.type == "cuda":
                    ev0 = torch.cuda.Event(enable_timing=True)
                    ev1 = torch.cuda.Event(enable_timing=True)
                    ev0.record()
                t_wall = time.perf_counter()
                n0 = launches_now()
            if step == w + K // 2:
                sampler.poke()
            if step == w + K:
                if device.type == "cuda":
                    ev1.record()
                    torch.cuda.synchronize()
                    ms = ev0.elapsed_time(ev1)
                else:
                    ms = (time.perf_counter() - t_wall) * 1e3
                wall_ms = (time.perf_counter() - t_wall) * 1e3
                sampler.end()
                barrier()
                region.ms.append(ms)
                region.wall_ms.append(wall_ms)
                region.launches += launches_now() - n0
                break
            timed = step >= w
            if region.name == "e2e":
                h2d_bytes = sum(t.numel() * t.element_size() for t in batch)
                if timed:
                    region.samples += int(batch[0].shape[0])
                if trainer is not None:
                    # pinned host tensors go straight into the step
                    # (copied H2D into the graph's static inputs)
                    train_step(batch, step, read_back=True)
                else:
                    train_step([on_device(t, False) for t in batch],
                               step, read_back=True)
            else:
                if resident is None or \
                        resident[0].shape[0] != batch[0].shape[0]:
                    resident = [on_device(t, True) for t in batch]
                if timed:
                    region.samples += int(resident[0].shape[0])
                train_step(resident, step, read_back=False)
        if region.target is None:
            # agree on the number of windows (max over ranks of the first)
            first = torch.tensor([region.ms[0]], dtype=torch.float64,
                                 device=device)
            if world > 1:
                dist.all_reduce(first, op=dist.ReduceOp.MAX)
            want = math.ceil(args.min_timed_ms / max(float(first), 1e-3))
            region.target = max(1, min(args.max_windows, want))
    if device.type == "cuda":
        torch.cuda.synchronize()
    clocks = sampler.close()
    assert bool(torch.isfinite(loss_host[:W + K]).all()), "non-finite loss"

    out = {"global_bsz": global_bsz, "clocks": clocks,
           "h2d_bytes": h2d_bytes, "net": net, "trainer": trainer}
    for region in todo:
        n = len(region.ms)
        # per window: max over ranks; totals over windows
        per_window = torch.tensor(region.ms + region.wall_ms,
                                  dtype=torch.float64, device=device)
        samples = torch.tensor([float(region.samples)], dtype=torch.float64,
                               device=device)
        if world > 1:
            dist.all_reduce(per_window, op=dist.ReduceOp.MAX)
            dist.all_reduce(samples, op=dist.ReduceOp.SUM)
        per_window = per_window.tolist()
        dev_ms, wall_ms = per_window[:n], per_window[n:]
        out[region.name] = {
            "windows_ms": dev_ms, "total_ms": sum(dev_ms),
            "wall_total_ms": sum(wall_ms), "steps": n * K,
            "samples": float(samples), "launches": region.launches,
        }
    return out


def step_profile(own):
    """What the framework's OWN step profiler (the numbers its goodput model
    is fitted to) booked during the run, rank 0: mean optimizer-step time and
    mean "sync" time = last gradient ready -> gradients reduced, i.e. the
    communication a step could not hide under backward. Own arm: %globaltimer
    stamps of the kernels; reference arm: its CUDA events. Bookkeeping only --
    never allowed to break the benchmark line."""
    try:
        if own:
            from adaptdl_b200.torch import _metrics
        else:
            from adaptdl.torch import _metrics
        steps = step_s = sync_s = 0.0
        for row in list(_metrics._metrics_state().profile.values()):
            count = row.get("optim_count", 0)
            if not count:
                continue
            steps += count
            step_s += row.get("optim_step_time", 0.0)
            sync_s += row.get("optim_sync_time", 0.0)
        if not steps:
            return None
        return {"steps": int(steps),
                "step_ms": round(1e3 * step_s / steps, 4),
                "sync_ms": round(1e3 * sync_s / steps, 4)}
    except Exception as exc:  # noqa: BLE001
        return {"error": str(exc)[:200]}


def run(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    own = args.impl == "own"
    if args.device == "cuda" and not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (use --device cpu "
                         "only for plumbing checks)")
    device = torch.device("cuda", local_rank) if args.device == "cuda" \
        else torch.device("cpu")
    if device.type == "cuda":
        torch.cuda.set_device(device)
        torch.backends.cudnn.benchmark = True

    if own:
        sys.path.insert(0, ROOT)
        import adaptdl_b200.torch as adl
    else:
        import numpy as np
        if not hasattr(np, "int"):       # numpy >= 1.24 dropped the aliases
            np.int, np.float = int, float
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        sys.path.insert(0, os.path.join(ROOT, "baseline", "shims"))
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref",
                                        "_ref_examples"))
        import adaptdl.torch as adl

    adl.init_process_group("nccl" if device.type == "cuda" else "gloo")
    workload = Workload(args.workload, own)
    K = args.steps
    bf16_params = (own and args.param_dtype == "bf16" and not args.no_graph
                   and device.type == "cuda")
    main = run_program(args, adl, device, rank, world, local_rank, workload,
                       args.param_dtype, ("e2e", "device"))
    net, trainer = main["net"], main["trainer"]
    global_bsz = main["global_bsz"]
    fp32_variant = None
    if bf16_params and not args.no_fp32_variant:
        # like-for-like precision on the gradient path: fp32 parameters and
        # fp32 gradients (what the reference arm reduces), same engine
        # an extra: whatever happens in it must not cost the headline line
        # that has already been measured
        try:
            extra = run_program(args, adl, device, rank, world, local_rank,
                                workload, "fp32", ("device",),
                                name="fp32-variant", sample_clocks=False)
            d = extra["device"]
            fp32_variant = {
                "param_dtype": "fp32", "grad_dtype": "fp32",
                "value": d["samples"] / (d["total_ms"] / 1e3),
                "ms_per_step": d["total_ms"] / d["steps"],
                "steps_timed": d["steps"]}
        except Exception as exc:  # noqa: BLE001
            fp32_variant = {"error": "{}: {}".format(type(exc).__name__,
                                                     str(exc)[:300])}

    dev, e2e = main["device"], main["e2e"]
    value = dev["samples"] / (dev["total_ms"] / 1e3)
    e2e_value = e2e["samples"] / (e2e["total_ms"] / 1e3)
    if rank == 0:
        line = {
            "metric": ("samples/sec (device-timed, max over ranks)"
                       if device.type == "cuda" else
                       "samples/sec (host-timed, max over ranks)"),
            "value": value, "unit": workload.unit, "n_gpus": world,
            "steps": K, "warmup": args.warmup,
            "warmup_run": warmup_steps(args),
            "ms_per_step": dev["total_ms"] / dev["steps"],
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": (value / PUBLISHED_BASELINE
                            if PUBLISHED_BASELINE else None),
            "dtype": "bf16" if device.type == "cuda" else "fp32",
            "data": "synthetic",
            "impl": args.impl,
            # the WORKLOAD (identical strings in both arms) ...
            "config": {
                "workload": workload.name,
                "model": workload.describe()[0],
                "global_batch": global_bsz,
                "local_batch": global_bsz // world,
                "seq_len": (SyntheticMLM.SEQ if workload.name == "bert"
                            else None),
                "parallelism": "dp{}".format(world),
                "optimizer": workload.describe()[1],
                "adaptive": "autoscale_batch_size(max=32x, local 32..1024)",
                "compute": (("channels_last, " if workload.channels_last
                             else "") + "bf16 autocast"
                            if device.type == "cuda" else "fp32, CPU / gloo"),
                "l2": workload.l2_note(),
            },
            # ... and how THIS arm implements it
            "param_dtype": ("bf16 weights + fp32 masters and fp32 optimizer "
                            "state in the fused optimizer" if bf16_params
                            else "fp32"),
            "grad_dtype": "bf16" if bf16_params else "fp32",
            "step_mode": ("cuda_graph" if (own and trainer is not None
                                           and trainer.replays > 0)
                          else "eager"),
            "windows": {"per_window_steps": K,
                        "device_ms": dev["windows_ms"],
                        "e2e_ms": e2e["windows_ms"]},
            "steps_timed": dev["steps"],
            "e2e": {"value": e2e_value, "unit": workload.unit,
                    "ms_per_step": e2e["total_ms"] / e2e["steps"],
                    "steps_timed": e2e["steps"],
                    "h2d_bytes_per_step": main["h2d_bytes"],
                    "d2h_bytes_per_step": 4},
            "wall_ms_per_step": {
                "device": dev["wall_total_ms"] / dev["steps"],
                "e2e": e2e["wall_total_ms"] / e2e["steps"]},
            "gpu_launches": (round(dev["launches"] * K / dev["steps"])
                             if own else None),
            "gpu_launches_e2e": (round(e2e["launches"] * K / e2e["steps"])
                                 if own else None),
            "clocks": main["clocks"],
        }
        if fp32_variant is not None:
            line["fp32_grad_variant"] = fp32_variant
        if own:
            red = net.reducer
            line["reducer"] = type(red).__name__
            line["graph_replays"] = trainer.replays
            line["eager_steps"] = trainer.eager_steps
            line["device_engine"] = net.engine is not None
            prov = getattr(red, "_provider", None)
            line["symmetric_memory"] = getattr(prov, "name", None)
            line["buckets"] = getattr(red, "num_buckets", None)
            line["nvls_launches"] = getattr(red, "nvls_launches", 0)
            line["oneshot_launches"] = getattr(red, "oneshot_launches", 0)
            from adaptdl_b200.torch import _metrics
            timer = _metrics.device_timer()
            line["device_timed_profile_steps"] = \
                timer.booked if timer is not None else 0
        line["step_profile"] = step_profile(own)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    if dist.is_initialized():
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank, world, local_rank = setup_env(args)
    if args.impl == "reference":
        ref = os.path.join(ROOT, "baseline", "_ref", "adaptdl")
        if not os.path.isdir(ref):
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable":
                                  "baseline/_ref not installed (run "
                                  "baseline/install_reference.sh)"}))
            return
    run(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
