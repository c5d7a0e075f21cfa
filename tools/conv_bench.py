"""Per-layer device time of the ResNet-18 convolutions (cuDNN) and of the
alternatives this repo offers for the slow ones.

    python tools/conv_bench.py [--out gpurun_out/conv_bench.json]

For every distinct convolution of ResNet-18 / CIFAR shape (batch 128,
channels-last bf16) the forward, data-gradient and weight-gradient kernels
are timed SEPARATELY (``aten::convolution_backward`` with an output mask),
each replayed as a CUDA graph with an L2 flush in between, and reported with
the TFLOP/s they reach.

The step profile (``profiles/r2_validate/step_profile_resnet_bf16.log``) says
where to look: 2 x 98 us of ``implicit_gemm_strided_dgrad`` and ~117 us of
non-tensor-core stem kernels in a 2 ms step. Two PyTorch-level workarounds
(the stride-2 data gradient as four stride-1 phase convolutions; the stem on
8 zero-padded input channels) were measured in round 2 and removed: whole-step
time 2.108 ms and 1.991 ms against 1.994 ms (``profiles/r2_validate``). On a machine without a GPU the
tool runs tiny shapes on the CPU (a smoke test of the plumbing, the numbers
mean nothing).
"""

import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# (name, count in the model, C_in, C_out, input H=W, kernel, stride)
RESNET18_CONVS = [
    ("stem", 1, 3, 64, 32, 3, 1),
    ("stage1 3x3", 4, 64, 64, 32, 3, 1),
    ("stage2 3x3 s2", 1, 64, 128, 32, 3, 2),
    ("stage2 1x1 s2", 1, 64, 128, 32, 1, 2),
    ("stage2 3x3", 3, 128, 128, 16, 3, 1),
    ("stage3 3x3 s2", 1, 128, 256, 16, 3, 2),
    ("stage3 1x1 s2", 1, 128, 256, 16, 1, 2),
    ("stage3 3x3", 3, 256, 256, 8, 3, 1),
    ("stage4 3x3 s2", 1, 256, 512, 8, 3, 2),
    ("stage4 1x1 s2", 1, 256, 512, 8, 1, 2),
    ("stage4 3x3", 3, 512, 512, 4, 3, 1),
]


def timer(device, iters, warmup):
    """``measure(fn) -> microseconds``: CUDA-graph replay + events on a GPU,
    wall clock on the CPU."""
    if device.type != "cuda":
        def measure(fn):
            fn()
            start = time.perf_counter()
            for _ in range(iters):
                fn()
            return (time.perf_counter() - start) / iters * 1e6
        return measure
    flush = torch.zeros(80 * 1024 * 1024, dtype=torch.float32, device=device)

    def measure(fn):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        graph.replay()
        torch.cuda.synchronize()
        total = 0.0
        for _ in range(iters):
            flush.add_(1)                         # evict the operands
            start, end = torch.cuda.Event(True), torch.cuda.Event(True)
            start.record()
            graph.replay()
            end.record()
            torch.cuda.synchronize()
            total += start.elapsed_time(end)
        return total / iters * 1e3
    return measure


def bench_layer(spec, batch, device, dtype, measure):
    name, count, cin, cout, size, k, stride = spec
    pad = k // 2
    nhwc = torch.channels_last
    x = torch.randn(batch, cin, size, size, device=device, dtype=dtype) \
        .contiguous(memory_format=nhwc)
    w = torch.randn(cout, cin, k, k, device=device, dtype=dtype) \
        .contiguous(memory_format=nhwc)
    y = F.conv2d(x, w, None, stride, pad)
    dy = torch.randn_like(y)
    flops = 2.0 * y.numel() * cin * k * k

    def backward(mask):
        return lambda: torch.ops.aten.convolution_backward(
            dy, x, w, None, (stride, stride), (pad, pad), (1, 1), False,
            (0, 0), 1, mask)
    row = {"layer": name, "count": count, "shape": [batch, cin, size, size],
           "out_channels": cout, "kernel": k, "stride": stride,
           "gflop": flops / 1e9}
    candidates = {
        "fprop": lambda: F.conv2d(x, w, None, stride, pad),
        "dgrad": backward((True, False, False)),
        "wgrad": backward((False, True, False)),
    }
    with torch.no_grad():
        for key, fn in candidates.items():
            micros = measure(fn)
            row[key + "_us"] = round(micros, 2)
            row[key + "_tflops"] = round(flops / micros / 1e6, 1)
    return row


def main(argv=None):
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("--out")
    parser.add_argument("--batch", type=int, default=None)
    parser.add_argument("--iters", type=int, default=20)
    parser.add_argument("--warmup", type=int, default=3)
    args = parser.parse_args(argv)
    cuda = torch.cuda.is_available()
    device = torch.device("cuda:0" if cuda else "cpu")
    dtype = torch.bfloat16 if cuda else torch.float32
    batch = args.batch or (128 if cuda else 2)
    if cuda:
        torch.backends.cudnn.benchmark = True
    measure = timer(device, args.iters if cuda else 1, args.warmup)
    rows = []
    for spec in RESNET18_CONVS:
        row = bench_layer(spec, batch, device, dtype, measure)
        rows.append(row)
        print(json.dumps(row), flush=True)
    total = {key: round(sum(r[key + "_us"] * r["count"] for r in rows), 1)
             for key in ("fprop", "dgrad", "wgrad")}
    print("per step (us), all layers:", total)
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"device": str(device), "batch": batch, "rows": rows,
                       "per_step_us": total}, f, indent=1)
    return rows


if __name__ == "__main__":
    main()
