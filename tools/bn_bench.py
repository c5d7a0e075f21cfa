"""Fused BatchNorm+residual+ReLU (csrc/adl_bn.cu) vs the PyTorch composition
on the ResNet-18 activation shapes (batch 128, channels-last bf16): forward
and forward+backward device time, effective bandwidth.

    python tools/bn_bench.py [--out gpurun_out/bn_bench.json]
"""

import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptdl_b200.ops import BatchNormAct2d  # noqa: E402


def time_us(fn, iters=30, warmup=5, flush=None):
    """Device time of ``fn`` replayed as a CUDA graph (no CPU launch cost)."""
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
    fn = graph.replay
    fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        total += s.elapsed_time(e)
    return total / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--iters", type=int, default=30)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.zeros(80 * 1024 * 1024, dtype=torch.float32, device=dev)
    rows = []
    for c, hw in ((64, 32), (128, 16), (256, 8), (512, 4)):
        shape = (args.batch, c, hw, hw)
        x = torch.randn(shape, device=dev).bfloat16().contiguous(
            memory_format=torch.channels_last).requires_grad_(True)
        r = torch.randn(shape, device=dev).bfloat16().contiguous(
            memory_format=torch.channels_last).requires_grad_(True)
        g = torch.randn(shape, device=dev).bfloat16().contiguous(
            memory_format=torch.channels_last)
        fused = BatchNormAct2d(c).to(dev)
        plain = torch.nn.BatchNorm2d(c).to(dev)

        def f_fused():
            return fused(x, r, True)

        def f_plain():
            return F.relu(plain(x) + r)

        def fb(fn):
            def run():
                x.grad = r.grad = None
                fn().backward(g)
            return run
        nbytes = x.numel() * 2
        row = {"shape": list(shape), "MB": nbytes / 2 ** 20}
        row["fused_fwd_us"] = time_us(f_fused, args.iters, flush=flush)
        row["torch_fwd_us"] = time_us(f_plain, args.iters, flush=flush)
        row["fused_fwd_bwd_us"] = time_us(fb(f_fused), args.iters,
                                          flush=flush)
        row["torch_fwd_bwd_us"] = time_us(fb(f_plain), args.iters,
                                          flush=flush)
        # minimum HBM traffic of the block: fwd x, res in / y out; bwd dy, y,
        # x in / dx, dres out
        row["fused_fwd_GBps_min_traffic"] = 3 * nbytes / row["fused_fwd_us"] \
            / 1e3
        row["fwd_speedup"] = row["torch_fwd_us"] / row["fused_fwd_us"]
        row["fwd_bwd_speedup"] = row["torch_fwd_bwd_us"] / \
            row["fused_fwd_bwd_us"]
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
