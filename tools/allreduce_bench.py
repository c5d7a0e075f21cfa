#!/usr/bin/env python
"""Fused all-reduce + statistics kernel vs NCCL all-reduce, multi-GPU.

    torchrun --nproc-per-node N tools/allreduce_bench.py [--out f.json]

For each bucket size: device time (CUDA events, max over ranks) of
(a) ``adl_allreduce_gns`` (in-place mean + both GNS statistics) in each of
its flavours (two-shot P2P, one-shot push, NVLS), (b) ``dist.all_reduce``
alone, (c) what the reference does around it for the same result (all-reduce
+ per-bucket norm kernels).

Two timings per flavour:

Everything is timed as CUDA-graph replays (no host launch cost), the way
the kernels run inside a captured training step.

``isolated``   one bucket = one optimizer step: the bucket kernel carries the
               fused finalize (statistics exchange + the step's closing peer
               barrier), i.e. a COMPLETE all-reduce whose result is visible on
               every rank when the kernel ends -- the number to hold against
               NCCL and the roofline;
``pipelined``  eight bucket kernels back to back + one closing kernel, per
               kernel: what a bucket costs in the middle of a backward pass.

Roofline: bytes each GPU must receive over NVLink / 770 GB/s (measured peer
copy; 900 nominal): two-shot 2(N-1)/N x B, NVLS (1 + 1/N) x B, one-shot
(N-1) x B.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, iters, warmup=5, stream=None):
    """Device time per call (max over ranks); events are recorded on the
    stream the kernels actually run on."""
    stream = stream or torch.cuda.current_stream()
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(True), torch.cuda.Event(True)
    start.record(stream)
    for _ in range(iters):
        fn()
    end.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([start.elapsed_time(end) / iters], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def graphed(fn, warmup=3):
    """``fn`` (kernels on the reducer's comm stream, joined to the current
    stream by events) captured into a CUDA graph: replaying it costs no
    Python / launch-API time per kernel, which at these sizes is most of an
    eager launch."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):
            fn()
        side.synchronize()
        dist.barrier()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    return graph


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--sizes-mb", default="0.0625,0.25,1,4,16,64,256")
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--ctas", default="")
    ap.add_argument("--nvls-ctas", default="32,64,96")
    ap.add_argument("--sweep", action="store_true",
                    help="also time the NVLS flavour at several grid sizes")
    args = ap.parse_args()
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from adaptdl_b200.parallel.reducer_cuda import CudaGradReducer
    from adaptdl_b200._native import (FLAVOUR_NVLS, FLAVOUR_ONESHOT,
                                      FLAVOUR_TWOSHOT)
    names = {FLAVOUR_TWOSHOT: "twoshot", FLAVOUR_ONESHOT: "oneshot",
             FLAVOUR_NVLS: "nvls"}
    dtype = torch.float32 if args.dtype == "fp32" else torch.bfloat16
    itemsize = 4 if args.dtype == "fp32" else 2
    peak = 770.0     # GB/s per direction, measured peer copy
    rows = []
    for mb in [float(x) for x in args.sizes_mb.split(",")]:
        numel = int(mb * (1 << 20) / itemsize)
        p = torch.nn.Parameter(torch.randn(numel, device=dev).to(dtype))
        if args.ctas:
            os.environ["ADAPTDL_B200_REDUCE_CTAS"] = args.ctas
        iters = 200 if mb <= 4 else 40
        variants = {}
        configs = [("auto", {}),
                   ("twoshot", {"ADAPTDL_B200_NVLS_MIN_MB": "1e9",
                                "ADAPTDL_B200_ONESHOT_KB": "0"})]
        if mb <= 8:
            configs.append(("oneshot", {"ADAPTDL_B200_NVLS_MIN_MB": "1e9",
                                        "ADAPTDL_B200_ONESHOT_KB": "1e9"}))
        nvls_grids = [int(c) for c in args.nvls_ctas.split(",")] \
            if args.sweep else [64]
        if world > 2 or args.sweep:
            configs += [("nvls%d" % c, {"ADAPTDL_B200_NVLS_MIN_MB": "0",
                                        "ADAPTDL_B200_ONESHOT_KB": "0",
                                        "ADAPTDL_B200_NVLS_MIN_WORLD": "2",
                                        "ADAPTDL_B200_NVLS_CTAS": str(c)})
                        for c in nvls_grids]
        picked = None
        for tag, env in configs:
            saved = {k: os.environ.get(k) for k in env}
            os.environ.update(env)
            red = CudaGradReducer([{"params": [p]}], world, rank,
                                  lambda: True, bucket_cap_mb=max(2 * mb, 1))
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
            arena = red.arenas[0]
            bucket = arena.buckets[0]
            arena.grad.normal_()
            nbytes = bucket.length * itemsize
            flavour = names[red._flavour[(0, bucket.index)]]
            if tag.startswith("nvls") and flavour != "nvls":
                del red, arena, bucket      # no multicast on this box
                continue
            red._k_before, red._accum_count = 0, 1

            def isolated():
                red._mark_sync_start()
                red._reduce(arena, bucket, 1.0 / world, True, last=True)
                red._finalize_step()

            def pipelined():
                for _ in range(8):
                    red._reduce(arena, bucket, 1.0 / world, True)
                isolated()
            g_iso, g_pipe = graphed(isolated), graphed(pipelined)
            t_iso = timed(g_iso.replay, iters) * 1e3
            t_pipe = timed(g_pipe.replay, max(iters // 8, 5)) * 1e3 / 9
            del g_iso, g_pipe
            variants[tag] = {"flavour": flavour, "isolated_us": t_iso,
                             "pipelined_us": t_pipe}
            if tag == "auto":
                picked = flavour
                provider, ctas = red._provider.name, red._reduce_ctas
            length = bucket.length
            del red, arena, bucket
        flat = torch.randn(length, device=dev).to(dtype)
        nccl = timed(lambda: dist.all_reduce(flat), iters) * 1e3

        def reference_like():
            local = flat.float().pow(2).sum(dtype=torch.float64)
            dist.all_reduce(flat)
            flat.div_(world)
            total = flat.float().pow(2).sum(dtype=torch.float64)
            return local, total
        ref = timed(reference_like, iters) * 1e3
        need = {"twoshot": 2.0 * (world - 1) / world * nbytes,
                "nvls": (1.0 + 1.0 / world) * nbytes,
                "oneshot": (world - 1.0) * nbytes}
        for tag, v in variants.items():
            floor_us = need[v["flavour"]] / (peak * 1e3)
            v["roofline_us"] = floor_us
            v["isolated_frac_of_770"] = floor_us / v["isolated_us"]
            v["pipelined_frac_of_770"] = floor_us / v["pipelined_us"]
        best = min(variants, key=lambda t: variants[t]["isolated_us"])
        row = {"MB": nbytes / 2 ** 20, "world": world, "dtype": args.dtype,
               "picked": picked, "best": best,
               "fused_us": variants["auto"]["isolated_us"],
               "nccl_us": nccl, "nccl_plus_norms_us": ref,
               "nccl_busbw_GBps": need["twoshot"] / nccl / 1e3,
               "provider": provider, "ctas": ctas,
               "variants": variants}
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
    if rank == 0 and args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)),
                    exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
