#!/usr/bin/env python
"""Fused all-reduce + statistics kernel vs NCCL all-reduce, multi-GPU.

    torchrun --nproc-per-node N tools/allreduce_bench.py [--out f.json]

For each bucket size: device time (CUDA events, max over ranks) of
(a) ``adl_allreduce_gns`` (in-place mean + both GNS statistics), (b)
``dist.all_reduce`` alone, (c) what the reference does around it for the same
result (all-reduce + per-bucket norm kernels). Bus bandwidth =
``2 (N-1)/N * bytes / time``; roofline = measured 770 GB/s per direction.
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--sizes-mb", default="0.0625,0.25,1,4,16,64,256")
    ap.add_argument("--ctas", default="")
    ap.add_argument("--sweep", action="store_true",
                    help="also time the NVLS flavour at 32/64/96 CTAs")
    args = ap.parse_args()
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from adaptdl_b200.parallel.reducer_cuda import CudaGradReducer
    rows = []
    for mb in [float(x) for x in args.sizes_mb.split(",")]:
        numel = int(mb * (1 << 20) / 4)
        p = torch.nn.Parameter(torch.randn(numel, device=dev))
        if args.ctas:
            os.environ["ADAPTDL_B200_REDUCE_CTAS"] = args.ctas
        iters = 200 if mb <= 4 else 40
        variants = {}
        # the fused kernel in its P2P flavour and its NVLS (multimem)
        # flavour at several grid sizes; "fused" = what the reducer picks
        configs = [("fused", {}), ("p2p", {"ADAPTDL_B200_NVLS_MIN_MB": "1e9"})]
        if args.sweep:
            configs += [("nvls%d" % c, {"ADAPTDL_B200_NVLS_MIN_MB": "0",
                                        "ADAPTDL_B200_NVLS_CTAS": str(c)})
                        for c in (32, 64, 96)]
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
            nbytes = bucket.length * 4
            variants[tag] = timed(
                lambda: red._reduce(arena, bucket, 1.0 / world, True),
                iters, stream=red._comm) * 1e3
            if tag == "fused":
                used_nvls = bool(getattr(red, "nvls_launches", 0))
                provider, ctas = red._provider.name, red._reduce_ctas
            length = bucket.length
            del red, arena, bucket
        fused = variants["fused"] / 1e3
        flat = torch.randn(length, device=dev)
        nccl = timed(lambda: dist.all_reduce(flat), iters)

        def reference_like():
            local = flat.float().pow(2).sum(dtype=torch.float64)
            dist.all_reduce(flat)
            flat.div_(world)
            total = flat.float().pow(2).sum(dtype=torch.float64)
            return local, total
        ref = timed(reference_like, iters)
        bus = 2 * (world - 1) / world * nbytes
        row = {"MB": nbytes / 2 ** 20, "world": world,
               "fused_us": fused * 1e3, "nccl_us": nccl * 1e3,
               "nccl_plus_norms_us": ref * 1e3,
               "fused_busbw_GBps": bus / fused / 1e6,
               "nccl_busbw_GBps": bus / nccl / 1e6,
               "fused_frac_of_770": bus / fused / 1e6 / 770.0,
               "provider": provider, "ctas": ctas, "nvls": used_nvls,
               "variants_us": variants}
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
