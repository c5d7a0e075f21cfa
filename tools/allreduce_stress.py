#!/usr/bin/env python
"""Stress test of the flag-synchronised fused all-reduce:
there is no sanitizer for cross-GPU release/acquire protocols, so this hammers
them instead.

    torchrun --nproc-per-node N tools/allreduce_stress.py --iters 20000 [--nvls]

Every iteration each rank
  * sleeps a random time ON THE DEVICE (skews the ranks against each other so
    every wait / flag ordering occurs),
  * back-propagates through a randomly chosen parameter so that its gradient
    encodes (rank, iteration, index) -- the ordinary hook path: bucket ready
    -> fused all-reduce on the comm stream -> finalize,
  * compares the reduced gradient and the |mean gradient|^2 statistic with
    their closed forms.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20000)
    ap.add_argument("--max-kb", type=int, default=4096)
    ap.add_argument("--nvls", action="store_true",
                    help="force the multimem flavour for every bucket")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cap-mb", type=float, default=0.25,
                    help="bucket cap: small = many kernels per step")
    args = ap.parse_args()
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    if args.nvls:
        # every bucket the one-shot flavour does not take goes through the
        # switch
        os.environ["ADAPTDL_B200_NVLS_MIN_MB"] = "0.05"
        os.environ["ADAPTDL_B200_NVLS_MIN_WORLD"] = "2"
    from adaptdl_b200.parallel.reducer_cuda import CudaGradReducer
    # ONE reducer over parameters of very different sizes with a small bucket
    # cap: every step launches a chain of bucket kernels of all flavours
    # (one-shot for the small ones, two-shot / NVLS for the large ones, the
    # last one carrying the fused finalize) without any barrier between them
    sizes = [int(2 ** e) for e in (8, 10, 12, 14, 16, 18, 20)
             if 2 ** e * 4 <= args.max_kb * 1024] + [3001, 77]
    params = [torch.nn.Parameter(torch.zeros(n, device=dev)) for n in sizes]
    red = CudaGradReducer([{"params": [p]} for p in params], world, rank,
                          lambda: True, bucket_cap_mb=args.cap_mb)
    if rank == 0:
        from collections import Counter
        print("buckets:", len(red.arenas[0].buckets), "flavours:",
              dict(Counter(red._flavour.values())), flush=True)
    bad = torch.zeros(1, dtype=torch.int64, device=dev)
    idx_cache = {n: torch.arange(n, device=dev, dtype=torch.float32)
                 for n in sizes}
    ranks = torch.arange(1, world + 1, dtype=torch.float64)
    for it in range(args.iters):
        a = float(it % 97 + 1) * 1e-2
        # gradient(rank, i) = a * (rank + 1) + (i % 13) * 1e-3
        red.zero()
        loss = 0
        for k, (p, n) in enumerate(zip(params, sizes)):
            weight = a * (rank + 1) + (idx_cache[n] % 13) * 1e-3
            loss = loss + (p * weight).sum()
        # device-side skew: up to a few hundred microseconds, different on
        # every rank and iteration
        torch.cuda._sleep(int(torch.randint(
            0, 400000, (1,), generator=torch.Generator().manual_seed(
                args.seed + it * world + rank))))
        loss.backward()
        stats = red.pop_stats()                  # waits for the finalize
        for g, (p, n) in enumerate(zip(params, sizes)):
            base = (idx_cache[n] % 13) * 1e-3
            want = a * (world + 1) / 2.0 + base
            bad += (~torch.isclose(p.grad, want, rtol=1e-5,
                                   atol=1e-6)).sum()
            total = float((want.double() ** 2).sum())
            bad += int(abs(float(stats.total_sqr[g]) - total)
                       > 1e-4 * total)
            # sum over replicas of |g_r|^2 has a closed form too
            b64 = base.double().cpu()
            local = float(((a * ranks[:, None] + b64[None, :]) ** 2).sum())
            bad += int(abs(float(stats.local_sqr[g]) - local)
                       > 1e-4 * local)
        if rank == 0 and (it + 1) % 2000 == 0:
            print("iter {}: mismatches so far {}".format(it + 1, int(bad)),
                  flush=True)
    total_bad = bad.clone()
    dist.all_reduce(total_bad)
    if rank == 0:
        print("STRESS_{} iters={} world={} nvls={} mismatches={}".format(
            "OK" if int(total_bad) == 0 else "FAILED", args.iters, world,
            args.nvls, int(total_bad)), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(total_bad) == 0 else 1)


if __name__ == "__main__":
    main()
