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
    args = ap.parse_args()
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    if args.nvls:
        os.environ["ADAPTDL_B200_NVLS_MIN_MB"] = "0"
    from adaptdl_b200.parallel.reducer_cuda import CudaGradReducer
    gen = torch.Generator().manual_seed(args.seed)        # same on all ranks
    sizes = [int(2 ** e) for e in range(8, 21) if 2 ** e * 4 <= args.max_kb
             * 1024]
    reducers = []
    for numel in sizes:
        p = torch.nn.Parameter(torch.zeros(numel, device=dev))
        red = CudaGradReducer([{"params": [p]}], world, rank, lambda: True,
                              bucket_cap_mb=max(8.0 * numel / 2 ** 20, 1))
        reducers.append((p, red))
    bad = torch.zeros(1, dtype=torch.int64, device=dev)
    idx_cache = {n: torch.arange(n, device=dev, dtype=torch.float32)
                 for n in sizes}
    for it in range(args.iters):
        which = int(torch.randint(len(sizes), (1,), generator=gen))
        p, red = reducers[which]
        n = sizes[which]
        a = float(it % 97 + 1) * 1e-2
        # gradient(rank, i) = a * (rank + 1) + (i % 13) * 1e-3
        base = (idx_cache[n] % 13) * 1e-3
        weight = a * (rank + 1) + base
        # device-side skew: up to a few hundred microseconds, different on
        # every rank and iteration
        torch.cuda._sleep(int(torch.randint(
            0, 400000, (1,), generator=torch.Generator().manual_seed(
                args.seed + it * world + rank))))
        red.zero()
        (p * weight).sum().backward()
        stats = red.pop_stats()                  # waits for the finalize
        want = a * (world + 1) / 2.0 + base
        bad += (~torch.isclose(p.grad, want, rtol=1e-5, atol=1e-6)).sum()
        # |mean gradient|^2 has a closed form too (the per-replica statistic
        # is checked against the torch oracle in tests/multigpu_check.py)
        total = float((want.double() ** 2).sum())
        bad += int(abs(float(stats.total_sqr.sum()) - total) > 1e-4 * total)
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
