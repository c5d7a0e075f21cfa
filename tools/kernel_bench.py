#!/usr/bin/env python
"""Single-GPU roofline of the local gradient kernels (CUDA events, warm-up,
L2-flush-by-size: every buffer is far larger than the 126 MB L2).

    python tools/kernel_bench.py [--mb 512] [--out profiles/kernel_bench.json]

Reports achieved HBM bandwidth of each kernel against the measured copy
bandwidth in MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, iters=20, warmup=5, stream=None):
    """Device time per call; events are recorded on the stream the kernels
    actually run on (the reducer's communication stream)."""
    stream = stream or torch.cuda.current_stream()
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(True), torch.cuda.Event(True)
    start.record(stream)
    for _ in range(iters):
        fn()
    end.record(stream)
    torch.cuda.synchronize()
    return start.elapsed_time(end) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--groups", type=int, default=62)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from adaptdl_b200.parallel.reducer_cuda import CudaGradReducer
    from adaptdl_b200.parallel.engine import DeviceEngine
    from adaptdl_b200.torch.scaling_rules import AdaScale
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    n_params = args.groups
    numel = args.mb * (1 << 20) // 4 // n_params
    params = [torch.nn.Parameter(torch.randn(numel, device=dev))
              for _ in range(n_params)]
    opt = torch.optim.SGD([{"params": [p]} for p in params], lr=0.1,
                          momentum=0.9, weight_decay=5e-4)
    flag = [True]
    red = CudaGradReducer(opt.param_groups, 1, 0, lambda: flag[0],
                          bucket_cap_mb=args.mb * 2)
    arena = red.arenas[0]
    arena.grad.normal_()
    nbytes = max(b.length for b in arena.buckets) * 4
    opt.state["gns"] = {"sqr_avg": 1.0, "var_avg": 0.0, "progress": 0.0,
                        "biased": False}
    engine = DeviceEngine(red, opt, AdaScale(), opt.state["gns"])
    engine.adopt_optimizer_state()
    engine.sync_ctrl(1.0, 0.999)
    bucket = max(arena.buckets, key=lambda b: b.length)
    red._k_before = 0
    results = {}

    def report(name, ms, passes):
        gbs = passes * nbytes / ms / 1e6
        results[name] = {"ms": ms, "bytes_moved": passes * nbytes,
                         "GBps": gbs, "frac_of_measured_hbm": gbs / hbm}
        print("{:<26s} {:8.3f} ms  {:8.1f} GB/s  {:5.1f}% of measured "
              "copy bandwidth ({} passes over {:.0f} MB)".format(
                  name, ms, gbs, 100 * gbs / hbm, passes, nbytes / 1e6))

    red._ensure(arena, "acc")
    red._ensure(arena, "prev")
    # (a) the shape these kernels have inside a training step: thin grids
    # (64 CTAs x 256 threads) that share the GPU with the backward kernels
    report("pair_norm_stash thin", timed(lambda: red._pair(arena, bucket), stream=red._comm), 3)
    report("fold_acc thin", timed(lambda: red._fold_acc(arena, bucket), stream=red._comm), 4)
    # (b) the same kernels with the GPU to themselves (2 CTAs/SM x 512)
    red._local_ctas, red._reduce_threads = 2 * red._sm_count, 512
    report("pair_norm_stash (F2)", timed(lambda: red._pair(arena, bucket), stream=red._comm), 3)
    report("fold_acc", timed(lambda: red._fold_acc(arena, bucket), stream=red._comm), 4)
    report("fold_final", timed(lambda: red._fold_final(arena, bucket), stream=red._comm), 4)
    report("allreduce_gns world=1",
           timed(lambda: red._reduce(arena, bucket, 0.5, False),
                 stream=red._comm), 2)
    arena_bytes = arena.grad.numel() * 4
    ms = timed(engine.optimizer_step)
    nbytes_saved, nbytes = nbytes, arena_bytes
    report("fused_sgd (F4, whole arena)", ms, 5)
    src = torch.empty_like(arena.grad)
    report("torch copy_ (reference)", timed(lambda: src.copy_(arena.grad)), 2)
    nbytes = nbytes_saved
    red._accum_count = 1

    def fin():
        red._finalize_step()
    ms = timed(fin, iters=50, stream=red._comm)
    results["finalize+estimator (62 groups)"] = {"ms": ms}
    print("{:<26s} {:8.3f} ms (latency-bound, one CTA)".format(
        "finalize+estimator", ms))
    results["_meta"] = {"arena_MB": nbytes / 1e6, "groups": n_params,
                        "hbm_gbs_measured": hbm,
                        "gpu": torch.cuda.get_device_name(0)}
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)),
                    exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
