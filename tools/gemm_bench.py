"""Fused tcgen05 Linear+bias+GELU vs the PyTorch composition (cuBLASLt GEMM +
GELU kernel): numerics against an fp32 reference and CUDA-event timings.

    python tools/gemm_bench.py [--out gpurun_out/gemm_bench.json]
"""

import argparse
import json
import sys
import os

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptdl_b200.ops import check_errors, gemm_bias_act  # noqa: E402


ITERS = [50]


def time_us(fn, iters=None, warmup=5, flush=None):
    iters = iters or ITERS[0]
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)                   # > L2: evict operands/outputs
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
    ap.add_argument("--shapes", default="4096x3072x768,16384x3072x768,"
                    "8192x4096x1024,4096x768x3072,512x3072x768,1000x3072x768")
    ap.add_argument("--block-n", default="128,256")
    ap.add_argument("--cluster-m", default="1,2,4")
    ap.add_argument("--iters", type=int, default=50)
    args = ap.parse_args()
    block_ns = [int(v) for v in args.block_n.split(",")]
    cluster_ms = [int(v) for v in args.cluster_m.split(",")]
    ITERS[0] = args.iters
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    flush = torch.zeros(80 * 1024 * 1024, dtype=torch.float32, device=dev)
    peaks = {}
    try:
        with open(os.path.join(os.path.dirname(__file__), "..",
                               "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    rows = []
    for shape in args.shapes.split(","):
        m, n, k = (int(v) for v in shape.split("x"))
        x = torch.randn(m, k, device=dev).bfloat16()
        w = (torch.randn(n, k, device=dev) / k ** 0.5).bfloat16()
        b = torch.randn(n, device=dev)
        ref_z = x.float() @ w.float().t() + b
        ref_y = F.gelu(ref_z)
        row = {"M": m, "N": n, "K": k}
        for bn in block_ns:
            for cm in cluster_ms:
                if n % bn or (cm == 22 and bn != 256):
                    continue
                tag = "bn%d_cm%d" % (bn, cm)
                y, z = gemm_bias_act(x, w, b, "gelu", True, block_n=bn,
                                     cluster_m=cm)
                torch.cuda.synchronize()
                check_errors()
                row["max_err_y_" + tag] = \
                    (y.float() - ref_y).abs().max().item()
                row["max_err_z_" + tag] = \
                    (z.float() - ref_z).abs().max().item()
                row["fused_%s_us" % tag] = time_us(
                    lambda: gemm_bias_act(x, w, b, "gelu", True, block_n=bn,
                                          cluster_m=cm), flush=flush)
                row["fused_nosave_%s_us" % tag] = time_us(
                    lambda: gemm_bias_act(x, w, b, "gelu", False, block_n=bn,
                                          cluster_m=cm), flush=flush)
        bb = b.bfloat16()
        yt = F.gelu(F.linear(x, w, bb))
        row["max_err_torch"] = (yt.float() - ref_y).abs().max().item()
        row["torch_linear_gelu_us"] = time_us(
            lambda: F.gelu(F.linear(x, w, bb)), flush=flush)
        row["torch_linear_only_us"] = time_us(
            lambda: F.linear(x, w, bb), flush=flush)
        row["flush_only_us"] = time_us(lambda: None, flush=flush)
        best = min(v for k_, v in row.items()
                   if k_.startswith("fused_bn"))
        row["best"] = min((v, k_) for k_, v in row.items()
                          if k_.startswith("fused_bn"))[1]
        flops = 2.0 * m * n * k
        row["fused_TFLOPs"] = flops / best / 1e6
        row["torch_TFLOPs"] = flops / row["torch_linear_gelu_us"] / 1e6
        row["speedup_vs_torch"] = row["torch_linear_gelu_us"] / best
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"rows": rows, "peaks": peaks}, f, indent=1)


if __name__ == "__main__":
    main()
