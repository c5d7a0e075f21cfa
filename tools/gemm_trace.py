"""Pipeline trace of the tcgen05 GEMM (CTA 0): per K-slice SM-clock stamps of
the TMA issue and of the MMA warp's wait on the full barrier.

    python tools/gemm_trace.py --shape 4096x3072x768 --cluster-m 22
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptdl_b200.ops import gemm_bias_act  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="4096x3072x768")
    ap.add_argument("--block-n", type=int, default=256)
    ap.add_argument("--cluster-m", type=int, default=22)
    ap.add_argument("--max-ctas", type=int, default=0)
    ap.add_argument("--nosave", action="store_true")
    args = ap.parse_args()
    m, n, k = (int(v) for v in args.shape.split("x"))
    dev = torch.device("cuda:0")
    x = torch.randn(m, k, device=dev).bfloat16()
    w = (torch.randn(n, k, device=dev) / k ** 0.5).bfloat16()
    b = torch.randn(n, device=dev)
    for _ in range(3):
        gemm_bias_act(x, w, b, "gelu", not args.nosave, args.block_n,
                      args.cluster_m, args.max_ctas)
    trace = torch.zeros(3, 256, dtype=torch.int64, device=dev)
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    gemm_bias_act(x, w, b, "gelu", not args.nosave, args.block_n,
                  args.cluster_m, args.max_ctas, trace=trace)
    e.record()
    torch.cuda.synchronize()
    t = trace.cpu()
    n_it = int((t[0] > 0).sum())
    t0 = int(t[0, 0])
    print("kernel {:.1f} us; {} K-slices traced in CTA 0 (K/64 = {} per tile)"
          .format(s.elapsed_time(e) * 1e3, n_it, k // 64))
    print(" it  tma_issue  wait_begin  wait_end  waited  load_latency")
    for i in range(n_it):
        issue, wb, we = (int(t[j, i]) - t0 for j in range(3))
        print("{:3d} {:9d} {:10d} {:9d} {:7d} {:10d}".format(
            i, issue, wb, we, we - wb, we - issue))
    if n_it > 4:
        ends = [int(t[2, i]) for i in range(n_it)]
        print("mean K-slice interval: {:.0f} cycles".format(
            (ends[-1] - ends[0]) / (n_it - 1)))


if __name__ == "__main__":
    main()
