"""Which part of the GEMM epilogue costs what: time the kernel with parts of
the epilogue switched off (diagnostic bits of ``act``; outputs are wrong)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptdl_b200.ops import gemm_bias_act  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.zeros(80 * 1024 * 1024, dtype=torch.float32, device=dev)


def t_us(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.add_(1)
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters * 1e3


for shape in ((4096, 3072, 768), (16384, 3072, 768)):
    m, n, k = shape
    x = torch.randn(m, k, device=dev).bfloat16()
    w = (torch.randn(n, k, device=dev) / k ** 0.5).bfloat16()
    b = torch.randn(n, device=dev)
    for cm in (22, 1):
        row = {}
        for tag, act, save in (("full", 1, True), ("nosave", 1, False),
                               ("no_tma_store", 1 | 8, True),
                               ("no_math", 1 | 16, True),
                               ("no_fence", 1 | 32, True),
                               ("no_store_no_math", 1 | 8 | 16, True),
                               ("no_store_math_fence", 1 | 8 | 16 | 32, True),
                               ("nothing", 1 | 8 | 16 | 32 | 64, True)):
            row[tag] = round(t_us(lambda: gemm_bias_act(
                x, w, b, act, save, 256, cm)), 1)
        print(shape, "cluster_m", cm, row, flush=True)
