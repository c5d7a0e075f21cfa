"""Fused dropout + residual + LayerNorm op (csrc/adl_ln.cu) against the
PyTorch composition, BERT-base shape ([4096, 768] bf16), CUDA-event timed
with an L2 flush between iterations (cold) and without (warm).

    python tools/ln_bench.py [--rows 4096 --width 768]
"""

import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, iters, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--width", type=int, default=768)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from adaptdl_b200.ops import dropout_add_layer_norm
    dev = torch.device("cuda:0")
    m, d = args.rows, args.width
    x = torch.randn(m, d, device=dev).bfloat16().requires_grad_(True)
    h = torch.randn(m, d, device=dev).bfloat16().requires_grad_(True)
    w = torch.ones(d, device=dev, requires_grad=True)
    b = torch.zeros(d, device=dev, requires_grad=True)
    g = torch.randn(m, d, device=dev).bfloat16()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def fused_fwd():
        return dropout_add_layer_norm(x, h, w, b, 0.1, True)

    def torch_fwd():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return F.layer_norm(x + F.dropout(h, 0.1, True), (d,), w, b)

    out = {}
    for name, fwd in (("fused", fused_fwd), ("torch", torch_fwd)):
        y = fwd()

        def bwd():
            torch.autograd.grad(y, [x, h, w, b], g, retain_graph=True)
        for tag, fl in (("cold", flush), ("warm", None)):
            out["{}_fwd_{}_us".format(name, tag)] = timed(fwd, args.iters, fl)
            out["{}_bwd_{}_us".format(name, tag)] = timed(bwd, args.iters, fl)
    bytes_fwd = m * d * (2 * 4 + 1)
    bytes_bwd = m * d * (2 * 4 + 1)
    out["ideal_fwd_us_at_6.5TBps"] = bytes_fwd / 6.5e6
    out["ideal_bwd_us_at_6.5TBps"] = bytes_bwd / 6.5e6
    print(json.dumps(out, indent=1))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
