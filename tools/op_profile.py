"""Which ATen ops (with input shapes) own the GPU time of one training step?

    python tools/op_profile.py --model bert [--top 40]

``tools/step_profile.py`` groups by KERNEL name, which cannot tell two
``direct_copy_kernel`` call sites apart; this one groups the same CUPTI
records by the ATen operator that launched them and its input shapes
(``torch.profiler`` with ``record_shapes``), for the bench configuration of
the model (bf16 parameters, bf16 autocast, fused ops on).
"""

import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="bert")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    from adaptdl_b200 import models
    from adaptdl_b200.torch import mixed_precision_params
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    if args.model == "bert":
        net = models.bert_base_mlm(max_len=128).to(dev)
        x = torch.randint(0, 28996, (args.batch, 128), device=dev)
        t = torch.randint(0, 28996, (args.batch, 128), device=dev)

        def loss_fn(out, tgt):
            return torch.nn.functional.cross_entropy(
                out.view(-1, out.shape[-1]), tgt.view(-1))
    else:
        net = models.get_model("ResNet18").to(dev).to(
            memory_format=torch.channels_last)
        x = torch.randn(128, 3, 32, 32, device=dev).contiguous(
            memory_format=torch.channels_last)
        t = torch.randint(0, 10, (128,), device=dev)
        loss_fn = torch.nn.functional.cross_entropy
    mixed_precision_params(net)
    opt = torch.optim.SGD(net.parameters(), lr=0.01)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = loss_fn(net(x), t)
        loss.backward()
        opt.step()
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA],
                 record_shapes=True) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    rows = collections.defaultdict(lambda: [0, 0.0])
    for evt in prof.key_averages(group_by_input_shape=True):
        # self device time: kernels launched by this op itself
        us = getattr(evt, "self_device_time_total", 0.0)
        if us <= 0:
            continue
        key = (evt.key, str(evt.input_shapes)[:110])
        rows[key][0] += evt.count
        rows[key][1] += us
    total = sum(v[1] for v in rows.values())
    print("total self device time per step: {:.1f} us".format(
        total / args.steps))
    for (name, shapes), (count, us) in sorted(
            rows.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print("{:9.1f} us {:5.1f}% x{:<4.0f} {:<42s} {}".format(
            us / args.steps, 100 * us / total, count / args.steps,
            name[:42], shapes))


if __name__ == "__main__":
    main()
