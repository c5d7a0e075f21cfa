"""Per-kernel GPU time of one training step under REAL cache conditions
(torch.profiler / CUPTI, no cache flush between kernels, unlike ncu).

    python tools/step_profile.py --model resnet18 --out gpurun_out/step_profile.json

Eager forward + backward + SGD step of the model with bf16 autocast and
channels-last, batch 128; kernels are grouped by name and sorted by total
device time. Use ADAPTDL_B200_FUSED_BN=0 / ADAPTDL_B200_FUSED_GEMM=0 to get
the PyTorch-composition baseline for the same model.
"""

import argparse
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet18")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default=None)
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--bf16-params", action="store_true",
                    help="store conv/linear weights in bf16 (no autocast "
                         "cast kernels); plain torch SGD, no masters")
    args = ap.parse_args()
    from adaptdl_b200 import models
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    if args.model == "bert":
        net = models.bert_base_mlm().to(dev)
        x = torch.randint(0, 28996, (args.batch // 4, 128), device=dev)
        t = torch.randint(0, 28996, (args.batch // 4, 128), device=dev)

        def loss_fn(out, tgt):
            return torch.nn.functional.cross_entropy(
                out.view(-1, out.shape[-1]), tgt.view(-1))
    else:
        net = models.get_model(
            {"resnet18": "ResNet18"}.get(args.model, args.model)).to(dev)
        net = net.to(memory_format=torch.channels_last)
        x = torch.randn(args.batch, 3, 32, 32, device=dev).contiguous(
            memory_format=torch.channels_last)
        t = torch.randint(0, 10, (args.batch,), device=dev)
        loss_fn = torch.nn.functional.cross_entropy
    if args.bf16_params:
        from adaptdl_b200.torch import mixed_precision_params
        mixed_precision_params(net)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = loss_fn(net(x), t)
        loss.backward()
        opt.step()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for evt in prof.events():
        if evt.device_type == torch.autograd.DeviceType.CUDA:
            a = agg[evt.name]
            a[0] += 1
            a[1] += evt.device_time
    rows = sorted(({"kernel": k, "launches_per_step": v[0] / args.steps,
                    "us_per_step": v[1] / args.steps}
                   for k, v in agg.items()),
                  key=lambda r: -r["us_per_step"])
    total = sum(r["us_per_step"] for r in rows)
    print("total GPU kernel time per step: {:.1f} us over {:.0f} launches"
          .format(total, sum(r["launches_per_step"] for r in rows)))
    for r in rows[:args.top]:
        print("{:9.1f} us {:5.1f}% x{:<5.0f} {}".format(
            r["us_per_step"], 100 * r["us_per_step"] / total,
            r["launches_per_step"], r["kernel"][:110]))
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"model": args.model, "batch": args.batch,
                       "total_us_per_step": total, "kernels": rows,
                       "env": {k: v for k, v in os.environ.items()
                               if k.startswith("ADAPTDL_B200")}}, f, indent=1)


if __name__ == "__main__":
    main()
