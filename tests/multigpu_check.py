"""Multi-GPU correctness of the fused all-reduce / statistics / broadcast.
Launched by torchrun (see test_gpu_kernels.py::test_multi_gpu_fused_allreduce
or directly:  torchrun --nproc-per-node N tests/multigpu_check.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from adaptdl_b200.parallel.reducer_cuda import CudaGradReducer
    from adaptdl_b200.parallel.reducer_torch import TorchGradReducer

    shapes = [(64, 3, 3, 3), (64,), (256, 128, 3, 3), (1000, 512), (7,),
              (1,), (512,), (2049, 33)]
    group_of = [0, 0, 1, 2, 2, 3, 3, 4]
    for dtype, rtol in ((torch.float32, 1e-5), (torch.bfloat16, 3e-2)):
        def build(cls):
            g = torch.Generator().manual_seed(0)
            params = [torch.nn.Parameter(
                torch.randn(s, generator=g).to(dev, dtype)) for s in shapes]
            groups = [{"params": []} for _ in range(5)]
            for p, gi in zip(params, group_of):
                groups[gi]["params"].append(p)
            flag = [True]
            red = cls(groups, world, rank, lambda: flag[0],
                      bucket_cap_mb=0.25)
            return params, red, flag

        pa, ra, fa = build(CudaGradReducer)
        pb, rb, fb = build(TorchGradReducer)
        if rank == 0 and dtype == torch.float32:
            print("provider:", ra._provider.name, "buckets:",
                  len(ra.arenas[0].buckets), flush=True)

        def backward(params, seed, scale=1.0):
            g = torch.Generator().manual_seed(seed * 1000 + rank)
            loss = 0
            for p in params:
                w = torch.randn(p.shape, generator=g).to(dev, dtype)
                loss = loss + (p * w).sum() * scale
            loss.backward()

        for it in range(12):
            accumulate = it % 3 == 2
            ra.zero(), rb.zero()
            if accumulate:
                fa[0] = fb[0] = False
                backward(pa, 100 + it, 0.5), backward(pb, 100 + it, 0.5)
            fa[0] = fb[0] = True
            backward(pa, it), backward(pb, it)
            sa, sb = ra.pop_stats(), rb.pop_stats()
            for x, y in zip(pa, pb):
                assert torch.allclose(x.grad.float(), y.grad.float(),
                                      rtol=rtol, atol=rtol), \
                    (it, x.shape, (x.grad.float() - y.grad.float()).abs().max())
            assert sa.count == sb.count == world * (2 if accumulate else 1)
            np.testing.assert_allclose(sa.local_sqr, sb.local_sqr, rtol=rtol)
            np.testing.assert_allclose(sa.total_sqr, sb.total_sqr, rtol=rtol)
            # identical bits on every rank (fixed summation order)
            gathered = [None] * world
            dist.all_gather_object(gathered, (sa.local_sqr.tobytes(),
                                              sa.total_sqr.tobytes()))
            assert all(g == gathered[0] for g in gathered)
        if os.environ.get("ADAPTDL_EXPECT_NVLS") == "1":
            assert getattr(ra, "nvls_launches", 0) > 0, \
                "NVLS flavour was requested but never launched"
        # every rank ends with the same averaged gradients
        flat = torch.cat([p.grad.reshape(-1).float() for p in pa])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(flat, ref)

        # broadcast: rank 0's tensors win, including one > staging chunk
        tensors = [torch.full((1000,), float(rank + 1), device=dev),
                   torch.full((3, 5), float(rank + 7), device=dev,
                              dtype=torch.bfloat16),
                   torch.arange(5 * 1024 * 1024 + 3, device=dev,
                                dtype=torch.float32) * (rank + 1)]
        ra.broadcast_parameters(tensors)
        torch.cuda.synchronize()
        assert float(tensors[0][0]) == 1.0 and float(tensors[1][0, 0]) == 7.0
        assert float(tensors[2][-1]) == float(5 * 1024 * 1024 + 2)

    # end-to-end through the public API
    os.environ["ADAPTDL_NUM_REPLICAS"] = str(world)
    os.environ["ADAPTDL_NUM_NODES"] = "1"
    os.environ["ADAPTDL_REPLICA_RANK"] = str(rank)
    import adaptdl_b200.torch as adl
    from adaptdl_b200 import collective
    collective.initialize("127.0.0.1", int(os.environ["MASTER_PORT"]) + 7,
                          rank, world)
    torch.manual_seed(rank)          # different init: broadcast must fix it
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(),
                                torch.nn.BatchNorm1d(32),
                                torch.nn.Linear(32, 4)).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    net = adl.AdaptiveDataParallel(model, opt)
    assert isinstance(net.reducer, CudaGradReducer)
    data = torch.utils.data.TensorDataset(torch.randn(512, 16),
                                          torch.randint(0, 4, (512,)))
    loader = adl.AdaptiveDataLoader(data, batch_size=8 * world,
                                    shuffle=True, drop_last=True)
    for _ in adl.remaining_epochs_until(2):
        for x, y in loader:
            opt.zero_grad()
            torch.nn.functional.cross_entropy(net(x.to(dev)),
                                              y.to(dev)).backward()
            opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(flat, ref), "replicas diverged"
    assert net.gns.get_progress() > 0 and np.isfinite(net.gain)
    dist.barrier()
    if rank == 0:
        print("MULTIGPU_OK world={} gain={:.4f}".format(world, net.gain),
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
