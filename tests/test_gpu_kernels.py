"""Numerics of the sm_100a gradient kernels against plain PyTorch fp32
references, and reducer-level equivalence (fused vs stock-torch reducer)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _params(shapes, dtype, device, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(device, dtype))
            for s in shapes]


SHAPES = [(64, 3, 3, 3), (64,), (64,), (128, 64, 3, 3), (128,), (1000, 37),
          (5,), (1,), (513, 129)]
GROUPS = [0, 1, 1, 2, 2, 3, 3, 3, 4]


def _make(cls, dtype, device, sync_flag, cap_mb=0.05):
    params = _params(SHAPES, dtype, device)
    groups = [{"params": []} for _ in range(max(GROUPS) + 1)]
    for p, g in zip(params, GROUPS):
        groups[g]["params"].append(p)
    red = cls(groups, 1, 0, lambda: sync_flag[0], bucket_cap_mb=cap_mb)
    return params, red


def _backward(params, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    loss = 0
    for p in params:
        w = torch.randn(p.shape, generator=g).to(p.device, p.dtype)
        loss = loss + (p * w).sum() * scale
    loss.backward()


def test_native_library_loads_and_driver_is_up():
    from adaptdl_b200 import _native
    lib = _native.load()
    assert lib.adl_symm_init() == 0, lib.adl_symm_last_error()
    assert lib.adl_sm_count(0) >= 100
    assert lib.adl_topo_vmm_fd_supported(0) == 1


@pytest.mark.parametrize("dtype,rtol", [(torch.float32, 1e-5),
                                        (torch.bfloat16, 2e-2),
                                        (torch.float16, 2e-3)])
def test_primitives_match_torch_reference(dtype, rtol):
    """pair -> accumulate x2 -> final reduce, statistics and buffers."""
    from adaptdl_b200.parallel.reducer_cuda import CudaGradReducer
    from adaptdl_b200.parallel.reducer_torch import TorchGradReducer
    dev = torch.device("cuda", 0)
    flag_a, flag_b = [True], [True]
    pa, ra = _make(CudaGradReducer, dtype, dev, flag_a)
    pb, rb = _make(TorchGradReducer, dtype, dev, flag_b)
    assert len(ra.arenas[0].buckets) > 2

    def both(seed, sync, scale=1.0):
        flag_a[0] = flag_b[0] = sync
        _backward(pa, seed, scale)
        _backward(pb, seed, scale)

    def compare_grads():
        for x, y in zip(pa, pb):
            assert torch.allclose(x.grad.float(), y.grad.float(),
                                  rtol=rtol, atol=rtol), (x.shape,)

    def compare_stats():
        sa, sb = ra.pop_stats(), rb.pop_stats()
        assert sa.count == sb.count
        np.testing.assert_allclose(sa.local_sqr, sb.local_sqr, rtol=rtol)
        np.testing.assert_allclose(sa.total_sqr, sb.total_sqr, rtol=rtol)
        assert (sa.pair is None) == (sb.pair is None)
        if sa.pair is not None:
            np.testing.assert_allclose(sa.pair[0], sb.pair[0], rtol=rtol)
            np.testing.assert_allclose(sa.pair[1], sb.pair[1], rtol=rtol)
        assert sa.sync_time is not None and sa.sync_time >= 0
        return sa

    # step 1: single sample (stash only); step 2: differenced pair
    for seed in (1, 2):
        ra.zero(), rb.zero()
        both(seed, True)
        compare_grads()
        s = compare_stats()
    assert s.pair is not None
    # step 3: 3 micro-batches accumulated, then the synchronising one
    ra.zero(), rb.zero()
    both(3, False, 0.5)
    both(4, False, 2.0)
    both(5, True)
    assert ra.accum_count == rb.accum_count == 3
    compare_grads()
    s = compare_stats()
    assert s.count == 3 and s.pair is None
    # padding stays zero (statistics of later steps rely on it)
    arena = ra.arenas[0]
    mask = torch.ones(arena.grad.numel(), dtype=torch.bool, device=dev)
    for b in arena.buckets:
        for seg in b.segments:
            mask[seg.start:seg.start + seg.length] = False
    assert float(arena.grad[mask].abs().sum()) == 0.0
    assert ra.launches > 0


def test_preconditioned_statistics():
    from adaptdl_b200.parallel.reducer_cuda import CudaGradReducer
    from adaptdl_b200.parallel.reducer_torch import TorchGradReducer
    dev = torch.device("cuda", 0)
    fa, fb = [True], [True]
    pa, ra = _make(CudaGradReducer, torch.float32, dev, fa)
    pb, rb = _make(TorchGradReducer, torch.float32, dev, fb)
    pre_a = {id(p): torch.rand_like(p) + 0.5 for p in pa}
    pre_b = {id(q): pre_a[id(p)].clone() for p, q in zip(pa, pb)}
    ra.set_preconditioner(lambda p: pre_a[id(p)])
    rb.set_preconditioner(lambda p: pre_b[id(p)])
    for seed in (1, 2):
        ra.zero(), rb.zero()
        _backward(pa, seed), _backward(pb, seed)
        sa, sb = ra.pop_stats(), rb.pop_stats()
        np.testing.assert_allclose(sa.total_sqr, sb.total_sqr, rtol=1e-5)
    np.testing.assert_allclose(sa.pair[0], sb.pair[0], rtol=1e-5)
    np.testing.assert_allclose(sa.pair[1], sb.pair[1], rtol=1e-5)


def test_non_finite_gradients_surface_in_statistics():
    from adaptdl_b200.parallel.reducer_cuda import CudaGradReducer
    dev = torch.device("cuda", 0)
    flag = [True]
    params, red = _make(CudaGradReducer, torch.float32, dev, flag)
    red.zero()
    (params[0] * float("nan")).sum().backward()
    stats = red.pop_stats()
    assert not np.isfinite(stats.total_sqr[0])
    assert np.all(np.isfinite(stats.total_sqr[1:]))


def test_gns_trajectory_matches_torch_reducer():
    """Same model, same data: the fused path must reproduce the oracle's
    gain / statistics trajectory (SURVEY 7.4 acceptance test)."""
    from unittest.mock import Mock
    from adaptdl_b200.parallel import make_reducer
    from adaptdl_b200.torch.gradient_noise_scale import GradientNoiseScale
    from adaptdl_b200.torch.scaling_rules import AdaScale
    dev = torch.device("cuda", 0)
    traj = {}
    for backend in ("cuda", "torch"):
        torch.manual_seed(0)
        model = torch.nn.Sequential(
            torch.nn.Linear(32, 64), torch.nn.ReLU(),
            torch.nn.Linear(64, 10)).to(dev)
        opt = torch.optim.SGD([{"params": [p]} for p in model.parameters()],
                              lr=0.05, momentum=0.9)
        adp = Mock(require_backward_grad_sync=True)
        red = make_reducer(opt.param_groups, 1, 0,
                           lambda: adp.require_backward_grad_sync,
                           backend=backend)
        gns = GradientNoiseScale(adp, opt, num_replicas=1, accum_scale=1.0,
                                 reducer=red)
        adp.gns = gns
        rule = AdaScale()
        rule.initialize(adp, opt, patch_optimizer=True)
        gen = torch.Generator().manual_seed(7)
        out = []
        for step in range(40):
            accumulate = step >= 20
            for micro in range(2 if accumulate else 1):
                adp.require_backward_grad_sync = \
                    (micro == 1) or not accumulate
                x = torch.randn(16, 32, generator=gen).to(dev)
                y = torch.randint(0, 10, (16,), generator=gen).to(dev)
                if micro == 0:
                    opt.zero_grad()
                torch.nn.functional.cross_entropy(model(x), y).backward()
            opt.step()
            out.append((gns.sqr_avg(), gns.var_avg(), gns.gain(4.0),
                        gns.get_progress()))
        traj[backend] = np.array(out)
    np.testing.assert_allclose(traj["cuda"], traj["torch"], rtol=2e-3)


def test_adaptive_data_parallel_resnet_step_uses_fused_reducer():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.smoke()


@pytest.mark.skipif(torch.cuda.device_count() < 2,
                    reason="needs >= 2 GPUs")
@pytest.mark.parametrize("provider", ["native", "torch"])
def test_multi_gpu_fused_allreduce(provider):
    n = min(torch.cuda.device_count(), 8)
    env = dict(os.environ, ADAPTDL_B200_SYMM=provider)
    for key in list(env):
        if key.startswith("ADAPTDL_") and key != "ADAPTDL_B200_SYMM":
            env.pop(key)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29611",
           os.path.join(ROOT, "tests", "multigpu_check.py")]
    proc = subprocess.run(cmd, env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-4000:]
    assert "MULTIGPU_OK" in proc.stdout
