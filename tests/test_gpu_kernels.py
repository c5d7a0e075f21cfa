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
@pytest.mark.parametrize("provider", ["native", "torch", "native-nvls"])
def test_multi_gpu_fused_allreduce(provider):
    """Fused all-reduce + statistics against the torch oracle on every GPU of
    the box (fp32 and bf16); "native-nvls" forces every bucket through the
    multimem (in-switch reduction) flavour."""
    n = min(torch.cuda.device_count(), 8)
    env = dict(os.environ)
    for key in list(env):
        if key.startswith("ADAPTDL_"):
            env.pop(key)
    env["ADAPTDL_B200_SYMM"] = provider.split("-")[0]
    if provider == "native-nvls":
        env["ADAPTDL_B200_NVLS_MIN_MB"] = "0"
        env["ADAPTDL_B200_NVLS_MIN_WORLD"] = "2"
        env["ADAPTDL_B200_ONESHOT_KB"] = "0"     # else it takes every bucket
        env["ADAPTDL_EXPECT_NVLS"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29611",
           os.path.join(ROOT, "tests", "multigpu_check.py")]
    proc = subprocess.run(cmd, env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-4000:]
    assert "MULTIGPU_OK" in proc.stdout


# ------------------------------------------------------------------------
# device engine (on-GPU estimator + fused optimizer) and CUDA-graph step
# ------------------------------------------------------------------------

_COUNTER = [0]


def _single_process_runtime():
    import adaptdl_b200.torch as adl
    from adaptdl_b200 import collective
    os.environ.pop("ADAPTDL_CHECKPOINT_PATH", None)
    if not torch.distributed.is_initialized():
        os.environ.setdefault("ADAPTDL_MASTER_ADDR", "127.0.0.1")
        adl.init_process_group("nccl")
    elif not collective.is_initialized():
        collective.initialize("127.0.0.1", 0, 0, 1)
    return adl


def _train(adl, make_opt, rule_fn, fused, graphed, steps=14, accum=False,
           model_fn=None, lag=None, autocast=None, scaler=None):
    """Train a small model through the public API; returns (params, gns
    dict, net)."""
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if model_fn is None:
        model = torch.nn.Sequential(
            torch.nn.Linear(32, 64), torch.nn.ReLU(),
            torch.nn.Linear(64, 64), torch.nn.Tanh(),
            torch.nn.Linear(64, 10)).to(dev)
    else:
        model = model_fn().to(dev)
    opt = make_opt(model)
    _COUNTER[0] += 1
    net = adl.AdaptiveDataParallel(model, opt, scaling_rule=rule_fn(),
                                   mp_scaler=scaler,
                                   name="t{}".format(_COUNTER[0]),
                                   fused_step=fused)
    assert (net.engine is not None) == bool(fused)
    gen = torch.Generator().manual_seed(3)
    data = torch.utils.data.TensorDataset(
        torch.randn(16 * steps * 2, 32, generator=gen),
        torch.randint(0, 10, (16 * steps * 2,), generator=gen))
    loader = adl.AdaptiveDataLoader(data, batch_size=16, drop_last=True)
    if accum:
        loader.autoscale_batch_size(64, local_bsz_bounds=(8, 16),
                                    gradient_accumulation=True)
        helper = loader._elastic
        orig = helper._sync_local_bsz

        def fixed():
            orig()
            helper._state.current_local_bsz = 16
            helper._state.accumulation_steps = 1
            return 16
        helper._sync_local_bsz = fixed
    trainer = adl.GraphedTrainStep(
        net, opt, lambda n, x, y: torch.nn.functional.cross_entropy(n(x), y),
        warmup=2, enabled=graphed, autocast_dtype=autocast,
        grad_scaler=scaler)
    losses = []
    for epoch in adl.remaining_epochs_until(adl.finished_epochs() + 1):
        for i, (x, y) in enumerate(loader):
            losses.append(trainer(x, y).clone())
            if i + 1 >= steps * (2 if accum else 1):
                break
    torch.cuda.synchronize()
    if net.engine is not None:
        net.engine.pull_gns_state(opt.state["gns"])
    else:
        net.gns._flush()
    gns = {k: np.array(v, dtype=float) for k, v in opt.state["gns"].items()
           if k in ("sqr_avg", "var_avg", "progress")}
    params = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    return params, gns, net, trainer, torch.stack(losses)


def _sgd(model):
    return torch.optim.SGD([{"params": [p]} for p in model.parameters()],
                           lr=0.05, momentum=0.9, weight_decay=5e-4,
                           nesterov=True)


@pytest.mark.parametrize("accum", [False, True])
def test_device_engine_matches_host_path(accum):
    """On-GPU estimator + fused SGD reproduce the host estimator + torch SGD
    (same LR factors => same parameters, same running averages)."""
    adl = _single_process_runtime()
    from adaptdl_b200.torch.scaling_rules import AdaScale
    pa, ga, net_a, _, la = _train(adl, _sgd, AdaScale, True, False,
                                  accum=accum)
    pb, gb, net_b, _, lb = _train(adl, _sgd, AdaScale, False, False,
                                  accum=accum)
    assert net_a.engine is not None and net_b.engine is None
    torch.testing.assert_close(la, lb, rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(pa, pb, rtol=2e-4, atol=2e-5)
    for key in ("sqr_avg", "var_avg", "progress"):
        np.testing.assert_allclose(ga[key], gb[key], rtol=2e-3, atol=1e-9)
    assert ga["progress"] > 0


@pytest.mark.parametrize("graphed", [False, True])
def test_device_engine_adam_preconditioned_statistics(graphed):
    """``scaling_rule=AdamScale()`` (Adam-preconditioned gradient statistics,
    reference gradient_noise_scale.py:289-330) stays on the device engine:
    the bucket kernels read the fused optimizer's second moments in place.
    Same trajectory as the host estimator + torch Adam."""
    adl = _single_process_runtime()
    from adaptdl_b200.torch.scaling_rules import AdamScale

    def make(model):
        return torch.optim.Adam([{"params": [p]} for p in model.parameters()],
                                lr=2e-3, betas=(0.9, 0.98))
    pa, ga, net_a, tr, la = _train(adl, make, AdamScale, True, graphed,
                                   steps=16)
    pb, gb, net_b, _, lb = _train(adl, make, AdamScale, False, False,
                                  steps=16)
    assert net_a.engine is not None and net_a.engine.precondition_stats
    assert net_b.engine is None
    if graphed:
        assert tr.replays > 0
    torch.testing.assert_close(la, lb, rtol=2e-3, atol=2e-4)
    torch.testing.assert_close(pa, pb, rtol=2e-3, atol=2e-4)
    for key in ("sqr_avg", "var_avg", "progress"):
        np.testing.assert_allclose(ga[key], gb[key], rtol=1e-2, atol=1e-7)
    assert ga["progress"] > 0


@pytest.mark.parametrize("graphed", [False, True])
def test_device_engine_with_grad_scaler(graphed):
    """fp16 autocast + ``GradScaler``: the engine stays on (no blocking
    ``.item()`` in ``scaler.step``), non-finite steps are skipped on the
    device, the statistics are divided by the loss scale. Same trajectory as
    the host path (which synchronises every step)."""
    adl = _single_process_runtime()
    from adaptdl_b200.torch.scaling_rules import AdaScale

    def run(fused, graph):
        # a large initial scale overflows fp16 in the first steps: both
        # paths must skip exactly those updates
        scaler = torch.amp.GradScaler("cuda", init_scale=2.0 ** 22,
                                      growth_interval=4)
        out = _train(adl, _sgd, AdaScale, fused, graph, steps=16,
                     autocast=torch.float16, scaler=scaler)
        return out + (scaler,)
    pa, ga, net_a, tr, la, sa = run(True, graphed)
    pb, gb, net_b, _, lb, sb = run(False, False)
    assert net_a.engine is not None and net_b.engine is None
    if graphed:
        assert tr.replays > 0
    assert sa.get_scale() == sb.get_scale() < 2.0 ** 22
    torch.testing.assert_close(la, lb, rtol=5e-3, atol=5e-3)
    torch.testing.assert_close(pa, pb, rtol=5e-3, atol=5e-4)
    for key in ("sqr_avg", "var_avg", "progress"):
        np.testing.assert_allclose(ga[key], gb[key], rtol=5e-2, atol=1e-7)
    assert ga["progress"] > 0


def test_step_profile_comes_from_device_stamps():
    """The goodput profile's step / sync durations are the finalize
    kernel's %globaltimer stamps (not host clocks), booked asynchronously;
    reference: torch/_metrics.py:43-59,104-127."""
    adl = _single_process_runtime()
    from adaptdl_b200.torch import _metrics, data as adl_data
    from adaptdl_b200.torch.scaling_rules import AdaScale
    _metrics._reset_for_tests()
    # "the training loader" is a per-process singleton (first wins): earlier
    # tests of this process own it, and only its steps are committed
    adl_data.AdaptiveDataLoaderHelper._training = None
    _, _, net, _, _ = _train(adl, _sgd, AdaScale, True, True, steps=20)
    timer = _metrics.device_timer()
    assert timer is not None and timer.reducer is net.reducer
    _metrics._book_device_records(wait=True)
    assert timer.booked >= 10 and timer.dropped == 0
    rows = [r for r in _metrics._metrics_state().profile.values()
            if r.get("optim_count")]
    assert rows
    for row in rows:
        mean_step = row["optim_step_time"] / row["optim_count"]
        mean_sync = row["optim_sync_time"] / row["optim_count"]
        # a device-timed step of this toy model: tens of microseconds to a
        # few milliseconds, never the host's python loop time
        assert 5e-6 < mean_step < 2e-2, mean_step
        assert 0 < mean_sync <= mean_step
    # accumulation micro-steps arrive with the optimizer step closing them
    _metrics._reset_for_tests()
    adl_data.AdaptiveDataLoaderHelper._training = None
    _, _, net2, _, _ = _train(adl, _sgd, AdaScale, True, False, steps=8,
                              accum=True)
    _metrics._book_device_records(wait=True)
    rows = [r for r in _metrics._metrics_state().profile.values()
            if r.get("optim_count")]
    assert rows and any(r.get("accum_count", 0) > 0 for r in rows)
    for row in rows:
        if row.get("accum_count"):
            assert row["accum_step_time"] > 0
    _metrics._reset_for_tests()


@pytest.mark.parametrize("opt_name", ["adam", "adamw", "sgd_plain"])
def test_fused_optimizers_match_torch(opt_name):
    adl = _single_process_runtime()
    from adaptdl_b200.torch.scaling_rules import LinearScale

    def make(model):
        if opt_name == "adam":
            return torch.optim.Adam(model.parameters(), lr=1e-3,
                                    weight_decay=1e-2)
        if opt_name == "adamw":
            return torch.optim.AdamW(
                [{"params": list(model.parameters())[:2], "lr": 2e-3},
                 {"params": list(model.parameters())[2:]}],
                lr=1e-3, weight_decay=5e-2, betas=(0.8, 0.95))
        return torch.optim.SGD(model.parameters(), lr=0.1)
    pa, _, _, _, la = _train(adl, make, LinearScale, True, False, steps=10)
    pb, _, _, _, lb = _train(adl, make, LinearScale, False, False, steps=10)
    torch.testing.assert_close(la, lb, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(pa, pb, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("accum", [False, True])
def test_graphed_step_matches_eager(accum):
    adl = _single_process_runtime()
    from adaptdl_b200.torch.scaling_rules import AdaScale

    def bn_model():
        return torch.nn.Sequential(
            torch.nn.Linear(32, 64), torch.nn.BatchNorm1d(64),
            torch.nn.ReLU(), torch.nn.Linear(64, 10))
    pa, ga, net_a, tr_a, la = _train(adl, _sgd, AdaScale, True, True,
                                     steps=12, accum=accum,
                                     model_fn=bn_model)
    pb, gb, net_b, tr_b, lb = _train(adl, _sgd, AdaScale, True, False,
                                     steps=12, accum=accum,
                                     model_fn=bn_model)
    assert tr_a.replays >= 6 and tr_b.replays == 0
    torch.testing.assert_close(la, lb, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(pa, pb, rtol=1e-5, atol=1e-6)
    for key in ("sqr_avg", "var_avg", "progress"):
        np.testing.assert_allclose(ga[key], gb[key], rtol=1e-6)
    # the host mirror trails the device but is alive
    assert net_a.gns.get_progress() > 0
    assert net_a.reducer.launches > 0


def test_checkpoint_roundtrip_with_device_engine(tmp_path):
    """optimizer.state_dict() sees the live device state (GNS averages,
    momentum buffers) and a reload restores it."""
    adl = _single_process_runtime()
    from adaptdl_b200.torch.scaling_rules import AdaScale
    _, gns, net, _, _ = _train(adl, _sgd, AdaScale, True, False, steps=6)
    import io
    buf = io.BytesIO()
    net._state.sync()
    net._state.save(buf)
    sd = torch.load(io.BytesIO(buf.getvalue()), weights_only=False)
    optim_sd = sd[0][1]
    np.testing.assert_allclose(optim_sd["state"]["gns"]["sqr_avg"],
                               gns["sqr_avg"])
    first = next(iter(k for k in optim_sd["state"] if k != "gns"))
    assert optim_sd["state"][first]["momentum_buffer"].abs().sum() > 0
    # perturb, reload, verify restoration
    before = torch.cat([p.detach().reshape(-1).clone()
                        for p in net.module.parameters()])
    with torch.no_grad():
        for p in net.module.parameters():
            p.add_(1.0)
    net._state.load(io.BytesIO(buf.getvalue()))
    net.engine.adopt_optimizer_state()
    net.engine.push_gns_state(net._state.optimizer.state["gns"])
    after = torch.cat([p.detach().reshape(-1)
                       for p in net.module.parameters()])
    torch.testing.assert_close(before, after)
    probe = {}
    net.engine.pull_gns_state(probe)
    np.testing.assert_allclose(probe["sqr_avg"], gns["sqr_avg"])


@pytest.mark.parametrize("opt_name", ["sgd", "adamw"])
def test_mixed_precision_params_follow_fp32_training(opt_name, tmp_path):
    """bf16-stored weights + fp32 masters inside the fused optimizer train like
    fp32 weights under bf16 autocast (same rounded weights in every forward),
    and a checkpoint restores the masters exactly."""
    adl = _single_process_runtime()
    from adaptdl_b200.torch.scaling_rules import LinearScale

    def mlp():
        return torch.nn.Sequential(
            torch.nn.Linear(32, 64), torch.nn.ReLU(),
            torch.nn.Linear(64, 64), torch.nn.Tanh(),
            torch.nn.Linear(64, 10))

    def mlp16():
        return adl.mixed_precision_params(mlp())

    def make(model):
        if opt_name == "sgd":
            return _sgd(model)
        return torch.optim.AdamW(model.parameters(), lr=2e-3,
                                 weight_decay=1e-2)
    pa, _, net_a, _, la = _train(adl, make, LinearScale, True, False,
                                 steps=12, model_fn=mlp16,
                                 autocast=torch.bfloat16)
    pb, _, net_b, _, lb = _train(adl, make, LinearScale, True, False,
                                 steps=12, model_fn=mlp,
                                 autocast=torch.bfloat16)
    weights = [p for p in net_a.module.parameters() if p.dim() >= 2]
    assert all(p.dtype == torch.bfloat16 for p in weights)
    assert all(p.grad.dtype == torch.bfloat16 for p in weights)
    masters = torch.cat([
        net_a._state.optimizer.state[p]["master_param"].reshape(-1)
        if p.dim() >= 2 else p.detach().reshape(-1)
        for p in net_a.module.parameters()])
    assert masters.dtype == torch.float32
    # bf16 forward/backward: the two runs round differently here and there
    torch.testing.assert_close(la, lb, rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(masters, pb, rtol=5e-2, atol=8e-3)
    # the 16-bit weights are the rounded masters
    for p in weights:
        m = net_a._state.optimizer.state[p]["master_param"]
        assert torch.equal(p.detach(), m.to(torch.bfloat16))
        assert (m - p.detach().float()).abs().max() > 0   # masters carry more
    # checkpoint: masters survive exactly (Optimizer.load_state_dict alone
    # would round them to bf16)
    import io
    buf = io.BytesIO()
    net_a._state.sync()
    net_a._state.save(buf)
    before = masters.clone()
    with torch.no_grad():
        for p in weights:
            net_a._state.optimizer.state[p]["master_param"].add_(1.0)
            p.add_(1.0)
    net_a._state.load(io.BytesIO(buf.getvalue()))
    net_a.engine.adopt_optimizer_state()
    net_a.engine.load_wide_state(net_a._state.wide_state)
    after = torch.cat([
        net_a._state.optimizer.state[p]["master_param"].reshape(-1)
        if p.dim() >= 2 else p.detach().reshape(-1)
        for p in net_a.module.parameters()])
    assert torch.equal(before, after)
    for p in weights:
        m = net_a._state.optimizer.state[p]["master_param"]
        assert torch.equal(p.detach(), m.to(torch.bfloat16))


# ---------------------------------------------------------------------------
# tcgen05 Linear + bias + GELU (csrc/adl_gemm.cu)
# ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("m,n,k,block_n,cluster_m", [
    (128, 128, 64, 128, 1), (256, 256, 128, 256, 2), (1000, 384, 192, 128, 4),
    (4096, 3072, 768, 256, 4), (4096, 3072, 768, 128, 2),
    (4096, 3072, 768, 256, 1), (77, 512, 1024, 0, 0), (640, 256, 64, 256, 4),
    (300, 768, 3072, 0, 0), (4096, 3072, 768, 256, 22), (256, 256, 64, 256, 22),
    (1000, 512, 192, 256, 22), (128, 256, 128, 256, 22)])
def test_tcgen05_linear_gelu_forward(m, n, k, block_n, cluster_m):
    from adaptdl_b200.ops import check_errors, gemm_bias_act
    torch.manual_seed(m + n + k)
    dev = torch.device("cuda:0")
    x = torch.randn(m, k, device=dev).bfloat16()
    w = (torch.randn(n, k, device=dev) / k ** 0.5).bfloat16()
    b = torch.randn(n, device=dev)
    y, z = gemm_bias_act(x, w, b, "gelu", True, block_n=block_n,
                         cluster_m=cluster_m)
    torch.cuda.synchronize()
    check_errors()
    ref_z = x.float() @ w.float().t() + b
    ref_y = torch.nn.functional.gelu(ref_z)
    # bf16 output rounding: 2^-8 relative
    assert torch.allclose(z.float(), ref_z, rtol=1e-2, atol=2e-2)
    assert torch.allclose(y.float(), ref_y, rtol=1e-2, atol=2e-2)
    y2, z2 = gemm_bias_act(x, w, None, "identity", False, block_n=block_n,
                           cluster_m=cluster_m)
    assert z2 is None
    assert torch.allclose(y2.float(), x.float() @ w.float().t(),
                          rtol=1e-2, atol=2e-2)
    check_errors()


@pytest.mark.gpu
def test_tcgen05_linear_gelu_autograd_matches_torch():
    from adaptdl_b200.ops import LinearGELU, check_errors
    torch.manual_seed(3)
    dev = torch.device("cuda:0")
    fused = LinearGELU(256, 512).to(dev)
    plain = torch.nn.Linear(256, 512).to(dev)
    plain.load_state_dict(fused.state_dict())
    x1 = torch.randn(4, 96, 256, device=dev, requires_grad=True)
    x2 = x1.detach().clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y1 = fused(x1)
        y2 = torch.nn.functional.gelu(plain(x2))
    assert y1.dtype == torch.bfloat16
    g = torch.randn_like(y1)
    y1.backward(g)
    y2.backward(g)
    check_errors()
    assert torch.allclose(y1.float(), y2.float(), rtol=2e-2, atol=2e-2)
    for a, b in ((x1.grad, x2.grad), (fused.weight.grad, plain.weight.grad),
                 (fused.bias.grad, plain.bias.grad)):
        assert a.dtype == b.dtype and a.shape == b.shape
        scale = b.abs().max().item()
        assert (a - b).abs().max().item() <= 3e-2 * scale + 1e-3


# ---------------------------------------------------------------------------
# fused BatchNorm (+ residual) (+ ReLU), channels-last (csrc/adl_bn.cu)
# ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,residual,relu", [
    ((32, 64, 16, 16), True, True), ((16, 128, 8, 8), False, True),
    ((8, 512, 4, 4), True, False), ((7, 256, 5, 3), False, False),
    ((128, 64, 32, 32), True, True)])
def test_fused_bn_act_matches_torch(dtype, shape, residual, relu):
    from adaptdl_b200.ops import BatchNormAct2d
    from adaptdl_b200.ops.bn_act import supported
    torch.manual_seed(1)
    dev = torch.device("cuda:0")
    c = shape[1]
    fused = BatchNormAct2d(c).to(dev)
    with torch.no_grad():
        fused.weight.uniform_(0.5, 1.5)
        fused.bias.normal_()
    plain = torch.nn.BatchNorm2d(c).to(dev)
    plain.load_state_dict(fused.state_dict())
    x32 = torch.randn(shape, device=dev) * 2 + 0.5
    x1 = x32.to(dtype).contiguous(memory_format=torch.channels_last) \
        .requires_grad_(True)
    assert supported(x1)
    x2 = x1.detach().float().requires_grad_(True)
    r1 = r2 = None
    if residual:
        r1 = torch.randn(shape, device=dev).to(dtype).contiguous(
            memory_format=torch.channels_last).requires_grad_(True)
        r2 = r1.detach().float().requires_grad_(True)
    y1 = fused(x1, r1, relu)
    y2 = plain(x2)                       # fp32 reference of the same op
    if residual:
        y2 = y2 + r2
    if relu:
        y2 = torch.relu(y2)
    g = torch.randn(shape, device=dev)
    y1.backward(g.to(dtype).contiguous(memory_format=torch.channels_last))
    y2.backward(g.to(dtype).float())
    tol = 1e-4 if dtype == torch.float32 else 3e-2
    assert y1.dtype == dtype
    assert y1.is_contiguous(memory_format=torch.channels_last)
    assert torch.allclose(y1.float(), y2, rtol=tol, atol=tol)
    assert torch.allclose(fused.running_mean, plain.running_mean,
                          rtol=1e-4, atol=1e-4)
    assert torch.allclose(fused.running_var, plain.running_var,
                          rtol=1e-3, atol=1e-3)
    assert int(fused.num_batches_tracked) == 1

    def close(a, b, t):
        scale = b.abs().max().item() + 1e-6
        return (a.float() - b).abs().max().item() <= t * scale
    gtol = 2e-4 if dtype == torch.float32 else 4e-2
    assert close(x1.grad, x2.grad, gtol)
    assert close(fused.weight.grad, plain.weight.grad, gtol)
    assert close(fused.bias.grad, plain.bias.grad, gtol)
    if residual:
        assert close(r1.grad, r2.grad, gtol)




@pytest.mark.gpu
def test_fused_bn_resnet_matches_unfused_model():
    """ResNet-18 with the fused BN kernels vs the same model with the fused
    path switched off (PyTorch composition): same loss and gradients."""
    import os
    from adaptdl_b200.models import resnet18
    torch.manual_seed(5)
    # TF32 convolutions turn 1e-7 input differences into 1e-3 ones
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    net = resnet18().to(dev).to(memory_format=torch.channels_last)
    x = torch.randn(32, 3, 32, 32, device=dev).contiguous(
        memory_format=torch.channels_last)
    t = torch.randint(0, 10, (32,), device=dev)
    results = []
    for flag in ("1", "0"):
        os.environ["ADAPTDL_B200_FUSED_BN"] = flag
        state = {k: v.clone() for k, v in net.state_dict().items()}
        net.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(net(x), t)
        loss.backward()
        results.append((loss.item(), [p.grad.clone()
                                      for p in net.parameters()]))
        net.load_state_dict(state)
    os.environ.pop("ADAPTDL_B200_FUSED_BN")
    torch.backends.cudnn.allow_tf32 = True
    assert abs(results[0][0] - results[1][0]) < 1e-3
    # a ReLU flip at an activation that is zero to rounding moves single
    # elements; compare in norm
    for a, b in zip(results[0][1], results[1][1]):
        assert (a - b).norm().item() <= 1e-2 * (b.norm().item() + 1e-6)


# ---------------------------------------------------------------------------
# fused dropout + residual + LayerNorm (csrc/adl_ln.cu)
# ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,p", [((4, 128, 768), 0.1), ((37, 200), 0.2),
                                     ((3, 5, 1024), 0.0), ((16, 8), 0.5),
                                     ((512, 512), 0.1)])
def test_fused_dropout_add_layer_norm(dtype, shape, p, monkeypatch):
    monkeypatch.setenv("ADAPTDL_B200_FUSED_LN", "1")     # opt-in op
    from adaptdl_b200.ops import dropout_add_layer_norm
    from adaptdl_b200.ops.layer_norm import supported
    torch.manual_seed(7)
    dev = torch.device("cuda:0")
    d = shape[-1]
    if dtype == torch.float32 and d > 512:
        pytest.skip("fp32 rows wider than 512 use the PyTorch composition")
    w = (torch.rand(d, device=dev) + 0.5).requires_grad_(True)
    b = torch.randn(d, device=dev).requires_grad_(True)
    w2 = w.detach().clone().requires_grad_(True)
    b2 = b.detach().clone().requires_grad_(True)
    x1 = (torch.randn(shape, device=dev) * 2).to(dtype).requires_grad_(True)
    h1 = torch.randn(shape, device=dev).to(dtype).requires_grad_(True)
    assert supported(h1)
    x2 = x1.detach().float().requires_grad_(True)
    h2 = h1.detach().float().requires_grad_(True)
    mask = (torch.rand(shape, device=dev) > p).to(torch.uint8)
    y1 = dropout_add_layer_norm(x1, h1, w, b, p, True, 1e-5, mask=mask)
    z2 = x2 + (h2 * mask.float() / (1 - p) if p > 0 else h2)
    # the kernel rounds z to the storage dtype (straight-through: gradient 1)
    z2 = z2 + (z2.to(dtype).float() - z2).detach()
    y2 = torch.nn.functional.layer_norm(z2, (d,), w2, b2, 1e-5)
    g = torch.randn(shape, device=dev)
    y1.backward(g.to(dtype))
    y2.backward(g.to(dtype).float())
    tol = 2e-4 if dtype == torch.float32 else 3e-2
    assert y1.dtype == dtype and y1.shape == tuple(shape)
    assert torch.allclose(y1.float(), y2, rtol=tol, atol=tol)

    def close(a, c, t):
        return (a.float() - c).abs().max().item() <= \
            t * (c.abs().max().item() + 1e-6)
    gtol = 5e-4 if dtype == torch.float32 else 4e-2
    assert close(x1.grad, x2.grad, gtol)
    assert close(h1.grad, h2.grad, gtol)
    assert close(w.grad, w2.grad, gtol)
    assert close(b.grad, b2.grad, gtol)


@pytest.mark.gpu
def test_fused_dropout_add_layer_norm_random_mask_and_eval(monkeypatch):
    monkeypatch.setenv("ADAPTDL_B200_FUSED_LN", "1")
    from adaptdl_b200.ops import dropout_add_layer_norm
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    x = torch.randn(64, 256, device=dev)
    h = torch.randn(64, 256, device=dev)
    w = torch.ones(256, device=dev)
    b = torch.zeros(256, device=dev)
    ref = torch.nn.functional.layer_norm(x + h, (256,), w, b, 1e-5)
    out = dropout_add_layer_norm(x, h, w, b, 0.3, False)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4)
    a = dropout_add_layer_norm(x, h, w, b, 0.3, True)
    c = dropout_add_layer_norm(x, h, w, b, 0.3, True)
    assert not torch.equal(a, c)            # fresh keep-mask per call
    assert torch.isfinite(a).all()




# ---------------------------------------------------------------------------
# transformer layout / reduction ops (csrc/adl_transformer.cu)
# ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(32, 128, 12, 64), (3, 7, 2, 8),
                                   (2, 130, 5, 128)])
def test_split_and_merge_heads_match_views(shape):
    from adaptdl_b200.ops.transformer import merge_heads, split_heads
    n, s, h, d = shape
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    qkv = torch.randn(n, s, 3 * h * d, device=dev).bfloat16() \
        .requires_grad_(True)
    ref = qkv.detach().clone().requires_grad_(True)
    q, k, v = split_heads(qkv, h, 3)
    rq, rk, rv = ref.view(n, s, 3, h, d).permute(2, 0, 3, 1, 4)
    for a, b in ((q, rq), (k, rk), (v, rv)):
        assert a.is_contiguous() and a.shape == (n, h, s, d)
        assert torch.equal(a, b)
    out = merge_heads(q * 2 + k - v)
    rout = (rq * 2 + rk - rv).transpose(1, 2).reshape(n, s, h * d)
    assert out.shape == (n, s, h * d) and torch.equal(out, rout)
    g = torch.randn_like(out)
    out.backward(g)
    rout.backward(g)
    assert torch.equal(qkv.grad, ref.grad)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(4096, 768), (4096, 3072), (37, 40),
                                   (1, 8), (5000, 2304)])
def test_colsum_matches_torch(dtype, shape):
    from adaptdl_b200.ops.transformer import colsum
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    x = torch.randn(shape, device=dev).to(dtype)
    got = colsum(x)
    want = x.double().sum(0)
    assert got.dtype == torch.float32 and got.shape == (shape[1],)
    tol = 1e-5 * (shape[0] ** 0.5) + 1e-6
    assert (got.double() - want).abs().max().item() <= \
        tol * (1 + want.abs().max().item())
    assert torch.equal(colsum(x), got)          # deterministic


@pytest.mark.gpu
def test_linear_with_fused_bias_grad_and_padded_logits():
    from adaptdl_b200.ops.transformer import linear, padded_logits
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    x = torch.randn(6, 33, 64, device=dev)
    w = torch.randn(96, 64, device=dev).bfloat16().requires_grad_(True)
    b = torch.randn(96, device=dev).requires_grad_(True)
    w2, b2 = (t.detach().clone().requires_grad_(True) for t in (w, b))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = linear(x, w, b)
        y2 = torch.nn.functional.linear(x, w2, b2)
    assert y.dtype == torch.bfloat16
    torch.testing.assert_close(y, y2)
    g = torch.randn_like(y)
    y.backward(g)
    y2.backward(g)
    torch.testing.assert_close(w.grad, w2.grad, rtol=2e-2, atol=2e-2)
    assert b.grad.dtype == torch.float32
    # the fp32 column sum is MORE accurate than the bf16 reduction autograd
    # does under autocast: compare against the fp64 truth
    truth = g.double().reshape(-1, 96).sum(0)
    assert (b.grad.double() - truth).abs().max() <= \
        (b2.grad.double() - truth).abs().max() + 1e-3

    # vocabulary of 100 (not a multiple of 8): padded GEMM, fp32 logits
    n = 100
    w = torch.randn(n, 64, device=dev).bfloat16().requires_grad_(True)
    b = torch.randn(n, device=dev).requires_grad_(True)
    w2, b2 = (t.detach().clone().requires_grad_(True) for t in (w, b))
    xs = torch.randn(5, 9, 64, device=dev, requires_grad=True)
    xs2 = xs.detach().clone().requires_grad_(True)
    tgt = torch.randint(0, n, (45,), device=dev)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = padded_logits(xs, w, b)
        loss = torch.nn.functional.cross_entropy(logits.view(-1, n), tgt)
        ref = torch.nn.functional.linear(xs2, w2, b2)
        loss2 = torch.nn.functional.cross_entropy(ref.view(-1, n), tgt)
    assert logits.dtype == torch.float32 and logits.is_contiguous()
    assert logits.shape == (5, 9, n)
    torch.testing.assert_close(logits, ref.float(), rtol=2e-2, atol=2e-2)
    loss.backward()
    loss2.backward()
    for a, c in ((xs.grad, xs2.grad), (w.grad, w2.grad), (b.grad, b2.grad)):
        torch.testing.assert_close(a.float(), c.float(), rtol=3e-2,
                                   atol=3e-3)


@pytest.mark.gpu
def test_bert_layer_with_fused_layout_ops_matches_plain_composition(
        monkeypatch):
    """One encoder layer + MLM head: the dedicated layout / reduction kernels
    against the same model with them switched off."""
    from adaptdl_b200.models.bert import MLMTask
    dev = torch.device("cuda:0")
    results = []
    for flag in ("1", "0"):
        monkeypatch.setenv("ADAPTDL_B200_FUSED_TRANSFORMER", flag)
        torch.manual_seed(3)
        net = MLMTask(1004, 128, 2, 256, 1, dropout=0.0, max_len=64).to(dev)
        x = torch.randint(0, 1004, (4, 48), device=dev)
        t = torch.randint(0, 1004, (4, 48), device=dev)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(x)
            loss = torch.nn.functional.cross_entropy(
                out.view(-1, 1004), t.view(-1))
        loss.backward()
        results.append((loss.item(),
                        [p.grad.float().clone() for p in net.parameters()]))
    assert abs(results[0][0] - results[1][0]) < 2e-2
    for a, b in zip(results[0][1], results[1][1]):
        assert (a - b).norm() <= 5e-2 * (b.norm() + 1e-3)


@pytest.mark.gpu
def test_linear_gelu_dropout_backward_is_one_fused_pass():
    """``linear_act(..., dropout_p)``: forward = GEMM(+bias+GELU) + dropout,
    backward = ONE kernel for the dropout and GELU derivatives; checked
    against the PyTorch composition replayed with the same keep-mask."""
    from adaptdl_b200.ops import linear_act
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    m, k, n, p = 256, 128, 256, 0.3
    x = (torch.randn(m, k, device=dev) * 0.5).bfloat16().requires_grad_(True)
    w = (torch.randn(n, k, device=dev) * 0.1).bfloat16().requires_grad_(True)
    b = torch.randn(n, device=dev).requires_grad_(True)
    y = linear_act(x, w, b, "gelu", dropout_p=p, training=True)
    keep = (y != 0)
    frac = keep.float().mean().item()
    assert 0.6 < frac < 0.8                      # ~ 1 - p
    x2, w2, b2 = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    ref = torch.nn.functional.gelu(x2 @ w2.t() + b2) * keep / (1 - p)
    torch.testing.assert_close(y.float(), ref, rtol=3e-2, atol=3e-2)
    g = torch.randn(m, n, device=dev)
    y.backward(g.bfloat16())
    ref.backward(g.bfloat16().float())
    for a, c in ((x.grad, x2.grad), (w.grad, w2.grad), (b.grad, b2.grad)):
        assert (a.float() - c).norm() <= 3e-2 * c.norm()
    # eval mode: no dropout
    y_eval = linear_act(x, w, b, "gelu", dropout_p=p, training=False)
    assert (y_eval != 0).float().mean().item() > 0.95
