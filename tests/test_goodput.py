"""Goodput model invariants, optimize() under bounds, and fitting."""
import itertools

import numpy as np
import pytest

from adaptdl_b200.goodput import (GoodputFunction, PerfParams, GradParams,
                                  fit_perf_params, _objective)

RNG = np.random.RandomState(0)
PERF_PARAMS = [PerfParams(*RNG.gamma(2.0, 2.0, [7])) for _ in range(10)]
PERF_PARAMS = [p._replace(gamma=1.0 + p.gamma % 9.0) for p in PERF_PARAMS]
GRAD_PARAMS = [GradParams(*RNG.gamma(2.0, 2.0, [2])) for _ in range(10)]


def groupby_indices(*args):
    _, indices = np.unique(np.stack(args), axis=1, return_inverse=True)
    indices = np.asarray(indices).reshape(-1)
    groups = {}
    for i, g in enumerate(indices):
        groups.setdefault(g, []).append(i)
    return list(groups.values())


@pytest.mark.parametrize("perf_params", PERF_PARAMS)
@pytest.mark.parametrize("grad_params", GRAD_PARAMS[:4])
def test_evaluate(perf_params, grad_params):
    init_batch_size = 16
    fn = GoodputFunction(perf_params, grad_params, init_batch_size)
    num_nodes = np.array([1, 2, 3, 4])
    num_replicas = np.array([1, 2, 4, 8])
    atomic_bsz = np.array([init_batch_size // 2, init_batch_size,
                           2 * init_batch_size])
    accum_steps = np.array([0, 1, 2, 3, 4])
    grid = np.array(list(itertools.product(num_nodes, num_replicas,
                                           atomic_bsz, accum_steps)))
    grid = grid[grid[:, 0] <= grid[:, 1]]
    grid = grid[grid[:, 1] * grid[:, 2] * (grid[:, 3] + 1)
                >= init_batch_size]
    num_nodes, num_replicas, atomic_bsz, accum_steps = grid.T
    batch_size = num_replicas * atomic_bsz * (accum_steps + 1)
    goodput = fn(num_nodes, num_replicas, atomic_bsz, accum_steps)
    throughput = fn.throughput(num_nodes, num_replicas, atomic_bsz,
                               accum_steps)
    efficiency = fn.efficiency(batch_size)
    assert np.all(0 <= throughput)
    assert np.all(0 <= efficiency) and np.all(efficiency <= 1 + 1e-9)
    assert np.allclose(goodput, throughput * efficiency)
    # efficiency is non-increasing in the batch size
    order = np.argsort(batch_size, kind="stable")
    assert np.all(np.diff(efficiency[order]) <= 1e-9)
    # throughput grows, sub-linearly, with the local batch size
    for idx in groupby_indices(num_nodes, num_replicas, accum_steps):
        idx = sorted(idx, key=lambda i: atomic_bsz[i])
        t, b = throughput[idx], atomic_bsz[idx]
        assert np.all(np.diff(t) >= -1e-9)
        assert np.all(np.diff(t / b) <= 1e-9)
    # scalability is sub-linear in the replica count
    for idx in groupby_indices(num_nodes, atomic_bsz, accum_steps):
        idx = sorted(idx, key=lambda i: num_replicas[i])
        t, r = throughput[idx], num_replicas[idx]
        assert np.all(np.diff(t / r) <= 1e-9)


@pytest.mark.parametrize("perf_params", PERF_PARAMS[:5])
@pytest.mark.parametrize("grad_params", GRAD_PARAMS[:3])
@pytest.mark.parametrize("accumulation", [False, True])
def test_optimize_bounds(perf_params, grad_params, accumulation):
    fn = GoodputFunction(perf_params, grad_params, 128)
    nodes = np.array([1, 1, 2, 4])[:, None]
    repl = np.array([1, 2, 4, 8, 16])[None, :]
    nodes_b, repl_b = np.broadcast_arrays(nodes, repl)
    keep = nodes_b <= repl_b
    for max_bsz, bounds in [(None, None), (1280, None), (1280, (64, 256)),
                            (None, (64, 256)), (4096, (32, 512))]:
        goodput, bsz, steps = fn.optimize(
            np.where(keep, nodes_b, 1), repl_b, max_batch_size=max_bsz,
            atomic_bsz_range=bounds, accumulation=accumulation)
        assert goodput.shape == bsz.shape == steps.shape == (4, 5)
        assert np.all(goodput >= 0)
        total = repl_b * bsz * (steps + 1)
        assert np.all(total >= 128)
        if not accumulation:
            assert np.all(steps == 0)
        else:
            lonely = np.logical_and(repl_b == 1, total > 128)
            assert np.all(steps[lonely] >= 1)
        if bounds:
            assert np.all(bsz >= bounds[0]) and np.all(bsz <= bounds[1])
        if max_bsz:
            # rounding up atomic_bsz may overshoot by < one sample/replica
            assert np.all(total <= max_bsz + repl_b * (steps + 1))
        # the reported goodput is the goodput of the reported config
        assert np.allclose(goodput, fn(np.where(keep, nodes_b, 1), repl_b,
                                       bsz, steps))
    # scalar in, scalar out
    g, b, s = fn.optimize(1, 4, max_batch_size=1280,
                          atomic_bsz_range=(64, 256),
                          accumulation=accumulation)
    assert np.isscalar(g) and isinstance(b, int) and isinstance(s, int)


def test_optimize_prefers_more_replicas_then_saturates():
    perf = PerfParams(0.121, 0.00568, 0.0236, 0.00634, 0.0118, 0.00317, 1.14)
    fn = GoodputFunction(perf, GradParams(sqr=0.00136, var=0.000502), 128)
    goodputs = [fn.optimize(1, n, max_batch_size=1280,
                            atomic_bsz_range=(64, 256),
                            accumulation=True)[0] for n in (1, 2, 4, 8)]
    assert goodputs == sorted(goodputs)
    speedup = np.array(goodputs) / goodputs[0]
    assert np.all(speedup <= np.array([1, 2, 4, 8]) + 1e-9)


def test_objective_gradient_matches_finite_differences():
    rng = np.random.RandomState(1)
    nodes = np.array([1, 1, 1, 2, 2, 4])
    repl = np.array([1, 2, 4, 4, 8, 16])
    bsz = np.array([32., 64, 128, 32, 64, 128])
    ta = rng.uniform(0.1, 1.0, 6)
    to = ta + rng.uniform(0.01, 0.5, 6)
    for _ in range(5):
        x = np.concatenate([rng.uniform(0.01, 0.3, 6),
                            [rng.uniform(1.1, 5)]])
        value, grad = _objective(x, nodes, repl, bsz, ta, to)
        num = np.zeros(7)
        for i in range(7):
            e = np.zeros(7)
            e[i] = 1e-6 * max(1.0, abs(x[i]))
            num[i] = (_objective(x + e, nodes, repl, bsz, ta, to, False)
                      - _objective(x - e, nodes, repl, bsz, ta, to, False)
                      ) / (2 * e[i])
        assert np.allclose(grad, num, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("noise", [0.001, 0.01])
def test_fit_recovers_own_model(noise):
    # data generated by the model class itself (+ one-sided noise): the fit
    # must reach a loss no worse than 1.1x the generating parameters' loss.
    from adaptdl_b200.goodput import _accum_time, _network_time
    rng = np.random.RandomState(3)
    size = (1000,)
    nodes = rng.randint(low=1, high=11, size=size)
    replicas = rng.randint(low=1, high=nodes + 1, size=size)
    local_bsz = rng.randint(32, 1024, size=size)
    params = PerfParams(0.1, 0.01, 0.5, 1.0, 1e-6, 1e-6, 1.2)
    accum = _accum_time(params, local_bsz) \
        + np.maximum(rng.normal(0, noise, size=size), 0.0)
    network = _network_time(params, nodes, replicas) \
        + np.maximum(rng.normal(0, noise, size=size), 0.0)
    optim = (accum ** params.gamma + network ** params.gamma) \
        ** (1 / params.gamma)
    result = fit_perf_params(nodes, replicas, local_bsz, accum, optim)
    args = (nodes, replicas, local_bsz.astype(float), accum, optim, False)
    loss_result = _objective(result, *args)
    loss_true = _objective(params, *args)
    assert abs(loss_result - loss_true) < 0.1 * loss_true \
        or loss_result < loss_true, result


def test_fit_single_replica_pins_network_terms():
    nodes = np.array([1, 1, 1])
    repl = np.array([1, 1, 1])
    bsz = np.array([32., 64, 128])
    ta = 0.05 + 0.001 * bsz
    fit = fit_perf_params(nodes, repl, bsz, ta, ta)
    assert fit.alpha_r == pytest.approx(1e-8)
    assert fit.beta_r == pytest.approx(1e-8)
    assert fit.alpha_n >= 1.1 * fit.alpha_r * 0.999
    assert abs(fit.alpha_c - 0.05) < 0.02 and abs(fit.beta_c - 0.001) < 3e-4
    # single batch size: alpha_c pinned to half the mean accum time
    fit = fit_perf_params([1], [1], [64.], [0.2], [0.2])
    assert fit.alpha_c == pytest.approx(0.1)


def test_fit_is_robust_to_arbitrary_small_profiles():
    """Whatever (few, noisy, degenerate) measurements a job has collected,
    the fit returns finite parameters inside their bounds."""
    from hypothesis import given, settings, strategies as st

    row = st.tuples(st.integers(1, 4), st.integers(1, 8),
                    st.integers(1, 512),
                    st.floats(1e-4, 10.0), st.floats(0.0, 5.0))

    @settings(max_examples=40, deadline=None)
    @given(st.lists(row, min_size=1, max_size=8))
    def check(rows):
        nodes = np.array([min(r[0], r[1]) for r in rows])
        replicas = np.array([r[1] for r in rows])
        bsz = np.array([r[2] for r in rows])
        accum = np.array([r[3] for r in rows])
        optim = accum + np.array([r[4] for r in rows])   # step >= local time
        params = fit_perf_params(nodes, replicas, bsz, accum, optim)
        values = np.array(params, dtype=float)
        assert np.isfinite(values).all(), params
        assert (values[:6] >= 0).all() and 1.0 <= params.gamma <= 10.0
        fn = GoodputFunction(params, GradParams(1.0, 1.0), 32)
        goodput, atomic, accum_steps = fn.optimize(
            1, 2, max_batch_size=1024, atomic_bsz_range=(8, 256),
            accumulation=True)
        assert np.isfinite(goodput) and goodput >= 0
        assert 8 <= atomic <= 256 and accum_steps >= 0
    check()
