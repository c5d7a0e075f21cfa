"""The goodput model and its fit against the reference's own module.

``baseline/_ref/adaptdl/goodput.py`` (the unmodified reference file; it needs
only numpy / scipy and ``autograd``, for which ``baseline/shims`` carries a
finite-difference stand-in) is loaded under a private name next to
``adaptdl_b200.goodput``: evaluation, the batch-size optimiser and the
performance-model fit must agree on random inputs.
"""

import importlib.util
import os
import sys

import numpy as np
import pytest

from adaptdl_b200 import goodput as own

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FILE = os.path.join(ROOT, "baseline", "_ref", "adaptdl", "goodput.py")

pytestmark = pytest.mark.skipif(
    not os.path.exists(REF_FILE),
    reason="reference package not installed (baseline/install_reference.sh)")


@pytest.fixture(scope="module")
def ref():
    shims = os.path.join(ROOT, "baseline", "shims")
    sys.path.insert(0, shims)
    try:
        spec = importlib.util.spec_from_file_location("_reference_goodput",
                                                      REF_FILE)
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
    finally:
        sys.path.remove(shims)
        for name in [n for n in sys.modules if n.split(".")[0] == "autograd"]:
            del sys.modules[name]
    # the reference predates numpy 1.24 (``np.int``): alias it for the
    # duration of this module's tests only
    patch = pytest.MonkeyPatch()
    patch.setattr(np, "int", int, raising=False)
    patch.setattr(np, "float", float, raising=False)
    yield module
    patch.undo()


def _random_models(rng, count):
    for _ in range(count):
        perf = dict(alpha_c=rng.uniform(0.01, 0.3), beta_c=rng.uniform(1e-4, 5e-3),
                    alpha_n=rng.uniform(0.0, 0.3), beta_n=rng.uniform(0.0, 0.05),
                    alpha_r=rng.uniform(0.0, 0.1), beta_r=rng.uniform(0.0, 0.02),
                    gamma=rng.uniform(1.0, 8.0))
        grad = dict(sqr=rng.uniform(1e-3, 1.0), var=rng.uniform(1e-3, 50.0))
        yield perf, grad, int(rng.choice([32, 128, 512]))


def test_evaluate_and_optimize_agree_with_the_reference(ref):
    rng = np.random.default_rng(7)
    for perf, grad, init_bsz in _random_models(rng, 25):
        theirs = ref.GoodputFunction(ref.PerfParams(**perf),
                                     ref.GradParams(**grad), init_bsz)
        ours = own.GoodputFunction(
            own.PerfParams(**perf),
            own.GradParams(grad["sqr"], grad["var"]), init_bsz)
        nodes = np.array([1, 1, 2, 4, 4])
        replicas = np.array([1, 4, 8, 16, 32])
        bsz = np.maximum(init_bsz // replicas, 1) * np.array([1, 2, 1, 3, 2])
        accum = np.array([0, 1, 0, 2, 0])
        # evaluate() refuses batch sizes below the initial one
        bsz = np.maximum(bsz, -(-init_bsz // (replicas * (accum + 1))))
        np.testing.assert_allclose(
            ours.evaluate(nodes, replicas, bsz, accum),
            theirs.evaluate(nodes, replicas, bsz, accum), rtol=1e-9)
        np.testing.assert_allclose(
            ours.throughput(nodes, replicas, bsz, accum),
            theirs.throughput(nodes, replicas, bsz, accum), rtol=1e-9)
        total = replicas * bsz * (accum + 1)
        np.testing.assert_allclose(ours.efficiency(total),
                                   theirs.efficiency(total), rtol=1e-9)
        for accumulation in (False, True):
            kwargs = dict(max_batch_size=32 * init_bsz,
                          atomic_bsz_range=(max(init_bsz // 64, 1),
                                            8 * init_bsz),
                          accumulation=accumulation)
            g1, b1, s1 = ours.optimize(nodes, replicas, **kwargs)
            g2, b2, s2 = theirs.optimize(nodes, replicas, **kwargs)
            np.testing.assert_allclose(g1, g2, rtol=1e-9)
            np.testing.assert_array_equal(b1, b2)
            np.testing.assert_array_equal(s1, s2)
        # scalar form, as the data loader calls it
        assert ours.optimize(2, 8, max_batch_size=16 * init_bsz,
                             atomic_bsz_range=(1, 4 * init_bsz),
                             accumulation=True) == pytest.approx(
            theirs.optimize(2, 8, max_batch_size=16 * init_bsz,
                            atomic_bsz_range=(1, 4 * init_bsz),
                            accumulation=True))


def _profile(perf, rng, noise=0.0):
    """Synthetic step-time observations from a known performance model."""
    rows = []
    for nodes, replicas in ((1, 1), (1, 2), (1, 4), (2, 8), (4, 16)):
        for bsz in (16, 32, 64, 128):
            rows.append((nodes, replicas, bsz))
    nodes, replicas, bsz = (np.array(c) for c in zip(*rows))
    accum_time = perf["alpha_c"] + perf["beta_c"] * bsz
    alpha = np.where(nodes > 1, perf["alpha_n"], perf["alpha_r"])
    beta = np.where(nodes > 1, perf["beta_n"], perf["beta_r"])
    network = np.where(replicas > 1,
                       alpha + beta * np.maximum(replicas - 2, 1e-8), 1e-8)
    gamma = perf["gamma"]
    optim = (accum_time ** gamma + network ** gamma) ** (1 / gamma)
    jitter = np.exp(noise * rng.standard_normal(len(rows)))
    return nodes, replicas, bsz, accum_time * jitter, optim * jitter


def test_perf_fit_lands_where_the_reference_fit_does(ref):
    rng = np.random.default_rng(11)
    for _ in range(4):
        perf = dict(alpha_c=rng.uniform(0.02, 0.2), beta_c=rng.uniform(5e-4, 4e-3),
                    alpha_n=rng.uniform(0.05, 0.3), beta_n=rng.uniform(0.005, 0.03),
                    alpha_r=rng.uniform(0.01, 0.08), beta_r=rng.uniform(0.001, 0.01),
                    gamma=rng.uniform(1.2, 4.0))
        args = _profile(perf, rng, noise=0.01)
        mine = own.fit_perf_params(*args)
        theirs = ref.fit_perf_params(*args)
        # the two optimisers (analytic vs finite-difference gradients) need
        # not stop at the same point of a flat valley: compare what the
        # parameters PREDICT on the observed configurations
        nodes, replicas, bsz, accum_time, optim_time = args

        def predict(module, params):
            fn = module.GoodputFunction(
                params, module.GradParams(1.0, 1.0), 16)
            return fn.throughput(nodes, replicas, bsz, 0)

        truth = replicas * bsz / optim_time
        np.testing.assert_allclose(predict(own, mine), truth, rtol=0.08)
        # (the reference's fit differentiates through the finite-difference
        # stand-in for ``autograd`` in baseline/shims)
        ref_pred = predict(ref, theirs)
        np.testing.assert_allclose(ref_pred, truth, rtol=0.08)
        np.testing.assert_allclose(predict(own, mine), ref_pred, rtol=0.08)


REF_SPEEDUP = "/root/reference/sched/adaptdl_sched/policy/speedup.py"


@pytest.mark.skipif(not os.path.exists(REF_SPEEDUP),
                    reason="reference checkout not available")
def test_speedup_function_agrees_with_the_reference(ref):
    """The scheduler's speedup function (numpy-only in the reference, loaded
    straight from the checkout) on top of either goodput model."""
    from adaptdl_b200.sched.policy import SpeedupFunction
    spec = importlib.util.spec_from_file_location("_reference_speedup",
                                                  REF_SPEEDUP)
    ref_speedup = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_speedup)
    rng = np.random.default_rng(3)
    nodes = np.array([1, 1, 1, 2, 2, 3, 4, 8, 0, 1])
    replicas = np.array([1, 2, 4, 4, 8, 12, 16, 64, 0, 3])
    for perf, grad, init_bsz in _random_models(rng, 12):
        for accumulation in (False, True):
            kwargs = dict(max_batch_size=16 * init_bsz,
                          atomic_bsz_range=(max(init_bsz // 32, 1),
                                            4 * init_bsz),
                          accumulation=accumulation)
            theirs = ref_speedup.SpeedupFunction(
                ref.GoodputFunction(ref.PerfParams(**perf),
                                    ref.GradParams(**grad), init_bsz),
                **kwargs)
            ours = SpeedupFunction(
                own.GoodputFunction(own.PerfParams(**perf),
                                    own.GradParams(grad["sqr"], grad["var"]),
                                    init_bsz), **kwargs)
            np.testing.assert_allclose(ours(nodes, replicas),
                                       theirs(nodes, replicas), rtol=1e-9)
            # memoised second query and the scalar form
            np.testing.assert_allclose(ours(nodes, replicas),
                                       theirs(nodes, replicas), rtol=1e-9)
            assert ours(2, 8) == pytest.approx(theirs(2, 8), rel=1e-9)
            assert ours(1, 1) == pytest.approx(1.0)
