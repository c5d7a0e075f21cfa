"""Goodput = throughput(model of step time) x statistical efficiency(GNS).

The model (parity with reference ``adaptdl/adaptdl/goodput.py``):

    T_accum(m)       = alpha_c + beta_c * m                (one micro-batch)
    T_net(nodes, N)  = alpha_n + beta_n * max(N-2, eps)    if nodes > 1
                       alpha_r + beta_r * max(N-2, eps)    elif N > 1
                       eps     + eps    * max(N-2, eps)    otherwise
    T_optim          = (T_accum^gamma + T_net^gamma)^(1/gamma)
    T_step           = a * T_accum + T_optim               (a = accum steps)
    throughput       = N * m * (a+1) / T_step
    efficiency(B)    = gain(B/B0) / (B/B0),  gain(s) = (var+sqr)/(var/s+sqr)

``fit_perf_params`` fits the seven parameters by L-BFGS-B on the RMS log
error of the accumulation- and optimisation-step times. The reference
differentiates its objective with the ``autograd`` package; here the
gradient is derived by hand (and checked against finite differences in
``tests/test_goodput.py``), so the only dependencies are numpy and scipy.
"""

import collections

import numpy as np
import scipy.optimize

__all__ = ["PerfParams", "GradParams", "GoodputFunction", "fit_perf_params"]

PerfParams = collections.namedtuple("PerfParams", [
    "alpha_c",  # constant term of compute time per micro-batch
    "beta_c",   # compute time per sample
    "alpha_n",  # inter-node all-reduce: constant term
    "beta_n",   # inter-node all-reduce: per-replica retrogression
    "alpha_r",  # intra-node all-reduce: constant term
    "beta_r",   # intra-node all-reduce: per-replica retrogression
    "gamma",    # overlap exponent in [1, 10]: 1 = none, large = perfect
])

GradParams = collections.namedtuple("GradParams", ["sqr", "var"])

_EPS = 1e-8


def _accum_time(p, atomic_bsz):
    return p.alpha_c + p.beta_c * atomic_bsz


def _network_time(p, num_nodes, num_replicas):
    multi_node = np.greater(num_nodes, 1)
    multi_repl = np.greater(num_replicas, 1)
    alpha = np.where(multi_node, p.alpha_n,
                     np.where(multi_repl, p.alpha_r, _EPS))
    beta = np.where(multi_node, p.beta_n,
                    np.where(multi_repl, p.beta_r, _EPS))
    return alpha + beta * np.maximum(np.subtract(num_replicas, 2), _EPS)


def _log_optim_time(gamma, accum_time, network_time):
    # log of the gamma-norm, via log-sum-exp for stability at large gamma.
    # Perfectly free communication (fitted alpha = beta = 0) is clamped to
    # _EPS seconds instead of taking log(0).
    return np.logaddexp(gamma * np.log(np.maximum(accum_time, _EPS)),
                        gamma * np.log(np.maximum(network_time, _EPS))) \
        / gamma


class GoodputFunction(object):
    """Predicts goodput for (num_nodes, num_replicas, atomic_bsz,
    accum_steps) and searches for the best batch-size configuration."""

    def __init__(self, perf_params, grad_params, init_batch_size):
        self._perf = PerfParams(*perf_params)
        self._grad = GradParams(*grad_params)
        self._init_batch_size = init_batch_size

    # kept for callers that reach into the model
    _perf_params = property(lambda self: self._perf)
    _grad_params = property(lambda self: self._grad)

    @staticmethod
    def _batch_size(num_replicas, atomic_bsz, accum_steps):
        return num_replicas * atomic_bsz * (accum_steps + 1)

    def throughput(self, num_nodes, num_replicas, atomic_bsz, accum_steps):
        """Samples per second: ``accum_steps`` local micro-steps followed by
        one step that also synchronises (gamma-norm of compute and network
        time: they overlap partially)."""
        local = _accum_time(self._perf, atomic_bsz)
        network = _network_time(self._perf, num_nodes, num_replicas)
        last = np.exp(_log_optim_time(self._perf.gamma, local, network))
        samples = self._batch_size(num_replicas, atomic_bsz, accum_steps)
        return samples / (accum_steps * local + last)

    def efficiency(self, batch_size):
        """Statistical efficiency relative to the initial batch size: how
        much progress one sample makes at ``batch_size`` (gain / scale)."""
        sqr, var = self._grad
        scale = batch_size / self._init_batch_size
        noise = var / scale + sqr
        usable = noise > 0
        gain = np.where(usable, (var + sqr) / np.where(usable, noise, 1.0),
                        1.0)
        return gain / scale

    def evaluate(self, num_nodes, num_replicas, atomic_bsz, accum_steps):
        """Goodput = throughput x efficiency (useful samples per second)."""
        samples = self._batch_size(num_replicas, atomic_bsz, accum_steps)
        if np.any(samples < self._init_batch_size):
            raise AssertionError("batch size below the initial batch size")
        speed = self.throughput(num_nodes, num_replicas, atomic_bsz,
                                accum_steps)
        return speed * self.efficiency(samples)

    __call__ = evaluate

    def optimize(self, num_nodes, num_replicas, max_batch_size=None,
                 atomic_bsz_range=None, accumulation=False):
        """Best ``(goodput, atomic_bsz, accum_steps)`` for each
        ``(num_nodes, num_replicas)`` (scalars or broadcastable arrays),
        searching 50 geometrically spaced global batch sizes between the
        initial and the maximum batch size."""
        assert np.all(np.less_equal(1, num_nodes))
        assert np.all(np.less_equal(num_nodes, num_replicas))
        init = self._init_batch_size
        if max_batch_size is None:
            max_batch_size = init
        assert init <= max_batch_size
        lo, hi = atomic_bsz_range or (None, None)
        min_atomic = lo or 1
        max_atomic = hi or max_batch_size
        shape = np.broadcast(num_nodes, num_replicas).shape
        scalar_out = np.isscalar(num_nodes) or np.isscalar(num_replicas)
        nodes = np.broadcast_to(num_nodes, shape).reshape(-1)
        replicas = np.broadcast_to(num_replicas, shape).reshape(-1)
        # candidates: rows = batch-size samples, cols = (nodes, replicas)
        floor_bsz = np.maximum(init, min_atomic * replicas)
        batch_size = np.geomspace(floor_bsz, max_batch_size)
        local_bsz = batch_size / replicas
        if accumulation:
            # split a too-large local batch into (accum_steps+1) micro-batches;
            # a single replica growing past the initial batch size needs >= 1
            # accumulation step or it has only one sample for the statistics.
            accum_steps = np.ceil(local_bsz / max_atomic - _EPS) - 1
            lonely = np.logical_and(replicas == 1, local_bsz > init + _EPS)
            accum_steps = np.where(lonely, np.maximum(accum_steps, 1),
                                   accum_steps).astype(int)
            atomic_bsz = np.ceil(local_bsz / (accum_steps + 1)
                                 - _EPS).astype(int)
        else:
            accum_steps = np.zeros_like(local_bsz, dtype=int)
            atomic_bsz = np.where(replicas == 1, init,
                                  np.ceil(local_bsz - _EPS)).astype(int)
        atomic_bsz = np.clip(atomic_bsz, min_atomic, max_atomic)
        goodput = self.evaluate(nodes, replicas, atomic_bsz, accum_steps)
        best = np.argmax(goodput, axis=0), np.arange(goodput.shape[1])
        goodput = goodput[best].reshape(shape)
        atomic_bsz = atomic_bsz[best].reshape(shape)
        accum_steps = accum_steps[best].reshape(shape)
        if scalar_out:
            return goodput.item(), atomic_bsz.item(), accum_steps.item()
        return goodput, atomic_bsz, accum_steps


# --------------------------------------------------------------------------
# Fitting
# --------------------------------------------------------------------------

def _rms(x):
    return np.sqrt(np.mean(np.square(x)))


def _objective(params, num_nodes, num_replicas, atomic_bsz,
               accum_step_time, optim_step_time, want_grad=True):
    """RMSLE(accum) + RMSLE(optim) + regularisers, and its gradient."""
    p = PerfParams(*params)
    multi_node = num_nodes > 1
    multi_repl = np.logical_and(~multi_node, num_replicas > 1)
    span = np.maximum(num_replicas - 2, _EPS)

    pa = p.alpha_c + p.beta_c * atomic_bsz
    pn = np.where(multi_node, p.alpha_n + p.beta_n * span,
                  np.where(multi_repl, p.alpha_r + p.beta_r * span,
                           _EPS + _EPS * span))
    log_pa, log_pn = np.log(pa), np.log(pn)
    la, ln = p.gamma * log_pa, p.gamma * log_pn
    log_s = np.logaddexp(la, ln)
    plo = log_s / p.gamma

    d1 = log_pa - np.log(accum_step_time)
    d2 = plo - np.log(optim_step_time)
    err1, err2 = _rms(d1), _rms(d2)
    reg1 = 1e-3 * (p.gamma - 1.0) ** 2
    rn, rr = p.beta_n / p.alpha_n, p.beta_r / p.alpha_r
    reg2 = 1e-2 * (rn ** 2 + rr ** 2)
    value = err1 + err2 + reg1 + reg2
    if not want_grad:
        return value

    n = float(len(d1))
    g1 = d1 / (n * err1) if err1 > 0 else np.zeros_like(d1)   # d err1/d d1
    g2 = d2 / (n * err2) if err2 > 0 else np.zeros_like(d2)
    wa, wn = np.exp(la - log_s), np.exp(ln - log_s)
    dplo_dpa = wa / pa
    dplo_dpn = wn / pn
    dplo_dgamma = (wa * log_pa + wn * log_pn) / p.gamma \
        - log_s / p.gamma ** 2
    # chain through pa (affects err1 and err2) and pn (err2 only)
    dpa = g1 / pa + g2 * dplo_dpa
    dpn = g2 * dplo_dpn
    grad = np.zeros(7)
    grad[0] = np.sum(dpa)
    grad[1] = np.sum(dpa * atomic_bsz)
    grad[2] = np.sum(dpn * multi_node) - 2e-2 * rn ** 2 / p.alpha_n
    grad[3] = np.sum(dpn * multi_node * span) + 2e-2 * rn / p.alpha_n
    grad[4] = np.sum(dpn * multi_repl) - 2e-2 * rr ** 2 / p.alpha_r
    grad[5] = np.sum(dpn * multi_repl * span) + 2e-2 * rr / p.alpha_r
    grad[6] = np.sum(g2 * dplo_dgamma) + 2e-3 * (p.gamma - 1.0)
    return value, grad


def _obj_fn(params, num_nodes, num_replicas, atomic_bsz,
            accum_step_time, optim_step_time):
    """Value of the fit objective at ``params`` (the name the reference's own
    ``fit_test.py`` evaluates; ``goodput.py:201-233`` there)."""
    return _objective(np.asarray(params, dtype=float), np.asarray(num_nodes),
                      np.asarray(num_replicas), np.asarray(atomic_bsz),
                      np.asarray(accum_step_time),
                      np.asarray(optim_step_time), want_grad=False)


# names under which the reference keeps the two time models
_predict_accum_time = _accum_time
_predict_network_time = _network_time


def fit_perf_params(num_nodes, num_replicas, atomic_bsz,
                    accum_step_time, optim_step_time):
    """Fit :class:`PerfParams` to measured per-configuration step times.

    Arguments are equal-length arrays, one entry per profiled
    ``(num_nodes, num_replicas, atomic_bsz)`` configuration.
    Parameters that the data cannot identify are pinned (single batch size
    -> alpha_c; no multi-node / multi-replica / N>2 data -> the matching
    network terms), and without multi-node data the inter-node terms are
    kept >= 1.1x the intra-node ones. Returns ``None`` when the objective is
    not finite at the solution (degenerate measurements).
    """
    num_nodes = np.asarray(num_nodes)
    num_replicas = np.asarray(num_replicas)
    atomic_bsz = np.asarray(atomic_bsz, dtype=float)
    accum_step_time = np.asarray(accum_step_time, dtype=float)
    optim_step_time = np.asarray(optim_step_time, dtype=float)

    x0 = [1e-1, 1e-2] * 3 + [1.0 + 1e-3]
    lower = [1e-8, 1e-8] * 3 + [1.0]
    upper = [np.inf, np.inf] * 3 + [10.0]

    def pin(i, value=None):
        x0[i] = lower[i] = upper[i] = lower[i] if value is None else value

    if len(np.unique(atomic_bsz)) == 1:
        pin(0, float(np.mean(accum_step_time)) / 2)
    if not np.any(num_nodes > 1):
        pin(2), pin(3)
    if not np.any(np.logical_and(num_nodes == 1, num_replicas > 1)):
        pin(4), pin(5)
    if not np.any(num_replicas > 2):
        pin(3), pin(5)

    args = (num_nodes, num_replicas, atomic_bsz,
            accum_step_time, optim_step_time)
    result = scipy.optimize.minimize(
        _objective, x0, args=args, jac=True, method="L-BFGS-B",
        bounds=scipy.optimize.Bounds(lower, upper, keep_feasible=True))
    params = np.array(result.x, dtype=float)
    if not (np.all(np.isfinite(params)) and np.isfinite(result.fun)):
        # L-BFGS silently hands back its starting point when the objective
        # is not finite there: that is not a fit
        return None
    if not np.any(num_nodes > 1):
        params[2] = max(params[2], params[4] * 1.1)
        params[3] = max(params[3], params[5] * 1.1)
    return PerfParams(*(float(v) for v in params))
