"""Gradient-noise-scale estimation and gradient-accumulation bookkeeping.

Tracks, per optimizer param group, running estimates of the squared norm of
the true gradient (``sqr_avg``) and of the trace of the per-sample-batch
gradient covariance (``var_avg``); their ratio is the gradient noise scale
that drives AdaScale learning-rate scaling and the goodput model's
statistical efficiency.

The *estimator* is the reference's (``torch/gradient_noise_scale.py:212-273``,
SURVEY App. A) and is a pure function of two per-group statistics of an
optimizer step:

* ``local_sqr`` -- the mean over replicas and micro-steps of ``|g/P|^2``
* ``total_sqr`` -- ``|mean gradient / P|^2``

What differs is where those statistics come from. The reference computes them
with ~10 tiny kernels per parameter plus a host sync per param group; here
they are by-products of the fused bucket all-reduce
(:mod:`adaptdl_b200.parallel`), read back once per step, lazily (at the next
point the host actually needs them -- normally ``optimizer.step``).
"""

import logging
import math

import numpy as np
import torch

from adaptdl_b200.parallel import make_reducer

LOG = logging.getLogger(__name__)

__all__ = ["GradientNoiseScale", "AdamGradientNoiseScale", "estimate"]

SMOOTHING = 0.999


def estimate(local_sqr, total_sqr, count, scale):
    """The replica/accumulation estimator: unbiased estimates of
    ``(|true grad|^2, tr(cov))`` from ``count > 1`` gradient samples whose
    mean squared norm is ``local_sqr`` and whose mean has squared norm
    ``total_sqr``; ``scale`` is the batch-size scale of the mean gradient."""
    grad_sqr = (count * total_sqr - local_sqr) / (count - 1)
    grad_var = (local_sqr - total_sqr) * scale / (count - 1)
    return grad_sqr, grad_var


class GradientNoiseScale(object):
    """Tracks gradient statistics and owns gradient accumulation.

    Arguments:
        adp: the data-parallel wrapper; only ``adp.require_backward_grad_sync``
            is read (is the current backward the one that synchronises).
        optimizer: its ``param_groups`` define the statistics groups and its
            ``state["gns"]`` holds the running averages (so they ride along
            in ``optimizer.state_dict()`` checkpoints).
        mp_scaler: optional ``torch.amp.GradScaler``.
        num_replicas: defaults to the torch.distributed world size.
        accum_scale: batch-size scale of one micro-batch *per step over all
            replicas* relative to the initial batch size.
        reducer: an existing :class:`~adaptdl_b200.parallel.reducer_base.
            GradReducer`; by default one is created over the optimizer's
            parameters.
    """

    def __init__(self, adp, optimizer, mp_scaler=None, num_replicas=None,
                 accum_scale=None, reducer=None, **reducer_kwargs):
        self._adp = adp
        self._optimizer = optimizer
        self._orig_optimizer_zero_grad = optimizer.zero_grad
        self._should_zero_grad = True
        self._mp_scaler = mp_scaler
        if num_replicas is None:
            num_replicas = (torch.distributed.get_world_size()
                            if torch.distributed.is_available()
                            and torch.distributed.is_initialized() else 1)
        self._num_replicas = num_replicas
        self._accum_scale = accum_scale or self._num_replicas
        self._smoothing = SMOOTHING
        self._listeners = []
        self._backward_listeners = []
        self._engine = None
        self._in_flush = False
        num_groups = len(optimizer.param_groups)
        self._optimizer.state.setdefault("gns", {
            "progress": 0.0,
            "prev_scale": 0.0,
            "sqr_avg": np.ones(num_groups),
            "var_avg": np.zeros(num_groups),
            "biased": False,
        })
        if reducer is None:
            rank = (torch.distributed.get_rank()
                    if torch.distributed.is_available()
                    and torch.distributed.is_initialized() else 0)
            reducer = make_reducer(
                optimizer.param_groups, self._num_replicas, rank,
                self._wants_sync, name="gns", **reducer_kwargs)
        self._reducer = reducer
        reducer._should_sync = self._wants_sync
        reducer._on_backward_end = self._after_backward
        self._install_preconditioner()
        self.reset_accumulation()

    # -- plumbing ----------------------------------------------------------

    def _wants_sync(self):
        return bool(self._adp.require_backward_grad_sync)

    def _install_preconditioner(self):
        pass

    @property
    def reducer(self):
        return self._reducer

    @property
    def _state(self):
        return self._optimizer.state["gns"]

    def add_listener(self, fn):
        """``fn()`` is invoked after every statistics update (used by
        AdaptiveDataParallel to publish gain / progress)."""
        self._listeners.append(fn)

    def attach_engine(self, engine):
        """Switch to the device-resident estimator: statistics are folded
        into the running averages on the GPU; this object then only mirrors
        them (with a fixed lag, see ``parallel/engine.py``)."""
        self._engine = engine

    def add_backward_listener(self, fn):
        """``fn(sync)`` is invoked at the end of every backward pass."""
        self._backward_listeners.append(fn)

    # -- accumulation ------------------------------------------------------

    def reset_accumulation(self, *args, **kwargs):
        """Zero gradients and restart gradient accumulation (this is what
        the patched ``optimizer.zero_grad`` does)."""
        self._flush()
        self._reducer.zero()

    @property
    def should_zero_grad(self):
        return self._should_zero_grad

    @property
    def accum_scale(self):
        return self._accum_scale

    @property
    def accum_count(self):
        return self._reducer.accum_count

    def set_accum_scale(self, accum_scale):
        if not np.isclose(self._accum_scale, accum_scale):
            self.reset_accumulation()
            self._accum_scale = accum_scale

    # -- estimates -----------------------------------------------------------

    @property
    def raw_sqr_avg(self):
        self._flush()
        view = np.asarray(self._state["sqr_avg"]).view()
        view.flags.writeable = False
        return view

    @property
    def raw_var_avg(self):
        self._flush()
        view = np.asarray(self._state["var_avg"]).view()
        view.flags.writeable = False
        return view

    def sqr_avg(self):
        """Estimate of the squared l2 norm of the true gradient."""
        self._flush()
        return float(np.sum(np.maximum(self._state["sqr_avg"], 0.0)))

    def var_avg(self):
        """Estimate of the trace of the gradient covariance."""
        self._flush()
        return float(np.sum(np.maximum(self._state["var_avg"], 1e-6)))

    def get_progress(self):
        return self._state["progress"]

    def set_progress(self, progress):
        self._state["progress"] = progress
        if self._engine is not None and self._engine.enabled:
            self._engine.set_progress(progress)

    def gain(self, scale):
        """AdaScale gain ratio at batch-size scale ``scale``."""
        var = self.var_avg()
        norm = self.sqr_avg()
        return (var + norm) / (var / scale + norm)

    # -- running averages ----------------------------------------------------

    def _update_avg(self, name, value, factor):
        state = self._state
        biased = state.get(name + "_biased", 0.0)
        unbias = state.get(name + "_unbias", 0.0)
        biased = factor * biased + (1.0 - factor) * value
        unbias = factor * unbias + (1.0 - factor)
        state[name + "_biased"] = biased
        state[name + "_unbias"] = unbias
        state[name] = biased / unbias

    def _reset_avg(self, name):
        self._state.pop(name + "_biased", None)
        self._state.pop(name + "_unbias", None)

    # -- per-backward callback (from the reducer) ------------------------------

    def _after_backward(self, sync):
        # Runs at the end of every backward. On a synchronising backward the
        # statistics are still in flight on the device: only bookkeeping
        # here, no host sync.
        self._should_zero_grad = bool(sync)
        if sync:
            self._before_update()
            self._pending = True
        for fn in self._backward_listeners:
            fn(sync)

    def _before_update(self):
        pass

    def before_captured_step(self, sync, k_before):
        pass

    _pending = False

    def _flush(self):
        """Fold the statistics of the last synchronised backward (if any)
        into the running averages. The one place the host waits for the
        device."""
        if self._engine is not None and self._engine.enabled:
            if self._in_flush:
                return
            self._in_flush = True
            try:
                self._pending = False
                self._reducer._stats_ready = None
                header = self._engine.mirror(self._state)
                if header is not None:
                    from adaptdl_b200.parallel.reducer_base import GradStats
                    stats = GradStats(None, None, 0, None,
                                      sync_time=float(header[4]) * 1e-9)
                    for fn in self._listeners:
                        fn(stats)
            finally:
                self._in_flush = False
            return
        if not self._pending:
            return
        self._pending = False
        stats = self._reducer.pop_stats()
        if stats is None:
            return
        self._update(stats)
        for fn in self._listeners:
            fn(stats)

    def _update(self, stats):
        mp_scale = (self._mp_scaler.get_scale()
                    if self._mp_scaler is not None else 1.0)
        total_sqr = stats.total_sqr / mp_scale ** 2
        if not np.all(np.isfinite(total_sqr)):
            LOG.warning("GradientNoiseScale detected invalid gradient at "
                        "scale %s, skipping statistics update", mp_scale)
            self._reducer.invalidate_stash()
            return
        k = stats.count // self._num_replicas
        count = stats.count
        scale = self._accum_scale * k
        state = self._state
        if count > 1:
            local_sqr = stats.local_sqr / count / mp_scale ** 2
            if state["biased"]:
                self._reset_avg("sqr_avg")
                self._reset_avg("var_avg")
            state["biased"] = False
            self._reducer.invalidate_stash()
        else:
            # a single gradient sample: difference it against the previous
            # step's gradient (biased; flagged so the averages restart once
            # real replicas/accumulation appear).
            state["biased"] = True
            if stats.pair is None:
                return
            prev_sqr, avg_sqr = stats.pair
            local_sqr = (prev_sqr / mp_scale ** 2 + total_sqr) / 2
            total_sqr = avg_sqr / mp_scale ** 2
            count = 2
            scale = 2 * self._accum_scale
        grad_sqr, grad_var = estimate(local_sqr, total_sqr, count, scale)
        theta = self._smoothing ** scale
        self._update_avg("sqr_avg", grad_sqr, theta)
        self._update_avg("var_avg", grad_var, theta)


class AdamGradientNoiseScale(GradientNoiseScale):
    """Statistics of the Adam-preconditioned gradient ``g / (sqrt(v_hat) +
    eps)`` (once Adam has taken 5 steps), for :class:`AdamScale`."""

    WARMUP_STEPS = 5

    def __init__(self, adp, optimizer, mp_scaler=None, num_replicas=None,
                 accum_scale=None, **kwargs):
        self._adam_param_group = {
            "beta": [g["betas"][1] for g in optimizer.param_groups],
            "eps": [g["eps"] for g in optimizer.param_groups],
        }
        self._group_of = {}
        for idx, group in enumerate(optimizer.param_groups):
            for p in group["params"]:
                self._group_of[id(p)] = idx
        super().__init__(adp, optimizer, mp_scaler, num_replicas,
                         accum_scale, **kwargs)

    def _install_preconditioner(self):
        self._reducer.set_preconditioner(self._calculate_preconditioner)

    def _calculate_preconditioner(self, param):
        state = self._optimizer.state.get(param, {})
        step = state.get("step", 0)
        step = float(step) if not torch.is_tensor(step) else float(step.item())
        if step < self.WARMUP_STEPS:
            return None
        idx = self._group_of[id(param)]
        beta2 = self._adam_param_group["beta"][idx]
        eps = self._adam_param_group["eps"][idx]
        correction = 1 - beta2 ** step
        return (state["exp_avg_sq"].sqrt() / math.sqrt(correction)).add_(eps)

    def _reset_adam_state(self, step=0):
        # NOTE (reference quirk, App. D3): with step=0 the factors are
        # (1-beta^0)/(1-beta^t) = 0, i.e. the moments are zeroed.
        if self._engine is not None and self._engine.enabled and step == 0:
            self._engine.reset_adam_state()      # same thing on the device
            return
        for group in self._optimizer.param_groups:
            beta1, beta2 = group["betas"]
            for param in group["params"]:
                state = self._optimizer.state.get(param, {})
                cur = state.get("step", 0)
                cur_f = float(cur.item()) if torch.is_tensor(cur) \
                    else float(cur)
                if cur_f > 0:
                    state["exp_avg"].mul_(
                        (1 - beta1 ** step) / (1 - beta1 ** cur_f))
                    state["exp_avg_sq"].mul_(
                        (1 - beta2 ** step) / (1 - beta2 ** cur_f))
                    if torch.is_tensor(cur):
                        cur.fill_(step)
                    else:
                        state["step"] = step

    def _before_update(self):
        scale = self._accum_scale * self.accum_count
        if not np.isclose(scale, self._state["prev_scale"]):
            self._reset_adam_state()
            self._state["prev_scale"] = scale

    def before_captured_step(self, sync, k_before):
        """A CUDA-graph replay runs the optimizer inside the graph, i.e.
        before the backward-end callback: do the scale-change reset ahead of
        the replay."""
        if not sync:
            return
        scale = self._accum_scale * (k_before + 1)
        if not np.isclose(scale, self._state["prev_scale"]):
            self._reset_adam_state()
            self._state["prev_scale"] = scale
