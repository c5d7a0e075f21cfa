"""Learning-rate scaling rules for adaptive batch sizes.

When the data loader grows the global batch by a factor ``scale`` over the
batch size the user tuned their learning rate for, a *scaling rule* decides
how much each parameter group's learning rate grows with it, applies that
factor around every ``optimizer.step()`` and advances the job's
scale-invariant *progress* by the step's gain (capabilities of the
reference's ``torch/scaling_rules.py:29-192``; the API -- class names,
``scale_lr``, ``initialize``, ``step``, ``zero_grad`` -- is the same):

===============  ==========================================================
``AdaScale``     per group ``(var + sqr) / (var / scale + sqr)`` from the
                 gradient-noise-scale running averages (SGD)
``AdamScale``    the AdaScale factor to the power 0.5 (Adam / AdamW /
                 RMSProp)
``LinearScale``  ``scale``
``SqrtScale``    ``sqrt(scale)``
``LEGWScale``    ``sqrt(scale)`` reached through a linear warm-up that is
                 measured in progress, so it survives rescaling
===============  ==========================================================

With the device engine (``parallel/engine.py``) active, ``step()`` is a
single fused optimizer launch per gradient arena: the factors were already
computed on the GPU by the statistics kernel.
"""

import math
import types
import warnings

import numpy as np

from adaptdl_b200.torch.data import current_dataloader

__all__ = ["ScalingRuleBase", "AdaScale", "AdamScale", "LinearScale",
           "SqrtScale", "LEGWScale"]


def _mark_first_step():
    """Rescale trace: the job is training again (first optimizer step of
    this process)."""
    global _mark_first_step
    from adaptdl_b200.utils import rescale_trace
    if rescale_trace.enabled():
        import torch
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        rescale_trace.mark("first_step_done")
    _mark_first_step = lambda: None      # noqa: E731 - one-shot


class _HookedOptimizerMethods(object):
    """Re-routes ``optimizer.step`` and ``optimizer.zero_grad`` of ONE
    optimizer instance to a scaling rule while keeping the originals
    reachable (the rule calls the real ``step`` under the scaled learning
    rates). Signature-compatible callables, so LR schedulers and user code
    that introspect ``optimizer.step`` keep working."""

    def __init__(self, optimizer):
        self.optimizer = optimizer
        self.real_step = optimizer.step
        self.real_zero_grad = optimizer.zero_grad

    def install(self, rule):
        # bound methods (torch's LR schedulers re-wrap ``optimizer.step`` and
        # expect ``step.__func__``)
        def step(_optimizer, *args, **kwargs):
            # what torch's LR-scheduler wrapper records on every call (its
            # "scheduler.step() before optimizer.step()" check); the fused
            # device step does not go through that wrapper
            _optimizer._opt_called = True
            return rule.step(*args, **kwargs)

        def zero_grad(_optimizer, *args, **kwargs):
            return rule.zero_grad(*args, **kwargs)
        for mine, real in ((step, self.real_step),
                           (zero_grad, self.real_zero_grad)):
            mine.__name__ = getattr(real, "__name__", mine.__name__)
            mine.__doc__ = getattr(real, "__doc__", None)
            mine.__wrapped__ = real
        # an LR scheduler built before the wrapper marked the step it saw;
        # carry the mark over, or every scheduler.step() warns that
        # optimizer.step() "has been overridden after initialization"
        if hasattr(self.real_step, "_wrapped_by_lr_sched"):
            step._wrapped_by_lr_sched = True
        self.optimizer.step = types.MethodType(step, self.optimizer)
        self.optimizer.zero_grad = types.MethodType(zero_grad,
                                                    self.optimizer)


class ScalingRuleBase(object):
    """Common machinery of the rules; subclasses provide :meth:`scale_lr`.

    A rule is normally handed to :class:`AdaptiveDataParallel`
    (``scaling_rule=AdaScale()``), which calls :meth:`initialize`; after
    that the user's ordinary ``optimizer.zero_grad()`` / ``optimizer.step()``
    calls go through the rule."""

    def __init__(self):
        self.adp = None
        self._optimizer = None
        self._hooks = None

    # -- to be provided by the rule ------------------------------------------

    def scale_lr(self, scale):
        """Learning-rate multiplier at batch-size ``scale``: a scalar, or an
        array with one entry per optimizer param group."""
        raise NotImplementedError

    # -- wiring --------------------------------------------------------------

    def initialize(self, adp, optimizer, patch_optimizer=False):
        """Bind to a data-parallel wrapper and its optimizer; with
        ``patch_optimizer`` the optimizer's own ``step`` / ``zero_grad``
        are redirected here."""
        self.adp = adp
        self._optimizer = optimizer
        self._hooks = _HookedOptimizerMethods(optimizer)
        if patch_optimizer:
            self._hooks.install(self)

    @property
    def _orig_optimizer_step(self):          # kept for API compatibility
        return self._hooks.real_step if self._hooks else None

    # -- what optimizer.zero_grad() / optimizer.step() become -----------------

    def zero_grad(self, *args, **kwargs):
        gns = self.adp.gns
        if not gns.should_zero_grad:
            # mid-accumulation: the micro-batch gradients must survive
            warnings.warn("zero_grad() ignored between gradient-accumulation "
                          "micro-steps")
            return
        gns.reset_accumulation(*args, **kwargs)

    def step(self, *args, **kwargs):
        """Apply one update with every group's learning rate multiplied by
        its factor (restored afterwards), then credit the step's gain to the
        progress counter. Does nothing on accumulation micro-steps."""
        adp = self.adp
        if adp is None:
            raise ValueError("scaling rule used before initialize(): no "
                             "AdaptiveDataParallel attached")
        if not adp.require_backward_grad_sync:
            return None
        _mark_first_step()
        engine = vars(adp).get("_engine")
        if engine is not None and engine.enabled and not (args or kwargs):
            # factors and progress were produced on the device; one fused
            # launch per gradient arena, no host synchronisation
            engine.optimizer_step()
            return None
        gns = adp.gns
        scale = gns.accum_scale * gns.accum_count
        groups = self._optimizer.param_groups
        base = [group["lr"] for group in groups]
        factors = np.broadcast_to(
            np.asarray(self.scale_lr(scale), dtype=float), (len(groups),))
        try:
            for group, lr, factor in zip(groups, base, factors):
                group["lr"] = float(lr * factor)
            outcome = self._hooks.real_step(*args, **kwargs)
        finally:
            for group, lr in zip(groups, base):
                group["lr"] = lr
        gns.set_progress(gns.get_progress() + gns.gain(scale))
        return outcome


class AdaScale(ScalingRuleBase):
    """Gain-ratio scaling (AdaScale SGD, Johnson et al. 2020): with gradient
    variance ``var`` and squared norm ``sqr`` per param group, a batch
    ``scale`` times larger is worth ``(var + sqr) / (var / scale + sqr)``
    times the learning rate -- ``scale`` while noise dominates, 1 once the
    gradient is essentially exact."""

    @staticmethod
    def _gain_ratio(var, sqr, scale):
        var = np.maximum(var, 1e-6)
        sqr = np.maximum(sqr, 0.0)
        return (var + sqr) / (var / scale + sqr)

    def scale_lr(self, scale):
        gns = self.adp.gns
        return self._gain_ratio(gns.raw_var_avg, gns.raw_sqr_avg, scale)


class AdamScale(AdaScale):
    """AdaScale for adaptive optimizers (Adam, AdamW, RMSProp): the gain
    ratio raised to ``power`` (0.5 by default)."""

    def scale_lr(self, scale, power=0.5):
        return np.power(AdaScale.scale_lr(self, scale), power)


class LinearScale(ScalingRuleBase):
    """Learning rate proportional to the batch size."""

    def scale_lr(self, scale):
        return scale


class SqrtScale(ScalingRuleBase):
    """Learning rate proportional to the square root of the batch size."""

    def scale_lr(self, scale):
        return math.sqrt(scale)


class LEGWScale(ScalingRuleBase):
    """Square-root scaling approached through a linear warm-up whose length
    grows with the batch size (LEGW, You et al. 2019). The warm-up position
    is the job's scale-invariant progress, not a step count, so a restart at
    another batch size continues the same ramp.

    Arguments:
        base_warmup_epochs: warm-up length, in epochs, at the initial batch
            size.
        data_size: samples per epoch.
    """

    def __init__(self, base_warmup_epochs, data_size):
        super().__init__()
        self._base_warmup_epochs = base_warmup_epochs
        self._data_size = data_size

    def _warmup_steps(self, scale):
        per_epoch = self._data_size / current_dataloader().batch_size
        return self._base_warmup_epochs * scale * per_epoch

    def scale_lr(self, scale):
        target = math.sqrt(scale)
        ramp = self._warmup_steps(scale)
        done = self.adp.gns.get_progress()
        return target * min(done / ramp, 1.0) if ramp > 0 else target
