"""Learning-rate scaling rules for adaptive batch sizes.

A rule maps the batch-size *scale* (current global batch / initial batch) to
a per-param-group learning-rate factor and wraps ``optimizer.step`` so the
factor is applied transparently; it also advances the scale-invariant
*progress* counter by the AdaScale gain after each update (parity: reference
``torch/scaling_rules.py:29-192``).

* :class:`AdaScale` -- ``(var + sqr) / (var/scale + sqr)`` per group
* :class:`AdamScale` -- AdaScale ** 0.5 (Adam, AdamW, RMSProp)
* :class:`LinearScale`, :class:`SqrtScale`
* :class:`LEGWScale` -- sqrt scaling with a progress-based linear warm-up
"""

import functools
import math
import warnings
from types import MethodType

import numpy as np

from adaptdl_b200.torch.data import current_dataloader

__all__ = ["ScalingRuleBase", "AdaScale", "AdamScale", "LinearScale",
           "SqrtScale", "LEGWScale"]


class ScalingRuleBase(object):
    """Base class. Typical use is implicit, through
    :class:`adaptdl_b200.torch.AdaptiveDataParallel`; stand-alone::

        optim = torch.optim.SGD(model.parameters(), lr=0.001)
        rule = AdaScale()
        model = AdaptiveDataParallel(model, optim, scaling_rule=rule)
        for batch in loader:
            optim.zero_grad(); loss(model(batch)).backward(); optim.step()
    """

    def __init__(self):
        self.adp = None
        self._optimizer = None
        self._orig_optimizer_step = None

    def scale_lr(self, scale):
        """LR factor(s) for batch-size scale ``scale`` (scalar or one value
        per param group)."""
        raise NotImplementedError

    def zero_grad(self, *args, **kwargs):
        if self.adp.gns.should_zero_grad:
            self.adp.gns.reset_accumulation(*args, **kwargs)
        else:
            warnings.warn("skipping zero_grad for accumulated gradient")

    def step(self, *args, **kwargs):
        """One optimizer step under the scaled learning rate; a no-op on
        gradient-accumulation micro-steps."""
        if not self.adp:
            raise ValueError("AdaptiveDataParallel instance is not set!")
        if not self.adp.require_backward_grad_sync:
            return None
        engine = getattr(self.adp, "__dict__", {}).get("_engine")
        if engine is not None and engine.enabled and not args \
                and not kwargs:
            # device-resident path: the LR factors and the progress counter
            # were produced on the GPU by the statistics kernel; one fused
            # launch per gradient arena applies the update. No host sync.
            engine.optimizer_step()
            return None
        gns = self.adp.gns
        scale = gns.accum_scale * gns.accum_count
        groups = self._optimizer.param_groups
        initial_lr = [pg["lr"] for pg in groups]
        scaled_lr = np.multiply(self.scale_lr(scale), initial_lr)
        for lr, pg in zip(np.broadcast_to(scaled_lr, (len(groups),)), groups):
            pg["lr"] = float(lr)
        try:
            result = self._orig_optimizer_step(*args, **kwargs)
        finally:
            for lr, pg in zip(initial_lr, groups):
                pg["lr"] = lr
        gns.set_progress(gns.get_progress() + gns.gain(scale))
        return result

    def _patch_optimizer(self):
        """Route ``optimizer.step`` / ``optimizer.zero_grad`` through this
        rule."""
        @functools.wraps(self._optimizer.step)
        def step_wrapper(optim, *args, **kwargs):
            return self.step(*args, **kwargs)

        @functools.wraps(self._optimizer.zero_grad)
        def zero_wrapper(optim, *args, **kwargs):
            return self.zero_grad(*args, **kwargs)

        self._optimizer.step = MethodType(step_wrapper, self._optimizer)
        self._optimizer.zero_grad = MethodType(zero_wrapper, self._optimizer)

    def initialize(self, adp, optimizer, patch_optimizer=False):
        self.adp = adp
        self._optimizer = optimizer
        self._orig_optimizer_step = optimizer.step
        if patch_optimizer:
            self._patch_optimizer()


class AdaScale(ScalingRuleBase):
    """AdaScale (Johnson et al., ICML 2020): scale the LR by the gain ratio
    estimated from the gradient noise scale, per param group."""

    def scale_lr(self, scale):
        var = np.maximum(self.adp.gns.raw_var_avg, 1e-6)
        sqr = np.maximum(self.adp.gns.raw_sqr_avg, 0.0)
        return (var + sqr) / (var / scale + sqr)


class AdamScale(AdaScale):
    """AdaScale variant for Adam / AdamW / RMSProp: gain ** power."""

    def scale_lr(self, scale, power=0.5):
        return np.power(super().scale_lr(scale=scale), power)


class LinearScale(ScalingRuleBase):

    def scale_lr(self, scale):
        return scale


class SqrtScale(ScalingRuleBase):

    def scale_lr(self, scale):
        return math.sqrt(scale)


class LEGWScale(ScalingRuleBase):
    """Linear-Epoch Gradual Warmup (You et al. 2019), adapted to elastic
    training: ``sqrt(scale)`` with a linear warm-up measured in
    scale-invariant *progress* rather than raw steps.

    Arguments:
        base_warmup_epochs: warm-up epochs at the initial batch size.
        data_size: number of samples in the dataset.
    """

    def __init__(self, base_warmup_epochs, data_size):
        super().__init__()
        self._base_warmup_epochs = base_warmup_epochs
        self._data_size = data_size

    def scale_lr(self, scale):
        dataloader = current_dataloader()
        total_steps = (self._base_warmup_epochs * scale * self._data_size
                       / dataloader.batch_size)
        peak = math.sqrt(scale)
        progress = self.adp.gns.get_progress()
        if progress < total_steps:
            return peak * (progress / total_steps)
        return peak
