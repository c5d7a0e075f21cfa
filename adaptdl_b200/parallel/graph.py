"""CUDA-graph training step.

``GraphedTrainStep`` captures one *whole* optimizer step -- gradient zeroing,
forward, backward (with the fused bucket all-reduce + statistics launched from
the autograd hooks on the communication stream), the on-device noise-scale
estimator, and the fused optimizer -- into a CUDA graph and replays it.
This is possible because, with the device engine (``parallel/engine.py``),
nothing on the step path needs the host: flag epochs come from a device
counter, learning-rate factors are read from device memory, statistics reach
the host through a pinned mailbox.

Small models are launch-bound on a B200 (ResNet-18/CIFAR at batch 128 is
~1000 kernel launches per step); replaying a graph removes the per-launch CPU
cost and lets the CPU run the data loader concurrently.

One graph is kept per distinct ``(sync, zero, micro-step index, input
shapes)`` so gradient accumulation and adaptive batch sizes work: a new
configuration runs eagerly for ``warmup`` steps (cuDNN autotuning, lazy
buffers), then is captured.
"""

import logging

import torch

from adaptdl_b200.ops import _count as _ops_count
from adaptdl_b200.utils.trace import traced


LOG = logging.getLogger(__name__)

__all__ = ["GraphedTrainStep"]


class _Captured(object):
    __slots__ = ("graph", "inputs", "loss", "launches", "ops_launches",
                 "n_rows")


class GraphedTrainStep(object):
    """Arguments:
        net: an :class:`adaptdl_b200.torch.AdaptiveDataParallel`.
        optimizer: the optimizer passed to ``net`` (patched by it).
        loss_fn: ``loss_fn(net, *inputs) -> scalar loss tensor``.
        autocast_dtype: e.g. ``torch.bfloat16`` (``None`` = no autocast).
        grad_scaler: the ``torch.amp.GradScaler`` given to ``net`` (fp16
            training); the device engine keeps it host-sync free, so the
            scaled step is captured like any other.
        warmup: eager steps per configuration before capturing.
        enabled: ``False`` forces the eager path (same call signature).

    ``step(*inputs)`` accepts CPU (ideally pinned) or device tensors, copies
    them into the graph's static input buffers and returns the (static,
    device) loss tensor of this step. Falls back to eager execution whenever
    the device engine is not active (CPU, AMP GradScaler, unsupported
    optimizer).
    """

    def __init__(self, net, optimizer, loss_fn, autocast_dtype=None,
                 warmup=3, enabled=True, channels_last=False,
                 grad_scaler=None):
        self.net = net
        self.optimizer = optimizer
        self.loss_fn = loss_fn
        # torch.amp.GradScaler (the one passed to ``net`` as ``mp_scaler``):
        # the step becomes scale(loss).backward(); step(optimizer); update()
        self.grad_scaler = grad_scaler
        self.autocast_dtype = autocast_dtype
        self.warmup = max(1, int(warmup))
        self.enabled = enabled
        self.channels_last = channels_last   # 4-D float inputs -> NHWC
        self._graphs = {}
        self._eager_runs = {}
        self.replays = 0
        self.eager_steps = 0

    # ------------------------------------------------------------------

    def _device(self):
        return self.net.reducer.device

    def _body(self, *inputs):
        self.optimizer.zero_grad()
        dev = self._device()
        if self.autocast_dtype is not None and dev.type == "cuda":
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                loss = self.loss_fn(self.net, *inputs)
        else:
            loss = self.loss_fn(self.net, *inputs)
        scaler = self.grad_scaler
        if scaler is not None and scaler.is_enabled():
            scaler.scale(loss).backward()
            scaler.step(self.optimizer)
            scaler.update()
        else:
            loss.backward()
            self.optimizer.step()
        return loss.detach()

    def _wants_nhwc(self, t):
        return (self.channels_last and torch.is_tensor(t) and t.dim() == 4
                and t.is_floating_point())

    def _to_device(self, inputs):
        dev = self._device()
        out = []
        for t in inputs:
            if torch.is_tensor(t):
                t = t.to(dev, non_blocking=True)
                if self._wants_nhwc(t):
                    t = t.contiguous(memory_format=torch.channels_last)
            out.append(t)
        return out

    def _can_graph(self):
        engine = getattr(self.net, "engine", None)
        return (self.enabled and engine is not None and engine.enabled
                and self._device().type == "cuda")

    def _key(self, inputs):
        net = self.net
        # imported here: adaptdl_b200.torch imports this module
        from adaptdl_b200.torch.data import current_dataloader
        dataloader = current_dataloader()
        if dataloader is not None and dataloader.training:
            sync = dataloader.is_optim_step()
        else:
            sync = net.require_backward_grad_sync
        zero = net.gns.should_zero_grad
        k_before = 0 if zero else net.reducer.accum_count
        shapes = tuple((tuple(t.shape), t.dtype) if torch.is_tensor(t)
                       else ("const", t) for t in inputs)
        return (bool(sync), bool(zero), int(k_before), shapes)

    # ------------------------------------------------------------------

    @traced("graphed_step")
    def __call__(self, *inputs):
        if not self._can_graph():
            self.eager_steps += 1
            return self._body(*self._to_device(inputs))
        key = self._key(inputs)
        cap = self._graphs.get(key)
        if cap is None:
            runs = self._eager_runs.get(key, 0)
            if runs < self.warmup:
                self._eager_runs[key] = runs + 1
                self.eager_steps += 1
                return self._body(*self._to_device(inputs))
            return self._capture(key, inputs)
        return self._replay(key, cap, inputs)

    step = __call__

    # ------------------------------------------------------------------

    def _capture(self, key, inputs):
        net, red = self.net, self.net.reducer
        dev = self._device()
        cap = _Captured()
        cap.inputs = [
            torch.empty(t.shape, dtype=t.dtype, device=dev,
                        memory_format=(torch.channels_last
                                       if self._wants_nhwc(t)
                                       else torch.contiguous_format))
            if torch.is_tensor(t) else t for t in inputs]
        for dst, src in zip(cap.inputs, inputs):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=True)
        # everything the host pushes to the device must be in place before
        # the capture starts (no pinned allocations / H2D copies inside)
        net._pre_forward()
        net.gns._flush()
        net.gns.before_captured_step(key[0], key[2])
        net.engine.sync_hyper()
        launches0 = red.launches
        ops0 = _ops_count.total()
        steps0 = red._steps
        torch.cuda.synchronize(dev)
        cap.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cap.graph, capture_error_mode="thread_local"):
            cap.loss = self._body(*cap.inputs)
        cap.launches = red.launches - launches0
        cap.ops_launches = _ops_count.total() - ops0
        cap.n_rows = None
        # the capture ran the step's host bookkeeping but no kernels: replay
        # once so the device actually performs this step
        assert red._steps - steps0 in (0, 1)
        cap.graph.replay()
        self._graphs[key] = cap
        self.replays += 1
        LOG.info("captured training-step graph %s (%d fused launches)",
                 key[:3], cap.launches)
        return cap.loss

    def _replay(self, key, cap, inputs):
        net, red = self.net, self.net.reducer
        sync, zero, k_before, _ = key
        for dst, src in zip(cap.inputs, inputs):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=True)
        # host prologue: what zero_grad()/forward() do on the host
        net.gns._flush()
        net._pre_forward()
        net.gns.before_captured_step(sync, k_before)
        net.engine.sync_hyper()
        cap.graph.replay()
        # host epilogue: what the backward hooks / optimizer.step() record
        red.replay_bookkeeping(sync, k_before)
        net.gns._after_backward(sync)
        if sync:
            net.engine._opt_steps_host += 1
        red.launches += cap.launches
        _ops_count.add(cap.ops_launches)
        self.replays += 1
        return cap.loss
