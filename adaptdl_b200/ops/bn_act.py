"""Fused BatchNorm (+ residual add) (+ ReLU), training mode, channels-last
(``csrc/adl_bn.cu``).

The ResNet / CIFAR-zoo workloads (reference ``examples/pytorch-cifar``) are
dominated by normalisation and elementwise passes: stock PyTorch runs
``bn -> (+ shortcut) -> relu`` as three kernels forward and two or three
backward, each a full trip over the activation. ``bn_act`` does the block
with a per-channel reduction, a C-element finalize and ONE elementwise pass
per direction (the second read of the activation comes out of L2).

``BatchNormAct2d`` is a drop-in ``nn.BatchNorm2d`` (same parameters, buffers
and state dict) whose ``forward(x, residual=None, relu=True)`` uses the fused
kernels for channels-last CUDA tensors in training mode and the PyTorch
composition everywhere else (CPU, eval mode, NCHW-contiguous inputs, odd
channel counts).
"""

import ctypes
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from adaptdl_b200.ops import _count

_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
_SM = {}


def _sm_count(device):
    n = _SM.get(device.index)
    if n is None:
        n = torch.cuda.get_device_properties(device).multi_processor_count
        _SM[device.index] = n
    return n


def supported(x):
    """Can ``x`` ([N, C, H, W] channels-last or [M, C]) take the fused path?"""
    if not x.is_cuda or x.dtype not in _DTYPES or \
            os.environ.get("ADAPTDL_B200_FUSED_BN", "1") == "0":
        return False
    if x.dim() == 4:
        if not x.is_contiguous(memory_format=torch.channels_last):
            return False
    elif x.dim() != 2 or not x.is_contiguous():
        return False
    c = x.shape[1]
    vec = 4 if x.dtype == torch.float32 else 8
    cb = min(c, 64)
    if c % vec or c % cb or cb % vec:
        return False
    tpr = c // vec
    return tpr <= 256 and 256 % tpr == 0 and 256 % (cb // vec) == 0 and \
        x.numel() > 0


def _grid(device, m, c, vec, unroll, waves=8):
    """CTAs of the elementwise pass (a row holds C / vec threads)."""
    rpi = 256 // (c // vec)
    need = (m + rpi * unroll - 1) // (rpi * unroll)
    return max(1, min(need, waves * _sm_count(device)))


def _reduce_grid(device, m, c, vec, unroll):
    """(channels per CTA, row chunks) of the reduction: ~2 CTAs per SM."""
    cb = min(c, 64)
    rpi = 256 // (cb // vec)
    need = (m + rpi * unroll - 1) // (rpi * unroll)
    return cb, max(1, min(need, 2 * _sm_count(device) // (c // cb)))


_TICKETS = {}


def _tickets(device, n):
    """``n`` zeroed int32 ticket counters. They live in a per-device ring:
    the kernels leave them at zero, and consecutive calls take different
    slots, so launches that overlap on different streams only share a
    counter if they are a full trip around the ring (>= 512 calls) apart;
    launches on one stream are ordered anyway."""
    state = _TICKETS.get(device.index)
    if state is None:
        state = [torch.zeros(4096, dtype=torch.int32, device=device), 0]
        _TICKETS[device.index] = state
    if state[1] + n > state[0].numel():
        state[1] = 0
    ptr = state[0].data_ptr() + 4 * state[1]
    state[1] += n
    return ptr


def _launch(args, dtype, backward, grid, grid_apply, device):
    from adaptdl_b200 import _native
    lib = _native.load()
    lib.adl_set_device(device.index)
    code = lib.adl_bn_act(ctypes.byref(args), _DTYPES[dtype], backward, grid,
                          grid_apply,
                          torch.cuda.current_stream(device).cuda_stream)
    if code < 0:
        raise RuntimeError("adl_bn_act rejected the call (code {})".format(
            code))
    _native.check(code, "adl_bn_act")
    _count.add(2)                     # reduce(+finalize) and apply


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _like(x, ref):
    """``x`` laid out like ``ref`` (channels-last for 4-D)."""
    if ref.dim() == 4:
        return x.contiguous(memory_format=torch.channels_last)
    return x.contiguous()


class _BnAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, residual,
                relu, momentum, eps, num_batches_tracked=None):
        from adaptdl_b200._native import BnArgs
        dev, c = x.device, x.shape[1]
        m = x.numel() // c
        vec = 4 if x.dtype == torch.float32 else 8
        if residual is not None:
            residual = _like(residual.to(x.dtype), x)
        y = torch.empty_like(x)           # preserve_format: channels-last
        mean = torch.empty(c, dtype=torch.float32, device=dev)
        rstd = torch.empty(c, dtype=torch.float32, device=dev)
        cb, grid = _reduce_grid(dev, m, c, vec, 6)
        scratch = torch.empty((grid + 1) * 2 * c, dtype=torch.float32,
                              device=dev)
        gamma = weight.float() if weight is not None else \
            torch.ones(c, dtype=torch.float32, device=dev)
        beta = bias.float() if bias is not None else \
            torch.zeros(c, dtype=torch.float32, device=dev)
        a = BnArgs()
        a.x, a.res, a.y = x.data_ptr(), _ptr(residual), y.data_ptr()
        a.gamma, a.beta = gamma.data_ptr(), beta.data_ptr()
        a.mean, a.rstd = mean.data_ptr(), rstd.data_ptr()
        a.running_mean, a.running_var = _ptr(running_mean), _ptr(running_var)
        a.num_batches_tracked = _ptr(num_batches_tracked)
        a.partial = scratch.data_ptr()
        a.coef = scratch.data_ptr() + grid * 2 * c * 4
        a.M, a.C, a.n_partial, a.relu = m, c, grid, int(relu)
        a.cb, a.counters = cb, _tickets(dev, c // cb)
        a.eps, a.momentum = eps, momentum
        _launch(a, x.dtype, 0, grid, _grid(dev, m, c, vec, 6), dev)
        ctx.save_for_backward(x, y if relu else None, gamma, mean, rstd)
        ctx.relu = bool(relu)
        ctx.has_res = residual is not None
        ctx.has_affine = (weight is not None, bias is not None)
        ctx.mark_non_differentiable(*[
            t for t in (running_mean, running_var, num_batches_tracked)
            if t is not None])
        return y

    @staticmethod
    def backward(ctx, dy):
        from adaptdl_b200._native import BnArgs
        x, y, gamma, mean, rstd = ctx.saved_tensors
        dev, c = x.device, x.shape[1]
        m = x.numel() // c
        vec = 4 if x.dtype == torch.float32 else 8
        dy = _like(dy.to(x.dtype), x)
        dx = torch.empty_like(x)
        # without an activation the residual's gradient is dy itself
        dres = torch.empty_like(x) if (ctx.has_res and ctx.relu) else None
        dgamma = torch.empty(c, dtype=torch.float32, device=dev)
        dbeta = torch.empty(c, dtype=torch.float32, device=dev)
        cb, grid = _reduce_grid(dev, m, c, vec, 3)
        scratch = torch.empty((grid + 1) * 2 * c, dtype=torch.float32,
                              device=dev)
        a = BnArgs()
        a.x, a.y, a.dy = x.data_ptr(), _ptr(y), dy.data_ptr()
        a.dx, a.dres = dx.data_ptr(), _ptr(dres)
        a.gamma, a.mean, a.rstd = gamma.data_ptr(), mean.data_ptr(), \
            rstd.data_ptr()
        a.dgamma, a.dbeta = dgamma.data_ptr(), dbeta.data_ptr()
        a.partial = scratch.data_ptr()
        a.coef = scratch.data_ptr() + grid * 2 * c * 4
        a.M, a.C, a.n_partial, a.relu = m, c, grid, int(ctx.relu)
        a.cb, a.counters = cb, _tickets(dev, c // cb)
        _launch(a, x.dtype, 1, grid, _grid(dev, m, c, vec, 3), dev)
        if ctx.has_res and not ctx.relu:
            dres = dy
        return (dx, dgamma if ctx.has_affine[0] else None,
                dbeta if ctx.has_affine[1] else None, None, None, dres,
                None, None, None, None)


def bn_act(x, weight, bias, running_mean=None, running_var=None,
           residual=None, relu=True, training=True, momentum=0.1, eps=1e-5,
           num_batches_tracked=None):
    """``act(batch_norm(x) + residual)`` (``act`` = ReLU or identity).
    ``num_batches_tracked`` (int64 scalar tensor) is incremented if given."""
    if training and supported(x) and \
            (residual is None or residual.shape == x.shape):
        return _BnAct.apply(x, weight, bias, running_mean, running_var,
                            residual, relu, momentum, eps,
                            num_batches_tracked)
    if num_batches_tracked is not None:
        num_batches_tracked.add_(1)
    out = F.batch_norm(x, running_mean, running_var, weight, bias, training,
                       momentum, eps)
    if residual is not None:
        out = out + residual
    return F.relu(out) if relu else out


class BatchNormAct2d(nn.BatchNorm2d):
    """``nn.BatchNorm2d`` + optional residual add + optional ReLU in one op.
    State-dict compatible with ``nn.BatchNorm2d``."""

    def forward(self, x, residual=None, relu=True):
        tracking = self.training and self.track_running_stats
        if tracking and self.momentum is None:   # cumulative average: not fused
            out = super().forward(x)
            if residual is not None:
                out = out + residual
            return F.relu(out) if relu else out
        use_batch = self.training or self.running_mean is None
        return bn_act(
            x, self.weight, self.bias,
            self.running_mean if self.track_running_stats else None,
            self.running_var if self.track_running_stats else None,
            residual, relu, use_batch,
            self.momentum if self.momentum is not None else 0.1, self.eps,
            self.num_batches_tracked if tracking else None)
