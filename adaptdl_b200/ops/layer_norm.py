"""Fused dropout + residual add + LayerNorm (``csrc/adl_ln.cu``).

``dropout_add_layer_norm(x, h, weight, bias, p, training, eps)`` computes
``LayerNorm(x + dropout(h))`` -- the tail of every transformer sub-layer
(reference workloads ``examples/BERT/model.py``, ``examples/transformer``) --
with one kernel forward and one (+ a small column reduction) backward instead
of PyTorch's seven launches. The keep-mask comes from ``Tensor.bernoulli_`` so
it follows the CUDA generator (and its CUDA-graph-safe Philox state); the
residual stream and the output use ``h``'s dtype (bf16 under autocast).

Falls back to the PyTorch composition on CPU and for widths the kernel does
not cover (D > 1024 in 16-bit, > 512 in fp32, D not a multiple of the vector
width).

On by default (``ADAPTDL_B200_FUSED_LN=0`` gives the PyTorch composition);
BERT-base step 11.26 -> 10.68 ms (``profiles/r2_validate``).
"""

import ctypes
import os

import torch
import torch.nn.functional as F

from adaptdl_b200.ops import _count

_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
_SM = {}


def supported(h):
    if not h.is_cuda or h.dtype not in _DTYPES or \
            os.environ.get("ADAPTDL_B200_FUSED_LN", "1") == "0":
        return False
    d = h.shape[-1]
    vec = 4 if h.dtype == torch.float32 else 8
    return d % vec == 0 and d // vec <= 128 and h.numel() > 0


def _grid(device, m, per_sm=2):
    """CTAs (8 warps = 8 rows at a time). The forward is pure load latency
    per row, so it takes one row per warp when the machine can hold them all
    (4 CTAs/SM); the backward carries per-CTA partial sums, so fewer, longer
    CTAs (2/SM = its register-limited residency)."""
    n = _SM.get(device.index)
    if n is None:
        n = torch.cuda.get_device_properties(device).multi_processor_count
        _SM[device.index] = n
    return max(1, min((m + 7) // 8, per_sm * n))


def _launch(args, dtype, backward, grid, device):
    from adaptdl_b200 import _native
    lib = _native.load()
    lib.adl_set_device(device.index)
    code = lib.adl_dropout_add_ln(
        ctypes.byref(args), _DTYPES[dtype], backward, grid,
        torch.cuda.current_stream(device).cuda_stream)
    if code < 0:
        raise RuntimeError("adl_dropout_add_ln rejected the call (code {})"
                           .format(code))
    _native.check(code, "adl_dropout_add_ln")
    _count.add(2 if backward else 1)


class _DropoutAddLN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h, weight, bias, mask, scale, eps):
        from adaptdl_b200._native import LnArgs
        d = h.shape[-1]
        x2 = x.reshape(-1, d).contiguous()
        h2 = h.reshape(-1, d).contiguous()
        m = h2.shape[0]
        dev = h.device
        y = torch.empty_like(h2)
        z = torch.empty_like(h2)
        mean = torch.empty(m, dtype=torch.float32, device=dev)
        rstd = torch.empty(m, dtype=torch.float32, device=dev)
        gamma = weight.float().contiguous()
        beta = bias.float().contiguous()
        a = LnArgs()
        a.x, a.h, a.y, a.z = x2.data_ptr(), h2.data_ptr(), y.data_ptr(), \
            z.data_ptr()
        a.mask = mask.data_ptr() if mask is not None else None
        a.gamma, a.beta = gamma.data_ptr(), beta.data_ptr()
        a.mean, a.rstd = mean.data_ptr(), rstd.data_ptr()
        a.M, a.D, a.scale, a.eps = m, d, scale, eps
        _launch(a, h2.dtype, 0, _grid(dev, m, 4), dev)
        ctx.save_for_backward(z, mask, gamma, mean, rstd)
        ctx.scale = scale
        ctx.shape = h.shape
        return y.view(h.shape)

    @staticmethod
    def backward(ctx, dy):
        from adaptdl_b200._native import LnArgs
        z, mask, gamma, mean, rstd = ctx.saved_tensors
        m, d = z.shape
        dev = z.device
        dy2 = dy.reshape(m, d).to(z.dtype).contiguous()
        dx = torch.empty_like(z)
        dh = torch.empty_like(z)
        grid = _grid(dev, m)
        partial = torch.empty(grid * 2 * d, dtype=torch.float32, device=dev)
        dgamma = torch.empty(d, dtype=torch.float32, device=dev)
        dbeta = torch.empty(d, dtype=torch.float32, device=dev)
        a = LnArgs()
        a.h, a.z, a.y, a.dh = dy2.data_ptr(), z.data_ptr(), dx.data_ptr(), \
            dh.data_ptr()
        a.mask = mask.data_ptr() if mask is not None else None
        a.gamma, a.mean, a.rstd = gamma.data_ptr(), mean.data_ptr(), \
            rstd.data_ptr()
        a.partial, a.dgamma, a.dbeta = partial.data_ptr(), \
            dgamma.data_ptr(), dbeta.data_ptr()
        a.M, a.D, a.n_partial, a.scale = m, d, grid, ctx.scale
        _launch(a, z.dtype, 1, grid, dev)
        return (dx.view(ctx.shape), dh.view(ctx.shape), dgamma, dbeta, None,
                None, None)


def dropout_add_layer_norm(x, h, weight, bias, p=0.0, training=True,
                           eps=1e-5, mask=None):
    """``layer_norm(x + dropout(h, p, training), weight, bias, eps)``.
    ``mask`` (uint8, 1 = keep) overrides the random keep-mask (tests)."""
    drop = training and p > 0.0
    if supported(h) and x.shape == h.shape:
        if drop and mask is None:
            mask = torch.empty(h.shape, dtype=torch.uint8,
                               device=h.device).bernoulli_(1.0 - p)
        if not drop:
            mask = None
        return _DropoutAddLN.apply(x.to(h.dtype), h, weight, bias, mask,
                                   1.0 / (1.0 - p) if drop else 1.0, eps)
    if mask is not None and drop:
        h = h * mask.to(h.dtype) / (1.0 - p)
    else:
        h = F.dropout(h, p, training)
    return F.layer_norm(x + h, (h.shape[-1],), weight, bias, eps)
