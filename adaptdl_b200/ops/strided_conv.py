"""Stride-2 3x3 convolution whose data gradient is computed phase by phase.

Measured on B200 (``profiles/r1_bn/step_profile_resnet_bf16.log``): cuDNN's
``implicit_gemm_strided_dgrad`` kernels take 98 us for each of the stride-2
3x3 convolutions that open ResNet-18's stages 2 and 3 -- 4.8 GFLOP apiece,
i.e. ~5 us at the tensor-core rate, and 10 % of the whole 2 ms training step.
A strided data gradient is four independent STRIDE-1 problems, one per parity
``(alpha, beta)`` of the input pixel ``(2a + alpha, 2b + beta)``:

    dx[.., 2a+alpha, 2b+beta] = sum_{u, v} dy[.., a+u, b+v] * w[.., R_alpha[u], S_beta[v]]
    R_0 = [1]      (one tap)          R_1 = [2, 0]   (two taps)

so the 9 filter taps split 1 + 2 + 2 + 4 over the phases and no multiply is
wasted on the zeros a dilated formulation inserts. Each phase is an ordinary
(fast-path) forward convolution of ``dy`` with a 1x1 / 1x2 / 2x1 / 2x2 slice
of the transposed filter; the four results interleave into ``dx``.

Opt-in (``ADAPTDL_B200_PHASE_DGRAD=1``): the numerics are tested against
autograd on CPU, the timing has not been taken on hardware yet (round 2). The
forward pass and the weight gradient stay on cuDNN's kernels.
"""

import os

import torch
import torch.nn.functional as F

_TAP_COUNT = (1, 2)              # filter rows (columns) seen by parity 0 / 1


def _taps(t, dim, parity):
    """Filter rows (``dim`` 2) or columns (``dim`` 3) of parity ``parity``:
    ``[1]`` or ``[2, 0]``. Slices and ``flip`` only -- index lists would
    need a host-to-device copy, which CUDA-graph capture forbids."""
    if parity == 0:
        return t.narrow(dim, 1, 1)
    return t[(slice(None),) * dim + (slice(0, 3, 2),)].flip(dim)


def enabled():
    return os.environ.get("ADAPTDL_B200_PHASE_DGRAD", "0") == "1"


def supported(x, conv):
    return (conv.kernel_size == (3, 3) and conv.stride == (2, 2)
            and conv.padding == (1, 1) and conv.dilation == (1, 1)
            and conv.groups == 1 and conv.bias is None
            and conv.padding_mode == "zeros"
            and x.dim() == 4 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0)


def phase_dgrad(dy, weight, input_shape):
    """Data gradient of ``conv2d(x, weight, stride=2, padding=1)`` for a 3x3
    ``weight`` [K, C, 3, 3], ``dy`` [N, K, H/2, W/2] -> [N, C, H, W]."""
    n, c, h, w = input_shape
    channels_last = dy.is_contiguous(memory_format=torch.channels_last) \
        and not dy.is_contiguous()
    dx = torch.empty(input_shape, dtype=dy.dtype, device=dy.device,
                     memory_format=torch.channels_last if channels_last
                     else torch.contiguous_format)
    flipped = weight.transpose(0, 1)                    # [C, K, 3, 3]
    for alpha, rows in enumerate(_TAP_COUNT):
        row_taps = _taps(flipped, 2, alpha)
        for beta, cols in enumerate(_TAP_COUNT):
            taps = _taps(row_taps, 3, beta)
            taps = taps.contiguous(
                memory_format=torch.channels_last if channels_last
                else torch.contiguous_format)
            # symmetric padding computes one extra leading row / column for
            # the two-tap phases; it is sliced away (cheaper than a padded
            # copy of dy, and the interleave below copies anyway)
            full = F.conv2d(dy, taps, padding=(rows - 1, cols - 1))
            dx[:, :, alpha::2, beta::2] = full[:, :, rows - 1:, cols - 1:]
    return dx


class _PhaseDgradConv(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, weight):
        ctx.save_for_backward(x, weight)
        return F.conv2d(x, weight, None, 2, 1)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = phase_dgrad(dy, weight.to(dy.dtype), x.shape)
        if ctx.needs_input_grad[1]:
            # under autocast the forward convolution ran in the low-precision
            # dtype of dy: hand cuDNN matching operands, return the
            # parameter's dtype
            _, dw, _ = torch.ops.aten.convolution_backward(
                dy, x.to(dy.dtype), weight.to(dy.dtype), None,
                (2, 2), (1, 1), (1, 1), False, (0, 0), 1,
                (False, True, False))
            dw = dw.to(weight.dtype)
        return dx, dw


def strided_conv3x3(x, conv):
    """``conv(x)`` for a stride-2 3x3 ``nn.Conv2d``; with the flag set (and
    a supported configuration) its backward uses :func:`phase_dgrad`."""
    if not (enabled() and supported(x, conv)):
        return conv(x)
    weight = conv.weight
    if torch.is_autocast_enabled(x.device.type):
        dtype = torch.get_autocast_dtype(x.device.type)
        x, weight = x.to(dtype), weight.to(dtype)
    return _PhaseDgradConv.apply(x, weight)
