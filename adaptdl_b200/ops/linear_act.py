"""Linear + bias + GELU on the 5th-gen tensor cores (``csrc/adl_gemm.cu``).

``linear_act(x, weight, bias, "gelu")`` computes ``gelu(x @ weight.T + bias)``
with ONE persistent tcgen05/TMA kernel: bias and the exact (erf) GELU are
applied to the fp32 accumulator in tensor memory and the pre-activation the
backward pass needs is stored from the same registers. Stock PyTorch runs
this as a cuBLASLt GEMM plus an elementwise kernel that re-reads the
``[tokens, 4·d_model]`` pre-activation. Backward: ``gelu_backward`` + two
cuBLAS GEMMs (dX, dW) + a bias-gradient reduction.

The workloads it serves are the reference's BERT and transformer examples
(``examples/BERT/model.py:91-113`` feed-forward ``linear1 -> gelu``).

On CPU, for non-bf16 inputs, or shapes outside the kernel's tiling
(K % 64, N % 128) the op is the plain PyTorch composition.
"""

import os

import torch
import torch.nn.functional as F

from adaptdl_b200.ops import _count

_ERR = {}            # device index -> int32 error flag tensor
_ACT = {None: 0, "identity": 0, "gelu": 1}


def _err_flag(device):
    flag = _ERR.get(device.index)
    if flag is None:
        flag = torch.zeros(1, dtype=torch.int32, device=device)
        _ERR[device.index] = flag
    return flag


def check_errors(device=None):
    """Raise if any fused GEMM reported a stuck pipeline (debug aid; syncs)."""
    for idx, flag in _ERR.items():
        if device is not None and torch.device(device).index != idx:
            continue
        code = int(flag.item())
        if code:
            flag.zero_()
            raise RuntimeError(
                "adl_gemm_bias_act: pipeline timeout (flag {})".format(code))


def supported(x, weight):
    if not x.is_cuda or os.environ.get("ADAPTDL_B200_FUSED_GEMM", "1") == "0":
        return False
    n, k = weight.shape
    return k % 64 == 0 and n % 128 == 0 and x.shape[-1] == k and \
        x.numel() // k > 0


def gemm_bias_act(x2d, weight, bias, act="gelu", save_preact=True,
                  block_n=0, cluster_m=0, max_ctas=0, trace=None):
    """Raw kernel call. ``x2d`` [M, K] bf16, ``weight`` [N, K] bf16, ``bias``
    [N] fp32 or None. Returns ``(y, z)`` (``z`` is None unless
    ``save_preact``). ``block_n`` (128/256) and ``cluster_m`` (1/2/4 CTAs
    sharing a multicast weight tile, or 22 = CTA pair issuing 2-SM
    ``tcgen05.mma.cta_group::2``) default to a shape-based choice.
    ``trace``: optional int64 ``[3, 256]`` tensor that receives SM-clock
    stamps of CTA 0's pipeline (TMA issue, MMA wait begin / end per
    K-slice)."""
    from adaptdl_b200 import _native
    lib = _native.load()
    assert x2d.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
    x2d = x2d.contiguous()
    weight = weight.contiguous()
    m, k = x2d.shape
    n = weight.shape[0]
    if bias is not None:
        bias = bias.float().contiguous()
    y = torch.empty((m, n), dtype=torch.bfloat16, device=x2d.device)
    z = torch.empty_like(y) if save_preact else None
    dev = x2d.device.index if x2d.device.index is not None \
        else torch.cuda.current_device()
    lib.adl_set_device(dev)
    code = lib.adl_gemm_bias_act(
        x2d.data_ptr(), weight.data_ptr(),
        bias.data_ptr() if bias is not None else None,
        y.data_ptr(), z.data_ptr() if z is not None else None,
        m, n, k, _ACT[act] if isinstance(act, (str, type(None))) else int(act),
        block_n, cluster_m, max_ctas,
        _err_flag(x2d.device).data_ptr(),
        trace.data_ptr() if trace is not None else None,
        torch.cuda.current_stream(x2d.device).cuda_stream)
    if code < 0:
        raise RuntimeError(
            "adl_gemm_bias_act rejected M={} N={} K={} (code {})".format(
                m, n, k, code))
    _native.check(code, "adl_gemm_bias_act")
    _count.add(1)
    return y, z


def _gelu_dropout_backward(dy2d, z, keep, scale):
    """``dy * keep * scale * gelu'(z)`` in one pass (``keep`` may be
    ``None``)."""
    from adaptdl_b200 import _native
    lib = _native.load()
    dy2d = dy2d.contiguous()
    dz = torch.empty_like(z)
    code = lib.adl_gelu_dropout_bwd(
        dy2d.data_ptr(), z.data_ptr(),
        keep.data_ptr() if keep is not None else None, dz.data_ptr(),
        z.numel(), float(scale), 1 if z.dtype == torch.bfloat16 else 2,
        torch.cuda.current_stream(z.device).cuda_stream)
    if code < 0:
        raise RuntimeError("adl_gelu_dropout_bwd rejected the call ({})"
                           .format(code))
    _native.check(code, "adl_gelu_dropout_bwd")
    _count.add(1)
    return dz


class _LinearAct(torch.autograd.Function):
    """``dropout(act(x @ W^T + b), p)``: the GEMM with its fused bias + GELU
    epilogue, ``aten::native_dropout`` for the mask (generator-driven, CUDA-
    graph safe), and a backward whose dropout and GELU derivatives are ONE
    elementwise pass (instead of ``masked_scale`` + ``gelu_backward``)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, p):
        x2d = x.reshape(-1, x.shape[-1])
        need_grad = any(ctx.needs_input_grad[:3])
        y, z = gemm_bias_act(x2d, weight, bias, act,
                             save_preact=need_grad and act == "gelu")
        keep = None
        if p > 0.0:
            y, keep = torch.native_dropout(y, p, True)
        ctx.act = act
        ctx.scale = 1.0 / (1.0 - p) if p > 0.0 else 1.0
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x2d, weight, z, keep)
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2d, weight, z, keep = ctx.saved_tensors
        dy2d = dy.reshape(-1, dy.shape[-1])
        if ctx.act == "gelu":
            dz = _gelu_dropout_backward(dy2d, z, keep, ctx.scale)
        elif keep is not None:
            dz = dy2d * keep * ctx.scale
        else:
            dz = dy2d
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = (dz @ weight).view(*dy.shape[:-1], weight.shape[1])
        if ctx.needs_input_grad[1]:
            dw = dz.t() @ x2d
        if ctx.has_bias and ctx.needs_input_grad[2]:
            from adaptdl_b200.ops.transformer import colsum
            db = colsum(dz.contiguous())
        return dx, dw, db, None, None


def linear_act(x, weight, bias=None, act="gelu", dropout_p=0.0,
               training=True):
    """``dropout(act(x @ weight.T + bias), dropout_p)``; fused tcgen05
    kernel for bf16 CUDA inputs (or under bf16 autocast), PyTorch composition
    otherwise."""
    p = float(dropout_p) if training else 0.0
    use_bf16 = x.is_cuda and (
        x.dtype == torch.bfloat16 or
        (torch.is_autocast_enabled("cuda") and
         torch.get_autocast_dtype("cuda") == torch.bfloat16))
    if use_bf16 and supported(x, weight):
        with torch.autocast("cuda", enabled=False):
            out = _LinearAct.apply(
                x.to(torch.bfloat16), weight.to(torch.bfloat16),
                bias.float() if bias is not None else None, act, p)
        return out
    out = F.linear(x, weight, bias)
    out = F.gelu(out) if act == "gelu" else out
    return F.dropout(out, p, True) if p > 0.0 else out


class LinearGELU(torch.nn.Linear):
    """``nn.Linear`` whose forward also applies GELU (fused on B200).
    Parameter names/shapes are those of ``nn.Linear``, so checkpoints are
    interchangeable with ``Linear`` + ``F.gelu``."""

    def forward(self, x):
        return linear_act(x, self.weight, self.bias, "gelu")
