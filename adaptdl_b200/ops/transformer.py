"""Layout and reduction ops of the transformer workloads
(``csrc/adl_transformer.cu``).

PyTorch's generic strided-copy / ``cat`` / ``sum`` kernels run at 5-25 % of
the copy bandwidth on the layout boundaries of a BERT step (fused-QKV split,
attention-head merge, bias gradients, the padded MLM logits); these ops do
the same data movement with 16-byte vectors and coalesced accesses:

``split_heads(qkv, nhead, parts)``   ``[N, S, parts*H*D] -> parts x [N, H, S, D]``
    (contiguous outputs: what cuDNN's attention kernels like best); the
    backward packs the ``parts`` gradients straight into ``[N, S, parts*H*D]``
    (instead of ``aten::cat`` + a strided ``copy_``).
``merge_heads(x)``                   ``[N, H, S, D] -> [N, S, H*D]``
``linear(x, weight, bias)``          ``F.linear`` whose bias gradient is one
    deterministic column-sum kernel.
``padded_logits(x, weight, bias)``   a Linear whose output width is not a
    multiple of 8 (BERT's 28 996-token MLM head): the GEMMs run on a weight
    padded to a multiple of 64 (tensor-core kernels for forward, dgrad and
    wgrad), and the user-visible contiguous ``[..., n]`` logits are produced by
    ONE pass -- in fp32 under autocast (what ``cross_entropy`` would cast them
    to anyway), so the slice, the cast and their two backward passes become
    two kernels instead of four.

All have pure-PyTorch fallbacks (CPU, unsupported shapes / dtypes).
"""

import os

import torch
import torch.nn.functional as F

from adaptdl_b200.ops import _count

_TICKETS = {}
_DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def _enabled(t):
    return t.is_cuda and os.environ.get("ADAPTDL_B200_FUSED_TRANSFORMER",
                                        "1") != "0"


def _lib(device):
    from adaptdl_b200 import _native
    lib = _native.load()
    lib.adl_set_device(device.index if device.index is not None
                       else torch.cuda.current_device())
    return lib


def _check(code, what):
    from adaptdl_b200 import _native
    if code < 0:
        raise RuntimeError("{} rejected the call (code {})".format(what,
                                                                    code))
    _native.check(code, what)
    _count.add(1)


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


# ---------------------------------------------------------------------------
# attention heads
# ---------------------------------------------------------------------------

def _heads_ok(t, d):
    return _enabled(t) and t.dtype in (torch.bfloat16, torch.float16) \
        and d % 8 == 0


def _permute(srcs, dst, a, b, w, h, d, merge):
    """split: ``srcs`` is the packed tensor; merge: a list of ``w`` planes
    ``[a, h, b, d]`` (separately allocated, contiguous)."""
    if torch.is_tensor(srcs):
        srcs = [srcs]
    lib = _lib(dst.device)
    ptrs = [t.data_ptr() for t in srcs] + [None] * (3 - len(srcs))
    _check(lib.adl_heads_permute(ptrs[0], ptrs[1], ptrs[2], dst.data_ptr(),
                                 a, b, w, h, d, merge, _stream(dst.device)),
           "adl_heads_permute")


class _SplitHeads(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, nhead, parts):
        n, s, e = qkv.shape
        d = e // (parts * nhead)
        qkv = qkv.contiguous()
        out = torch.empty((parts, n, nhead, s, d), dtype=qkv.dtype,
                          device=qkv.device)
        _permute(qkv, out, n, s, parts, nhead, d, 0)
        ctx.dims = (n, s, parts, nhead, d)
        return tuple(out[i] for i in range(parts))

    @staticmethod
    def backward(ctx, *grads):
        n, s, parts, nhead, d = ctx.dims
        ref = next(g for g in grads if g is not None)
        # the attention backward hands over three separately allocated
        # [N, H, S, D] gradients: the pack kernel reads them in place
        planes = [torch.zeros_like(ref) if g is None else g.contiguous()
                  for g in grads]
        dqkv = torch.empty((n, s, parts * nhead * d), dtype=ref.dtype,
                           device=ref.device)
        _permute(planes, dqkv, n, s, parts, nhead, d, 1)
        return dqkv, None, None


class _MergeHeads(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        n, h, s, d = x.shape
        x = x.contiguous()
        out = torch.empty((n, s, h * d), dtype=x.dtype, device=x.device)
        _permute([x], out, n, s, 1, h, d, 1)
        ctx.dims = (n, h, s, d)
        return out

    @staticmethod
    def backward(ctx, grad):
        n, h, s, d = ctx.dims
        grad = grad.contiguous()
        out = torch.empty((n, h, s, d), dtype=grad.dtype, device=grad.device)
        _permute(grad, out, n, s, 1, h, d, 0)
        return out


def split_heads(qkv, nhead, parts=3):
    """``qkv`` ``[N, S, parts * nhead * D]`` -> ``parts`` contiguous tensors
    ``[N, nhead, S, D]``."""
    n, s, e = qkv.shape
    d = e // (parts * nhead)
    if _heads_ok(qkv, d):
        return _SplitHeads.apply(qkv, nhead, parts)
    x = qkv.view(n, s, parts, nhead, d).permute(2, 0, 3, 1, 4)
    return tuple(x[i] for i in range(parts))


def merge_heads(x):
    """``[N, H, S, D]`` -> ``[N, S, H * D]``."""
    n, h, s, d = x.shape
    if _heads_ok(x, d):
        return _MergeHeads.apply(x)
    return x.transpose(1, 2).reshape(n, s, h * d)


# ---------------------------------------------------------------------------
# column sums (bias gradients)
# ---------------------------------------------------------------------------

def _tickets(device, n):
    buf = _TICKETS.get(device.index)
    if buf is None or buf.numel() < n:
        buf = torch.zeros(max(n, 1024), dtype=torch.int32, device=device)
        _TICKETS[device.index] = buf
    return buf


def colsum(x2d):
    """fp32 column sums of a contiguous ``[M, N]`` matrix."""
    m, n = x2d.shape
    vec = 4 if x2d.dtype == torch.float32 else 8
    if not (_enabled(x2d) and x2d.dtype in _DT and n % vec == 0
            and x2d.is_contiguous() and m > 0):
        return x2d.sum(0, dtype=torch.float32)
    dev = x2d.device
    blocks = (n + 63) // 64
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    rows_per_iter = 256 // (64 // vec)
    chunks = max(1, min((2 * sms + blocks - 1) // blocks,
                        (m + 4 * rows_per_iter - 1) // (4 * rows_per_iter)))
    out = torch.empty(n, dtype=torch.float32, device=dev)
    partial = torch.empty(blocks * chunks * 64, dtype=torch.float32,
                          device=dev)
    lib = _lib(dev)
    _check(lib.adl_colsum(x2d.data_ptr(), out.data_ptr(), partial.data_ptr(),
                          _tickets(dev, blocks).data_ptr(), m, n,
                          _DT[x2d.dtype], chunks, _stream(dev)),
           "adl_colsum")
    return out


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        # the bias stays in its own (fp32) dtype outside: its gradient is
        # returned in that dtype, straight from the fp32 column sums
        return F.linear(x, weight,
                        bias.to(x.dtype) if bias is not None else None)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = (dy2 @ weight).view(x.shape)
        if ctx.needs_input_grad[1]:
            dw = dy2.t() @ x.reshape(-1, x.shape[-1])
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2.contiguous())     # fp32; autograd casts if needed
        return dx, dw, db


def linear(x, weight, bias=None):
    """``F.linear`` (same numerics: the GEMMs are the library's) with the
    bias gradient computed by :func:`colsum`."""
    if not _enabled(x) or bias is None:
        return F.linear(x, weight, bias)
    if torch.is_autocast_enabled("cuda"):
        dt = torch.get_autocast_dtype("cuda")
        with torch.autocast("cuda", enabled=False):
            return _Linear.apply(x.to(dt), weight.to(dt), bias)
    return _Linear.apply(x, weight, bias)


# ---------------------------------------------------------------------------
# padded logits (vocabulary not a multiple of 8)
# ---------------------------------------------------------------------------

class _SliceCast(torch.autograd.Function):
    """``padded[..., :n].float()`` in one pass; backward: one pass into a
    zero-padded bf16 buffer."""

    @staticmethod
    def forward(ctx, padded, n):
        ld = padded.shape[-1]
        m = padded.numel() // ld
        out = torch.empty(padded.shape[:-1] + (n,), dtype=torch.float32,
                          device=padded.device)
        lib = _lib(padded.device)
        _check(lib.adl_slice_cast(padded.data_ptr(), out.data_ptr(), m, n,
                                  ld, 0, _stream(padded.device)),
               "adl_slice_cast")
        ctx.dims = (padded.shape, n, ld, m)
        return out

    @staticmethod
    def backward(ctx, grad):
        shape, n, ld, m = ctx.dims
        grad = grad.contiguous().float()
        out = torch.empty(shape, dtype=torch.bfloat16, device=grad.device)
        lib = _lib(grad.device)
        _check(lib.adl_slice_cast(grad.data_ptr(), out.data_ptr(), m, n, ld,
                                  1, _stream(grad.device)), "adl_slice_cast")
        return out, None


def padded_logits(x, weight, bias, multiple=64):
    """``F.linear(x, weight, bias)`` for an output width ``n`` that is not a
    multiple of 8. On CUDA under bf16 autocast the result is fp32 (see the
    module docstring); otherwise it has ``x``'s dtype."""
    n = weight.shape[0]
    pad = (-n) % multiple
    if pad == 0 or not x.is_cuda:
        return F.linear(x, weight, bias)
    wp = F.pad(weight, (0, 0, 0, pad))
    bp = F.pad(bias, (0, pad)) if bias is not None else None
    yp = F.linear(x, wp, bp)
    if _enabled(yp) and yp.dtype == torch.bfloat16 and n % 4 == 0 and \
            yp.is_contiguous() and torch.is_autocast_enabled("cuda"):
        return _SliceCast.apply(yp, n)
    return yp[..., :n].contiguous()
