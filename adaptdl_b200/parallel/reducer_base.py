"""Gradient reducer: owns flat gradient storage, bucketing, the per-backward
control flow, and the gradient statistics that feed the noise-scale
estimator.

This replaces what the reference gets from ``DistributedDataParallel``'s C++
Reducer plus its own per-parameter hooks:

============================  ============================================
reference (per parameter)     here (per bucket, one fused primitive each)
============================  ============================================
DDP bucket copy + all-reduce  ``REDUCE``: in-place mean over replicas
``(g/P).pow(2).sum()`` hook   fused into FOLD / REDUCE (L statistic)
``grad.div_(accum_count)``    folded into REDUCE's scale factor
``grad.float()/amp`` copies   none (kernels cast in registers)
``_normsqr_groups`` + .item   fused into REDUCE / PAIR (T statistic)
``prev_grads`` clone          PAIR writes the stash in the same pass
============================  ============================================

Primitives (implemented by subclasses on a bucket of the flat arena; G is
the gradient arena whose views are ``param.grad``, A the accumulation arena,
Pv the previous-step stash; L/T are per-param-group float64 vectors):

``FOLD_ACC``    A += G;  L += |G|^2;  G = 0          (accumulation step)
``FOLD_FINAL``  L += |G|^2;  G += A;  A = 0          (last micro-step, k>1)
``REDUCE``      G = s * sum_r G_r;  T += |G|^2; optionally L += sum_r |G_r|^2
``PAIR``        T=|G|^2; Pp=|Pv|^2; Pa=|(G+Pv)/2|^2; Pv = G   (N=1, k=1)

|x|^2 is taken per param group, after dividing by the optional
preconditioner.
"""

import logging
import time

import numpy as np
import torch
from torch.autograd import Variable

from adaptdl_b200.parallel import layout
from adaptdl_b200.utils import print_exc
from adaptdl_b200.utils.trace import traced

LOG = logging.getLogger(__name__)


class GradStats(object):
    """Host-side statistics of one optimizer step (amp-scaled units)."""

    __slots__ = ("local_sqr", "total_sqr", "count", "pair", "sync_time")

    def __init__(self, local_sqr, total_sqr, count, pair=None,
                 sync_time=None):
        self.local_sqr = local_sqr   # sum over replicas & micro-steps of |g|^2
        self.total_sqr = total_sqr   # |averaged gradient|^2
        self.count = count           # replicas * micro-steps
        self.pair = pair             # (prev_sqr, avg_sqr) or None
        self.sync_time = sync_time   # seconds, or None


class _Arena(object):
    """All gradients of one dtype: flat storage + bucket plan."""

    def __init__(self, dtype, params, param_ids, groups):
        self.dtype = dtype
        self.params = params          # registration order
        self.param_ids = param_ids    # global ordinal of each param
        self.groups = groups          # param-group index per param
        self.total = 0
        self.buckets = []
        self.grad = None              # flat G
        self.acc = None               # flat A (lazy)
        self.prev = None              # flat Pv (lazy)
        self.pinv = None              # flat preconditioner (lazy)
        self.views = []               # per param view into ``grad``
        self.bucket_of = {}           # local param index -> bucket index
        self.pending = []             # per bucket: grads still expected
        self.done = []                # per bucket: processed this backward


class GradReducer(object):
    """Base class; see module docstring. Subclasses implement the
    primitives and storage allocation.

    Arguments:
        param_groups: ``optimizer.param_groups`` (group index = statistics
            group).
        world_size, rank: data-parallel replicas and this replica's rank.
        should_sync: callable -> bool, asked at the start of every backward:
            is this the micro-step that ends with a sync + optimizer step?
        bucket_cap_mb: soft cap of one bucket (one fused launch).
    """

    FIRST_BUCKET_BYTES = 1 << 20
    # The parameters backward reaches LAST (the first layers) get a small
    # bucket of their own: its reduction is the only one that cannot overlap
    # with backward, so it should be latency- not bandwidth-sized.
    LAST_BUCKET_BYTES = 64 << 10

    def __init__(self, param_groups, world_size, rank, should_sync,
                 bucket_cap_mb=25, name="reducer"):
        self.world_size = int(world_size)
        self.rank = int(rank)
        self.name = name
        self._should_sync = should_sync
        self.num_groups = len(param_groups)
        self._bucket_cap = int(bucket_cap_mb * 1024 * 1024)
        self._accum_count = 0
        self._in_backward = False
        self._sync = True
        self._k_before = 0
        self._prev_valid = False
        self._stats_ready = None
        self._precond_fn = None
        self._hooks = []
        self._on_backward_end = None   # callable(sync: bool)
        self._sync_t0 = None
        self._n_done = 0
        self.launches = 0            # fused-primitive launches (telemetry)
        self.device = None

        by_dtype = {}
        seen = set()
        ordinal = 0
        for gidx, group in enumerate(param_groups):
            for p in group["params"]:
                if not p.requires_grad or id(p) in seen:
                    continue
                seen.add(id(p))
                if p.is_sparse:
                    raise ValueError("sparse parameters are not supported")
                if self.device is None:
                    self.device = p.device
                elif p.device != self.device:
                    raise ValueError("all parameters must live on one "
                                     "device (one replica = one GPU)")
                slot = by_dtype.setdefault(p.dtype, ([], [], []))
                slot[0].append(p)
                slot[1].append(ordinal)
                slot[2].append(gidx)
                ordinal += 1
        if self.device is None:
            self.device = torch.device("cpu")
        self.arenas = [_Arena(dt, *slot) for dt, slot in by_dtype.items()]
        for arena in self.arenas:
            itemsize = torch.empty((), dtype=arena.dtype).element_size()
            arena.total, arena.buckets = layout.plan_arena(
                [p.numel() for p in arena.params], arena.groups, itemsize,
                self._bucket_cap, self.FIRST_BUCKET_BYTES, self.world_size,
                last_bucket_cap_bytes=self.LAST_BUCKET_BYTES)
            for b in arena.buckets:
                for seg in b.segments:
                    arena.bucket_of[seg.param_index] = b.index
        self._attach()

    # ------------------------------------------------------------------
    # storage
    # ------------------------------------------------------------------

    def _alloc_flat(self, arena, kind):
        """Allocate a zeroed flat buffer of ``arena.total`` elements.
        ``kind`` in {"grad", "acc", "prev", "pinv"}."""
        return torch.zeros(max(arena.total, 1), dtype=arena.dtype,
                           device=self.device)

    def _attach(self):
        for arena in self.arenas:
            arena.grad = self._alloc_flat(arena, "grad")
            arena.views = [None] * len(arena.params)
            for b in arena.buckets:
                for seg in b.segments:
                    p = arena.params[seg.param_index]
                    piece = arena.grad[seg.start:seg.start + seg.length]
                    if p.is_contiguous() or not _is_dense(p):
                        view = piece.view(p.shape)
                    else:
                        # dense permuted layout (channels_last): give .grad
                        # the parameter's strides so autograd's accumulation
                        # and the optimizer run on matching layouts; the
                        # kernels only ever see the flat bytes.
                        view = piece.as_strided(p.shape, p.stride())
                    arena.views[seg.param_index] = view
                    if p.grad is not None:
                        view.copy_(p.grad)
                    p.grad = view
            arena.pending = [len(b.segments) for b in arena.buckets]
            arena.done = [False] * len(arena.buckets)
            for local_idx, p in enumerate(arena.params):
                handle = p.register_post_accumulate_grad_hook(
                    self._make_hook(arena, local_idx))
                self._hooks.append(handle)

    def detach(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def _ensure(self, arena, kind):
        buf = getattr(arena, kind)
        if buf is None:
            buf = self._alloc_flat(arena, kind)
            setattr(arena, kind, buf)
        return buf

    # ------------------------------------------------------------------
    # public control
    # ------------------------------------------------------------------

    @property
    def accum_count(self):
        """Micro-steps accumulated since the last :meth:`zero`."""
        return self._accum_count

    def set_preconditioner(self, fn):
        """``fn(param) -> tensor|None``: element-wise divisor applied before
        every squared norm (Adam-preconditioned statistics)."""
        self._precond_fn = fn

    def zero(self):
        """Zero gradients and reset accumulation (the patched
        ``optimizer.zero_grad``)."""
        for arena in self.arenas:
            arena.grad.zero_()
            if arena.acc is not None and self._accum_count > 0:
                arena.acc.zero_()
        self._accum_count = 0
        self._reset_partials()

    def invalidate_stash(self):
        """Forget the previous-step gradient (after a non-finite step)."""
        self._prev_valid = False

    @traced("pop_stats")
    def pop_stats(self):
        """Statistics of the last synchronised backward (blocks until the
        device has produced them), or ``None``."""
        handle, self._stats_ready = self._stats_ready, None
        if handle is None:
            return None
        return self._resolve_stats(handle)

    # ------------------------------------------------------------------
    # backward-pass control flow
    # ------------------------------------------------------------------

    def _make_hook(self, arena, local_idx):
        @print_exc
        def hook(param):
            if not self._in_backward:
                self._begin_backward()
            view = arena.views[local_idx]
            grad = param.grad
            if grad is not None and grad.data_ptr() != view.data_ptr():
                # someone replaced .grad (e.g. zero_grad(set_to_none=True)
                # on the inner module): fold it back into the arena.
                view.copy_(grad)
                param.grad = view
            b = arena.bucket_of[local_idx]
            arena.pending[b] -= 1
            if arena.pending[b] == 0 and not arena.done[b]:
                self._process_bucket(arena, b)
        return hook

    def _begin_backward(self):
        self._in_backward = True
        self._sync = bool(self._should_sync())
        self._k_before = self._accum_count
        self._n_done = 0
        if self._precond_fn is not None and not self._device_preconditioner():
            self._refresh_preconditioner()
        self._on_begin_backward()
        Variable._execution_engine.queue_callback(self._end_backward)

    @print_exc
    @traced("backward_end")
    def _end_backward(self):
        # buckets whose parameters did not all receive gradients (unused
        # parameters) are flushed here, in order.
        for arena in self.arenas:
            for b in range(len(arena.buckets)):
                if not arena.done[b]:
                    self._process_bucket(arena, b)
            arena.pending = [len(bk.segments) for bk in arena.buckets]
            arena.done = [False] * len(arena.buckets)
        self._accum_count += 1
        if self._sync:
            self._stats_ready = self._finalize_step()
        else:
            self._mark_accum_step()
        self._in_backward = False
        if self._on_backward_end is not None:
            self._on_backward_end(self._sync)

    @traced("bucket")
    def _process_bucket(self, arena, b):
        arena.done[b] = True
        bucket = arena.buckets[b]
        self._n_done += 1
        # the step's last bucket: everything after it (statistics exchange,
        # estimator, optimizer) is on the critical path, so the subclass may
        # fold the finalize into this launch
        last = self._sync and self._n_done == self.num_buckets
        if last:
            # "end of the local backward" for the sync-time measurement
            self._mark_sync_start()
        if not self._sync:
            self._fold_acc(arena, bucket)
        elif self._k_before > 0:
            self._fold_final(arena, bucket)
            scale = 1.0 / (self.world_size * (self._k_before + 1))
            self._reduce(arena, bucket, scale, want_local=False, last=last)
        elif self.world_size > 1:
            self._reduce(arena, bucket, 1.0 / self.world_size,
                         want_local=True, last=last)
        else:
            self._pair(arena, bucket, last=last)

    @property
    def num_buckets(self):
        return sum(len(arena.buckets) for arena in self.arenas)

    def _refresh_preconditioner(self):
        for arena in self.arenas:
            pinv = self._ensure(arena, "pinv")
            for b in arena.buckets:
                for seg in b.segments:
                    p = arena.params[seg.param_index]
                    val = self._precond_fn(p)
                    dst = pinv[seg.start:seg.start + seg.length]
                    if val is None:
                        dst.fill_(1.0)
                    else:
                        dst.copy_(val.reshape(-1))
            # padding must not produce 0/0
            self._fill_padding_ones(arena, pinv)

    def _fill_padding_ones(self, arena, pinv):
        cursor = 0
        for b in sorted(arena.buckets, key=lambda x: x.start):
            for seg in sorted(b.segments, key=lambda s: s.start):
                if seg.start > cursor:
                    pinv[cursor:seg.start].fill_(1.0)
                cursor = seg.start + seg.length
        if cursor < pinv.numel():
            pinv[cursor:].fill_(1.0)

    # ------------------------------------------------------------------
    # to be provided by subclasses
    # ------------------------------------------------------------------

    def _on_begin_backward(self):
        pass

    def _device_preconditioner(self):
        """True if the kernels derive the preconditioner themselves (no
        per-parameter refresh on the host)."""
        return False

    def _mark_sync_start(self):
        self._sync_t0 = time.time()

    def _mark_accum_step(self):
        """End of an accumulation micro-step (no synchronisation)."""

    def _reset_partials(self):
        raise NotImplementedError

    def _fold_acc(self, arena, bucket):
        raise NotImplementedError

    def _fold_final(self, arena, bucket):
        raise NotImplementedError

    def _reduce(self, arena, bucket, scale, want_local, last=False):
        raise NotImplementedError

    def _pair(self, arena, bucket, last=False):
        raise NotImplementedError

    def _finalize_step(self):
        raise NotImplementedError

    def _resolve_stats(self, handle):
        raise NotImplementedError

    def broadcast_parameters(self, tensors, src=0):
        """Broadcast ``tensors`` (params/buffers) from ``src`` in place."""
        raise NotImplementedError


def _is_dense(p):
    try:
        return torch.empty_like(p).stride() == p.stride()
    except RuntimeError:
        return False


def count_of(reducer):
    return reducer.world_size * max(reducer.accum_count, 1)


def as_numpy64(t):
    return np.asarray(t.detach().cpu().numpy(), dtype=np.float64)
