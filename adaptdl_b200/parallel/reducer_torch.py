"""Reference implementation of the reducer primitives with stock torch ops
and ``torch.distributed`` collectives (gloo on CPU, NCCL on GPU).

This is (a) the plumbing path that runs anywhere, (b) the correctness oracle
for the fused sm_100a kernels in ``reducer_cuda.py``, and (c) the fallback
when peer mappings cannot be established. It is deliberately simple: one
``all_reduce`` per bucket, segment sums for the statistics.
"""

import time

import numpy as np
import torch
import torch.distributed as dist

from adaptdl_b200.parallel.reducer_base import GradReducer, GradStats


class TorchGradReducer(GradReducer):

    def __init__(self, param_groups, world_size, rank, should_sync,
                 bucket_cap_mb=25, process_group=None, name="reducer"):
        self._pg = process_group
        self._seg_cache = {}
        super().__init__(param_groups, world_size, rank, should_sync,
                         bucket_cap_mb, name)
        n = self.num_groups + 1          # last slot collects padding
        dev = self.device
        self._L = torch.zeros(n, dtype=torch.float64, device=dev)
        self._T = torch.zeros(n, dtype=torch.float64, device=dev)
        self._Pp = torch.zeros(n, dtype=torch.float64, device=dev)
        self._Pa = torch.zeros(n, dtype=torch.float64, device=dev)
        self._had_pair = False
        self._sync_events = None

    # -- helpers ---------------------------------------------------------

    def _seg_index(self, arena, bucket):
        """(lengths, group-of-piece) covering the whole bucket, padding
        pieces mapped to the dummy group ``num_groups``."""
        key = (id(arena), bucket.index)
        hit = self._seg_cache.get(key)
        if hit is not None:
            return hit
        lengths, groups = [], []
        cursor = bucket.start
        for seg in sorted(bucket.segments, key=lambda s: s.start):
            if seg.start > cursor:
                lengths.append(seg.start - cursor)
                groups.append(self.num_groups)
            lengths.append(seg.length)
            groups.append(seg.group)
            cursor = seg.start + seg.length
        end = bucket.start + bucket.length
        if end > cursor:
            lengths.append(end - cursor)
            groups.append(self.num_groups)
        hit = (torch.tensor(lengths, dtype=torch.int64, device=self.device),
               torch.tensor(groups, dtype=torch.int64, device=self.device))
        self._seg_cache[key] = hit
        return hit

    def _slice(self, arena, bucket, kind):
        buf = self._ensure(arena, kind)
        return buf[bucket.start:bucket.start + bucket.length]

    def _add_norms(self, out, arena, bucket, values):
        """out[group] += sum over the group's elements of (values/P)^2."""
        x = values.to(torch.float64)
        if arena.pinv is not None and self._precond_fn is not None:
            x = x / self._slice(arena, bucket, "pinv").to(torch.float64)
        lengths, groups = self._seg_index(arena, bucket)
        sums = torch.segment_reduce(x * x, "sum", lengths=lengths,
                                    unsafe=True)
        out.index_add_(0, groups, sums)
        self.launches += 1

    # -- primitives --------------------------------------------------------

    def _reset_partials(self):
        if hasattr(self, "_L"):
            self._L.zero_()
            self._T.zero_()
            self._Pp.zero_()
            self._Pa.zero_()

    def _on_begin_backward(self):
        if self._sync:
            self._T.zero_()
            self._Pp.zero_()
            self._Pa.zero_()
            self._had_pair = False
        if self._k_before == 0:
            self._L.zero_()

    def _mark_sync_start(self):
        if self.device.type == "cuda":
            start = torch.cuda.Event(enable_timing=True)
            start.record()
            self._sync_events = [start, None]
        else:
            self._sync_t0 = time.time()

    def _fold_acc(self, arena, bucket):
        g = self._slice(arena, bucket, "grad")
        a = self._slice(arena, bucket, "acc")
        self._add_norms(self._L, arena, bucket, g)
        a.add_(g)
        g.zero_()

    def _fold_final(self, arena, bucket):
        g = self._slice(arena, bucket, "grad")
        a = self._slice(arena, bucket, "acc")
        self._add_norms(self._L, arena, bucket, g)
        g.add_(a)
        a.zero_()

    def _reduce(self, arena, bucket, scale, want_local, last=False):
        g = self._slice(arena, bucket, "grad")
        if want_local:
            self._add_norms(self._L, arena, bucket, g)
        if self.world_size > 1:
            dist.all_reduce(g, group=self._pg)
        if scale != 1.0:
            g.mul_(scale)
        self._add_norms(self._T, arena, bucket, g)

    def _pair(self, arena, bucket, last=False):
        g = self._slice(arena, bucket, "grad")
        pv = self._slice(arena, bucket, "prev")
        self._add_norms(self._T, arena, bucket, g)
        if self._prev_valid:
            self._add_norms(self._Pp, arena, bucket, pv)
            self._add_norms(self._Pa, arena, bucket, (g + pv) / 2)
            self._had_pair = True
        pv.copy_(g)

    def _finalize_step(self):
        pair_mode = self.world_size == 1 and self._k_before == 0
        if pair_mode:
            self._prev_valid = True
        elif self.world_size > 1:
            dist.all_reduce(self._L, group=self._pg)
        rows = [self._L, self._T]
        if pair_mode and self._had_pair:
            rows += [self._Pp, self._Pa]
        packed = torch.stack(rows)[:, :self.num_groups]
        count = self.world_size * self._accum_count
        if self.device.type == "cuda":
            host = torch.empty(packed.shape, dtype=packed.dtype,
                               pin_memory=True)
            host.copy_(packed, non_blocking=True)
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            self._sync_events[1] = end
            return (host, end, count, self._sync_events)
        sync_time = time.time() - self._sync_t0 if self._sync_t0 else 0.0
        return (packed.clone(), None, count, sync_time)

    def _resolve_stats(self, handle):
        host, event, count, timing = handle
        if event is not None:
            event.synchronize()
            sync_time = timing[0].elapsed_time(timing[1]) / 1e3
        else:
            sync_time = timing
        arr = np.asarray(host.numpy(), dtype=np.float64)
        pair = (arr[2], arr[3]) if arr.shape[0] == 4 else None
        return GradStats(arr[0].copy(), arr[1].copy(), count, pair,
                         sync_time)

    def broadcast_parameters(self, tensors, src=0):
        if self.world_size <= 1:
            return
        tensors = [t for t in tensors if t is not None]
        if not tensors:
            return
        if hasattr(dist, "_broadcast_coalesced"):
            pg = self._pg or dist.group.WORLD
            dist._broadcast_coalesced(pg, [t.detach() for t in tensors],
                                      256 * 1024 * 1024, src)
        else:  # pragma: no cover
            for t in tensors:
                dist.broadcast(t.detach(), src, group=self._pg)
