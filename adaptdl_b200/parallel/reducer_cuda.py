"""Gradient reducer on hand-written sm_100a kernels over NVLink peer memory.

One fused launch per bucket (``adl_allreduce_gns``, csrc/adl_kernels.cu):
two-shot all-reduce by direct loads/stores on the peers' gradient arenas,
``1/(N*accum)`` scaling, and both gradient-noise-scale statistics from the
same registers -- replacing, per step, the reference's NCCL bucket
all-reduces, ~10 element-wise/reduction launches per parameter, a second
NCCL all-reduce for the statistics, one full fp32 copy of the gradients and
``1 + num_param_groups`` host synchronisations.

Three flavours of the bucket kernel, chosen per bucket when the reducer is
built (all ranks derive the same choice from the bucket plan): one-shot push
for latency-bound buckets, NVLS (``multimem``) for the largest ones, two-shot
P2P in between. The kernel of the step's LAST bucket also runs the statistics
exchange and the noise-scale estimator in its last CTA (``fuse_fin``), so the
exposed tail of a step is one small launch with two flag rounds.

Stream plumbing: every primitive runs on a dedicated high-priority
communication stream, ordered after the producing backward kernels by an
event; the compute stream waits for the communication stream once, at the
end of backward. Statistics travel to the host through a pinned mailbox
written by ``adl_finalize_stats`` (no ``.item()``); the host waits for them
lazily, on a CUDA event, when it first needs them.
"""

import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from adaptdl_b200 import _native
from adaptdl_b200._native import (ReduceArgs, LocalArgs, FinalizeArgs,
                                  BcastArgs, MAX_RANKS, MAX_CTAS, MBOX_HDR,
                                  XCHG_TAIL, CLOCK_DOUBLES, SITES_PER_STEP,
                                  FLAVOUR_TWOSHOT, FLAVOUR_ONESHOT,
                                  FLAVOUR_NVLS, PINV_FLAT, PINV_ADAM,
                                  MB_SYNC_NS, MB_ERR, check)
from adaptdl_b200.parallel import layout, symm
from adaptdl_b200.parallel.reducer_base import GradReducer, GradStats

_DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
_PAD_BYTES = 2 * MAX_CTAS * MAX_RANKS * 4
_STAGING_BYTES = 16 << 20
_ALIGN = 512
_TIMEOUT_NS = int(float(os.environ.get("ADAPTDL_B200_TIMEOUT_S", "20")) * 1e9)


def _round_up(x, m):
    return (x + m - 1) // m * m


class CudaGradReducer(GradReducer):

    def __init__(self, param_groups, world_size, rank, should_sync,
                 bucket_cap_mb=25, process_group=None, name="reducer"):
        self._lib = _native.load()
        self._pg = process_group
        if world_size > MAX_RANKS:
            raise ValueError("at most {} replicas per NVLink domain"
                             .format(MAX_RANKS))
        self._region = None
        self._provider = None
        self._site = 0          # launch ordinal within the current step
        self._steps = 0         # finalize launches so far (== device step_ctr)
        self.engine = None      # DeviceEngine (device-resident estimator)
        self._seg = {}
        self._had_pair = False
        self._pending_events = None
        self._comm = None
        super().__init__(param_groups, world_size, rank, should_sync,
                         bucket_cap_mb, name)

    # ------------------------------------------------------------------
    # storage: one symmetric region holds pad | stats exchange | staging |
    # every arena's gradient buffer
    # ------------------------------------------------------------------

    def _attach(self):
        dev = self.device
        if dev.type != "cuda":
            raise ValueError("CudaGradReducer needs CUDA parameters")
        for arena in self.arenas:
            if arena.dtype not in _DTYPE_CODE:
                raise ValueError("unsupported gradient dtype {}"
                                 .format(arena.dtype))
        torch.cuda.set_device(dev)
        check(self._lib.adl_set_device(dev.index), "adl_set_device")
        G = self.num_groups
        if G > self._lib.adl_max_groups():
            raise ValueError("too many param groups for the fused "
                             "statistics kernels ({} > {})".format(
                                 G, self._lib.adl_max_groups()))
        self._xchg_bytes = _round_up(2 * (4 * G + XCHG_TAIL) * 8, _ALIGN)
        offsets, cursor = {}, 0
        offsets["pad"] = cursor
        cursor += _round_up(_PAD_BYTES, _ALIGN)
        offsets["xchg"] = cursor
        cursor += self._xchg_bytes
        offsets["staging"] = cursor
        cursor += _STAGING_BYTES if self.world_size > 1 else 0
        for i, arena in enumerate(self.arenas):
            offsets[("grad", i)] = cursor
            itemsize = torch.empty((), dtype=arena.dtype).element_size()
            cursor += _round_up(max(arena.total, 1) * itemsize, _ALIGN)
        # NVLS: buckets at least this large go through the switch
        # Measured at N=8 (graph replays, profiles/r2_n8/allreduce_n8.json):
        # the two-shot P2P flavour wins up to 64 MB (16 MB: 66 vs 100 us,
        # 64 MB: 210 vs 217 us per kernel), NVLS with 96 CTAs wins at
        # 256 MB (708 vs 778 us) -- the switch path pays a second pass over
        # the local bucket for the per-replica statistic. So only buckets of
        # at least 128 MB (single huge parameters) take it.
        self._nvls_min_bytes = int(float(os.environ.get(
            "ADAPTDL_B200_NVLS_MIN_MB", "128")) * (1 << 20))
        self._nvls_ctas = max(1, min(int(os.environ.get(
            "ADAPTDL_B200_NVLS_CTAS", "96")), MAX_CTAS - 1))
        # the switch moves (1 + 1/N) B per GPU against 2 (N-1)/N B for the
        # two-shot flavour: no gain at N = 2 (measured 2x slower), 1.55x
        # fewer bytes at N = 8
        self._nvls_min_world = int(os.environ.get(
            "ADAPTDL_B200_NVLS_MIN_WORLD", "4"))
        # one-shot push: buckets whose pushed bytes ((N-1) x bucket) stay
        # below this are latency-bound; each gets N lanes of staging
        self._oneshot_push_bytes = int(float(os.environ.get(
            "ADAPTDL_B200_ONESHOT_KB", "1024")) * 1024)
        self._flavour = {}
        for i, arena in enumerate(self.arenas):
            itemsize = torch.empty((), dtype=arena.dtype).element_size()
            for b in arena.buckets:
                nbytes = b.length * itemsize
                if self.world_size > 1 and \
                        (self.world_size - 1) * nbytes <= \
                        self._oneshot_push_bytes:
                    self._flavour[(i, b.index)] = FLAVOUR_ONESHOT
                    offsets[("stage", i, b.index)] = cursor
                    cursor += _round_up(self.world_size * nbytes, _ALIGN)
                else:
                    self._flavour[(i, b.index)] = FLAVOUR_TWOSHOT
        self._provider = symm.make_provider(
            self._pg if self._pg is not None else
            (dist.group.WORLD if self.world_size > 1 else None),
            dev, self.world_size)
        self._region = self._provider.allocate(cursor)
        self._offsets = offsets
        _, self._pad_ptrs = self._region.carve(offsets["pad"], _PAD_BYTES)
        _, self._xchg_ptrs = self._region.carve(offsets["xchg"],
                                                self._xchg_bytes)
        if self.world_size > 1:
            self._staging, self._staging_ptrs = self._region.carve(
                offsets["staging"], _STAGING_BYTES)
        self._grad_ptrs = {}
        self._grad_mc = {}
        self._stage_ptrs = {}
        for key, off in offsets.items():
            if isinstance(key, tuple) and key[0] == "stage":
                _, ai, bi = key
                arena = self.arenas[ai]
                itemsize = torch.empty((), dtype=arena.dtype).element_size()
                nbytes = self.world_size * arena.buckets[bi].length * itemsize
                _, self._stage_ptrs[(ai, bi)] = self._region.carve(off, nbytes)
        for i, arena in enumerate(self.arenas):
            itemsize = torch.empty((), dtype=arena.dtype).element_size()
            view, ptrs = self._region.carve(
                offsets[("grad", i)], max(arena.total, 1) * itemsize,
                arena.dtype)
            arena._symm_grad = view
            self._grad_ptrs[i] = ptrs
            self._grad_mc[i] = (self._region.mc_ptr + offsets[("grad", i)]
                                if self._region.mc_ptr else 0)
            for b in arena.buckets:
                if self._grad_mc[i] and \
                        self.world_size >= self._nvls_min_world and \
                        self._flavour[(i, b.index)] == FLAVOUR_TWOSHOT and \
                        b.length * itemsize >= self._nvls_min_bytes:
                    self._flavour[(i, b.index)] = FLAVOUR_NVLS
        # statistics, error word, timers, mailbox
        self._stats = torch.zeros(4, G, dtype=torch.float64, device=dev)
        self._result = torch.zeros(4, G, dtype=torch.float64, device=dev)
        self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        self._t_start = torch.zeros(1, dtype=torch.int64, device=dev)
        self._step_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        self._pair_state = torch.zeros(1, dtype=torch.int32, device=dev)
        # step clock: %globaltimer of the previous step mark + the running
        # accumulation-step totals (adl_kernels.cu, finalize_body)
        self._last_stamp = torch.zeros(1, dtype=torch.int64, device=dev)
        self._clock = torch.zeros(CLOCK_DOUBLES, dtype=torch.float64,
                                  device=dev)
        self._ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        self._amp_scale = None      # device float tensor (GradScaler._scale)
        self._fin_fused = False
        self.fuse_finalize = os.environ.get(
            "ADAPTDL_B200_FUSE_FINALIZE", "1") != "0"
        self.oneshot_launches = 0
        self.nvls_launches = 0
        self._ring = 16
        self._slot = MBOX_HDR + 4 * G
        self._mailbox = torch.zeros(self._ring * self._slot,
                                    dtype=torch.float64).pin_memory()
        self._comm = torch.cuda.Stream(dev, priority=-1)
        self._sm_count = max(self._lib.adl_sm_count(dev.index), 1)
        self._reduce_ctas = int(os.environ.get(
            "ADAPTDL_B200_REDUCE_CTAS", "32"))
        self._reduce_ctas = max(1, min(self._reduce_ctas, MAX_CTAS - 1))
        # CTA size of the bucket kernels that run next to backward (two-shot
        # and one-shot flavours): 256 threads keep half of an SM's registers
        # and warp slots free for the backward kernels; NVLS-sized buckets
        # are bandwidth-bound and keep 512
        self._reduce_threads = int(os.environ.get(
            "ADAPTDL_B200_REDUCE_THREADS", "256"))
        self._local_ctas = max(1, int(os.environ.get(
            "ADAPTDL_B200_LOCAL_CTAS", "64")))
        super()._attach()
        for i, arena in enumerate(self.arenas):
            self._build_seg_tables(i, arena)

    def _alloc_flat(self, arena, kind):
        if kind == "grad":
            return arena._symm_grad
        if kind == "pinv":
            return torch.ones(max(arena.total, 1), dtype=arena.dtype,
                              device=self.device)
        return torch.zeros(max(arena.total, 1), dtype=arena.dtype,
                           device=self.device)

    def _build_seg_tables(self, arena_idx, arena):
        vec = layout.VEC_BYTES // torch.empty(
            (), dtype=arena.dtype).element_size()
        for b in arena.buckets:
            rows = sorted(layout.segment_table(b, vec))
            n_vec = b.length // vec
            ends = [r[1] for r in rows]
            groups = [r[2] for r in rows]
            # every vector must resolve to a segment: stretch each end to
            # the next start (padding is zero) and the last to the bucket end
            for j in range(len(rows) - 1):
                ends[j] = max(ends[j], rows[j + 1][0])
            ends[-1] = n_vec
            self._seg[(arena_idx, b.index)] = (
                torch.tensor(ends, dtype=torch.int32, device=self.device),
                torch.tensor(groups, dtype=torch.int32, device=self.device),
                n_vec, vec)

    # ------------------------------------------------------------------

    def _arena_index(self, arena):
        for i, a in enumerate(self.arenas):
            if a is arena:
                return i
        raise KeyError

    def _itemsize(self, arena):
        return layout.VEC_BYTES // self._seg[
            (self._arena_index(arena), arena.buckets[0].index)][3]

    def _order_after_compute(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._comm.wait_event(ev)

    def _local_grid(self, n_vec):
        """Grid of a kernel that has the GPU to itself (stand-alone use)."""
        return max(1, min(2 * self._sm_count,
                          (n_vec + 2 * 512 - 1) // (2 * 512)))

    def _thin_grid(self, n_vec):
        """``(CTAs, threads)`` of a local pass that runs NEXT TO backward on
        the high-priority comm stream: a fraction of the SMs' thread and
        register budget (a full-width grid stalls the backward kernels for
        the ~10 us of launch + latency every such kernel has, measured as
        +10-15 us of step time per bucket), four vectors per tensor in
        flight per thread for bandwidth."""
        threads = self._reduce_threads
        per_cta = threads * 8
        return max(1, min(self._local_ctas,
                          (n_vec + per_cta - 1) // per_cta)), threads

    def _local_args(self, arena, bucket, mode):
        ai = self._arena_index(arena)
        ends, groups, n_vec, vec = self._seg[(ai, bucket.index)]
        itemsize = layout.VEC_BYTES // vec
        off = bucket.start * itemsize
        args = LocalArgs()
        args.g = arena.grad.data_ptr() + off
        if mode in (0, 1):
            args.a = self._ensure(arena, "acc").data_ptr() + off
        if mode == 2:
            args.pv = self._ensure(arena, "prev").data_ptr() + off
        self._set_pinv(args, arena, ai, bucket.start, itemsize)
        args.n_vec = n_vec
        args.segs.seg_end = ends.data_ptr()
        args.segs.seg_group = groups.data_ptr()
        args.segs.n_seg = ends.numel()
        args.n_groups = self.num_groups
        return args, n_vec

    def _set_pinv(self, args, arena, ai, start_elem, itemsize):
        """Preconditioner of the squared norms: the device engine's Adam
        second moments (read in place, no copies), or the host path's flat
        divisor arena, or none."""
        engine = self.engine
        if engine is not None and engine.enabled and \
                engine.precondition_stats:
            moments, wide = engine.second_moments(ai)
            args.pinv = moments.data_ptr() + \
                start_elem * moments.element_size()
            args.pinv_mode = PINV_ADAM
            args.pinv_wide = int(wide)
            args.pinv_coef = engine.pinv_coef.data_ptr()
        elif self._precond_fn is not None and arena.pinv is not None:
            args.pinv = arena.pinv.data_ptr() + start_elem * itemsize
            args.pinv_mode = PINV_FLAT
        else:
            args.pinv = None

    def _row(self, r):
        return self._stats.data_ptr() + r * self.num_groups * 8

    # -- primitives ------------------------------------------------------

    def _reset_partials(self):
        if getattr(self, "_stats", None) is not None:
            self._stats.zero_()

    def _on_begin_backward(self):
        self._had_pair = False

    def _device_preconditioner(self):
        engine = self.engine
        return engine is not None and engine.enabled and \
            engine.precondition_stats

    def _mark_sync_start(self):
        check(self._lib.adl_stamp(
            self._t_start.data_ptr(),
            torch.cuda.current_stream(self.device).cuda_stream), "adl_stamp")
        self.launches += 1

    def _mark_accum_step(self):
        # after this micro-step's folds, on the communication stream
        self._order_after_compute()
        check(self._lib.adl_step_mark(
            self._last_stamp.data_ptr(), self._clock.data_ptr(),
            self._comm.cuda_stream), "adl_step_mark")
        self.launches += 1

    def reset_step_clock(self):
        """Forget the previous step mark: the next step's interval is not
        measured (start of a training loop, after evaluation / a
        checkpoint)."""
        with torch.cuda.stream(self._comm):
            self._last_stamp.zero_()
            self._clock.zero_()

    def set_amp_scale(self, scale_tensor):
        """Device tensor holding the AMP loss scale the gradients carry
        (``GradScaler._scale``); the device estimator divides it out."""
        self._amp_scale = scale_tensor

    def _launch_local(self, arena, bucket, mode, flag=0, last=False):
        args, n_vec = self._local_args(arena, bucket, mode)
        args.s0 = self._row(1) if mode == 2 else self._row(0)
        args.s1 = self._row(2)
        args.s2 = self._row(3)
        args.flag = flag
        if mode == 2 and self.engine is not None and self.engine.enabled:
            args.flag_ptr = self._pair_state.data_ptr()
        fin = None
        if last and mode == 2 and self.fuse_finalize:
            if flag:
                self._had_pair = True
            fin = self._finalize_args()
            args.fuse_fin = 1
            args.ticket = self._ticket.data_ptr()
            self._fin_fused = True
        self._order_after_compute()
        grid, threads = self._thin_grid(n_vec)
        check(self._lib.adl_local(
            ctypes.byref(args), ctypes.byref(fin) if fin is not None else None,
            mode, _DTYPE_CODE[arena.dtype], grid, threads,
            self._comm.cuda_stream), "adl_local")
        self.launches += 1

    def _fold_acc(self, arena, bucket):
        self._launch_local(arena, bucket, 0)

    def _fold_final(self, arena, bucket):
        self._launch_local(arena, bucket, 1)

    def _pair(self, arena, bucket, last=False):
        self._launch_local(arena, bucket, 2, flag=int(self._prev_valid),
                           last=last)
        self._had_pair = self._had_pair or self._prev_valid

    def _reduce(self, arena, bucket, scale, want_local, last=False):
        ai = self._arena_index(arena)
        ends, groups, n_vec, vec = self._seg[(ai, bucket.index)]
        itemsize = layout.VEC_BYTES // vec
        off = bucket.start * itemsize
        args = ReduceArgs()
        for p in range(self.world_size):
            args.buf[p] = self._grad_ptrs[ai][p] + off
            args.pad[p] = self._pad_ptrs[p]
        args.rank, args.world = self.rank, self.world_size
        args.step_ctr = self._step_ctr.data_ptr()
        args.site = self._next_site()
        args.n_vec = n_vec
        args.scale = scale
        args.want_local = int(want_local)
        args.segs.seg_end = ends.data_ptr()
        args.segs.seg_group = groups.data_ptr()
        args.segs.n_seg = ends.numel()
        args.n_groups = self.num_groups
        self._set_pinv(args, arena, ai, bucket.start, itemsize)
        args.L = self._row(0)
        args.T = self._row(1)
        args.err = self._err.data_ptr()
        args.timeout_ns = _TIMEOUT_NS
        slice_vec = n_vec // self.world_size
        flavour = self._flavour.get((ai, bucket.index), FLAVOUR_TWOSHOT) \
            if self.world_size > 1 else FLAVOUR_TWOSHOT
        if flavour == FLAVOUR_NVLS:
            # in-switch reduction: multimem.ld_reduce / multimem.st
            args.mc_buf = self._grad_mc[ai] + off
            self.nvls_launches += 1
            threads = 512
            per_cta = threads * 8
            grid = max(1, min(self._nvls_ctas,
                              (slice_vec + per_cta - 1) // per_cta))
        elif flavour == FLAVOUR_ONESHOT:
            for p in range(self.world_size):
                args.stage[p] = self._stage_ptrs[(ai, bucket.index)][p]
            self.oneshot_launches += 1
            threads = self._reduce_threads
            grid = max(1, min(self._reduce_ctas,
                              (n_vec + threads - 1) // threads))
        elif self.world_size > 1:
            # each thread keeps 16/W vectors in flight per iteration. Within
            # the CTA cap, prefer more CTAs over more iterations: an
            # iteration is a full NVLink round trip (~3 us), so a 1 MB bucket
            # on 2 CTAs spent 8 round trips where 8 CTAs need 2 (measured at
            # N=8, profiles/r2_n8/allreduce_n8.json: 40 us per 1 MB kernel)
            threads = self._reduce_threads
            per_iter = threads * max(16 // self.world_size, 1)
            grid = max(1, min(self._reduce_ctas,
                              (slice_vec + 2 * per_iter - 1)
                              // (2 * per_iter)))
        else:
            threads = 512
            grid = self._local_grid(n_vec)
        fin = None
        if last and self.fuse_finalize:
            fin = self._finalize_args()
            args.fuse_fin = 1
            args.ticket = self._ticket.data_ptr()
            self._fin_fused = True
        self._order_after_compute()
        check(self._lib.adl_allreduce_gns(
            ctypes.byref(args), ctypes.byref(fin) if fin is not None else None,
            _DTYPE_CODE[arena.dtype], flavour, grid, threads,
            self._comm.cuda_stream), "adl_allreduce_gns")
        self.launches += 1

    def _next_site(self):
        self._site += 1
        if self._site >= SITES_PER_STEP:
            raise RuntimeError("too many fused launches in one step")
        return self._site

    def _finalize_args(self):
        """Arguments of this step's finalize (stand-alone launch, or fused
        into the last bucket's kernel)."""
        pair_mode = self.world_size == 1 and self._k_before == 0
        engine = self.engine if (self.engine is not None
                                 and self.engine.enabled) else None
        if engine is not None:
            n_rows = 4 if pair_mode else 2
        else:
            n_rows = 4 if (pair_mode and self._had_pair) else 2
        self._fin_rows = n_rows
        G = self.num_groups
        args = FinalizeArgs()
        for p in range(self.world_size):
            args.xchg[p] = self._xchg_ptrs[p]
            args.pad[p] = self._pad_ptrs[p]
        args.rank, args.world = self.rank, self.world_size
        args.step_ctr = self._step_ctr.data_ptr()
        args.site = self._next_site()
        args.n_rows = n_rows
        args.n_groups = G
        for r in range(4):
            args.rows[r] = self._row(r)
        args.sum_mask = 0b0011 if self.world_size > 1 else 0
        args.micro_steps = self._k_before + 1
        args.pair_mode = int(pair_mode)
        args.pair_flag = int(self._had_pair)
        args.pair_state = self._pair_state.data_ptr()
        args.mailbox = self._mailbox.data_ptr()
        args.ring, args.slot_doubles = self._ring, self._slot
        args.result = self._result.data_ptr()
        args.t_start = self._t_start.data_ptr()
        if engine is not None:
            args.gns_state = engine.state.data_ptr()
            args.gns_ctrl = engine.ctrl.data_ptr()
            args.lr_factor = engine.lr_factor.data_ptr()
        args.err = self._err.data_ptr()
        args.timeout_ns = _TIMEOUT_NS
        args.last_stamp = self._last_stamp.data_ptr()
        args.clock = self._clock.data_ptr()
        if self._amp_scale is not None:
            args.amp_scale = self._amp_scale.data_ptr()
        return args

    def _finalize_step(self):
        if self._fin_fused:
            # the last bucket's kernel already carried the finalize
            self._fin_fused = False
            n_rows = self._fin_rows
        else:
            args = self._finalize_args()
            n_rows = self._fin_rows
            self._order_after_compute()
            check(self._lib.adl_finalize_stats(ctypes.byref(args),
                                               self._comm.cuda_stream),
                  "adl_finalize_stats")
            self.launches += 1
        if self.world_size == 1 and self._k_before == 0:
            self._prev_valid = True
        done = torch.cuda.Event()
        done.record(self._comm)
        # the optimizer (compute stream) must see the reduced gradients
        torch.cuda.current_stream(self.device).wait_event(done)
        count = self.world_size * self._accum_count
        handle = self.note_step_finalized(done, count, n_rows)
        return handle

    def note_step_finalized(self, done, count, n_rows):
        """Host bookkeeping for one finalize launch (also used when a
        captured CUDA graph containing the launch is replayed)."""
        step = self._steps
        self._steps += 1
        self._site = 0
        return (step, done, count, n_rows)

    def replay_bookkeeping(self, sync, k_before):
        """Host state after a captured step (one backward pass) has been
        replayed: mirrors what the autograd hooks record in eager mode."""
        self._sync = bool(sync)
        self._k_before = int(k_before)
        self._accum_count = self._k_before + 1
        if sync:
            pair_mode = self.world_size == 1 and self._k_before == 0
            self._stats_ready = self.note_step_finalized(
                None, self.world_size * self._accum_count,
                4 if pair_mode else 2)

    def read_slot(self, step, wait_event=None, spin=True):
        """Mailbox slot of optimizer step ``step`` (0-based) as a numpy
        view; waits until the device has published it."""
        if wait_event is not None:
            wait_event.synchronize()
        lo = (step % self._ring) * self._slot
        arr = self._mailbox.numpy()[lo:lo + self._slot]
        if spin:
            import time
            t0 = time.time()
            while int(arr[0]) != step + 1:
                if int(arr[0]) > step + 1:
                    raise RuntimeError(
                        "statistics mailbox overrun: wanted step {} but the "
                        "slot holds {}".format(step + 1, int(arr[0])))
                if time.time() - t0 > _TIMEOUT_NS * 1e-9:
                    raise RuntimeError("timed out waiting for the "
                                       "statistics of step {}".format(step))
                time.sleep(0)
        return arr

    def peek_slot(self, step):
        """Non-blocking :meth:`read_slot`: the slot's numpy view if the
        device has published optimizer step ``step``, ``None`` if not yet,
        ``False`` if the ring has already moved past it."""
        lo = (step % self._ring) * self._slot
        arr = self._mailbox.numpy()[lo:lo + self._slot]
        seq = int(arr[0])
        if seq == step + 1:
            return arr
        return False if seq > step + 1 else None

    def _resolve_stats(self, handle):
        step, done, count, n_rows = handle
        arr = self.read_slot(step, wait_event=done)
        G = self.num_groups
        if int(arr[MB_ERR]) != 0:
            raise RuntimeError(
                "fused all-reduce timed out waiting for a peer "
                "(error word {})".format(int(arr[MB_ERR])))
        n = n_rows * G
        rows = np.array(arr[MBOX_HDR:MBOX_HDR + n],
                        dtype=np.float64).reshape(n_rows, G)
        pair = (rows[2], rows[3]) if n_rows == 4 else None
        return GradStats(rows[0], rows[1], count, pair,
                         sync_time=float(arr[MB_SYNC_NS]) * 1e-9)

    # -- broadcast -----------------------------------------------------------

    @staticmethod
    def _memory_bytes(t):
        """The tensor's bytes in memory order (layouts are identical on all
        ranks, so broadcasting memory order preserves e.g. channels_last)."""
        t = t.detach()
        if t.is_contiguous():
            flat = t.reshape(-1)
        else:
            from adaptdl_b200.parallel.reducer_base import _is_dense
            if not _is_dense(t):
                raise ValueError("broadcast needs dense tensors")
            flat = t.as_strided((t.numel(),), (1,))
        return flat.view(torch.uint8)

    def broadcast_parameters(self, tensors, src=0):
        if self.world_size <= 1:
            return
        tensors = [self._memory_bytes(t) for t in tensors
                   if t is not None and t.numel() > 0]
        stream = torch.cuda.current_stream(self.device)
        chunk, chunk_bytes = [], 0

        def flush():
            nonlocal chunk, chunk_bytes
            if not chunk:
                return
            n = sum(t.numel() for t in chunk)
            n_pad = _round_up(n, 16)
            if self.rank == src:
                torch.cat(chunk, out=self._staging[:n])
            args = BcastArgs()
            for p in range(self.world_size):
                args.staging[p] = self._staging_ptrs[p]
                args.pad[p] = self._pad_ptrs[p]
            args.rank, args.world, args.src = self.rank, self.world_size, src
            args.step_ctr = self._step_ctr.data_ptr()
            args.site = self._next_site()
            args.dst = self._staging_ptrs[self.rank]
            args.n_vec = n_pad // 16
            args.err = self._err.data_ptr()
            args.timeout_ns = _TIMEOUT_NS
            grid = max(1, min(16, (n_pad // 16 + 1023) // 1024))
            check(self._lib.adl_bcast_pull(ctypes.byref(args), grid,
                                           stream.cuda_stream),
                  "adl_bcast_pull")
            self.launches += 1
            if self.rank != src:
                cursor, pieces = 0, []
                for t in chunk:
                    nb = t.numel()
                    pieces.append(self._staging[cursor:cursor + nb])
                    cursor += nb
                # one multi-tensor launch instead of one copy per tensor
                torch._foreach_copy_(chunk, pieces)
            chunk, chunk_bytes = [], 0

        for t in tensors:
            nb = t.numel()
            if nb > _STAGING_BYTES:
                flush()
                for lo in range(0, nb, _STAGING_BYTES):
                    piece = t[lo:lo + _STAGING_BYTES]
                    chunk, chunk_bytes = [piece], piece.numel()
                    flush()
                continue
            if chunk_bytes + nb > _STAGING_BYTES:
                flush()
            chunk.append(t)
            chunk_bytes += nb
        flush()
