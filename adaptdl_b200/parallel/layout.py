"""Flat gradient layout: arenas -> buckets -> segments.

All gradients of one dtype live in one flat *arena* (``param.grad`` tensors
are views into it). An arena is cut into *buckets* -- the unit of one fused
all-reduce launch -- in reverse parameter order, so buckets complete roughly
in the order backward produces gradients. Inside a bucket each parameter is a
*segment* ``(start, length, group)``; ``group`` is the optimizer param-group
index used for per-group gradient-noise-scale statistics.

Alignment rules (what the sm_100a kernels rely on):

* every segment starts on a 16-byte boundary (so a 128-bit vector never
  straddles two statistic groups); padding is zero and never contributes to a
  norm;
* every bucket starts on a ``BUCKET_ALIGN_BYTES`` boundary and its length is
  a multiple of ``16 B x world_size`` so the two-shot kernel's rank slices
  are whole vectors.

Pure Python / no torch dependency: unit-testable on its own.
"""

import collections
import math

VEC_BYTES = 16
BUCKET_ALIGN_BYTES = 512

Segment = collections.namedtuple(
    "Segment", ["param_index", "start", "length", "group"])
# start/length in ELEMENTS relative to the arena start.

Bucket = collections.namedtuple(
    "Bucket", ["index", "start", "length", "segments"])
# start/length in elements relative to the arena start; length is padded.


def _round_up(x, m):
    return (x + m - 1) // m * m


def plan_arena(numels, groups, itemsize, bucket_cap_bytes,
               first_bucket_cap_bytes=None, world_size=1,
               last_bucket_cap_bytes=None):
    """Plan one arena.

    Arguments:
        numels: number of elements of each parameter, in *registration*
            order (the plan walks them in reverse).
        groups: optimizer param-group index of each parameter.
        itemsize: bytes per element.
        bucket_cap_bytes: soft cap of a bucket's payload.
        first_bucket_cap_bytes: cap of the first bucket to be produced by
            backward (small, so communication starts early); defaults to
            ``bucket_cap_bytes``.
        world_size: replicas sharing each bucket (slice alignment).
        last_bucket_cap_bytes: if given, the parameters backward reaches
            last (the first ones registered) are split off into a final
            bucket of at most this payload (at least one parameter): that
            bucket's reduction cannot overlap with backward, so it is kept
            latency-sized.

    Returns ``(total_elements, buckets)``; ``buckets[i].segments`` are in
    arena order, and ``buckets`` is in expected completion order.
    """
    assert len(numels) == len(groups)
    vec = VEC_BYTES // itemsize
    # a common multiple of both requirements (for world sizes that are not
    # powers of two -- an elastic job may run on 3, 5, 6, 7 GPUs -- rounding
    # one up to the other would break the 512-byte starts)
    bucket_align = math.lcm(max(BUCKET_ALIGN_BYTES // itemsize, 1),
                            vec * world_size)
    if first_bucket_cap_bytes is None:
        first_bucket_cap_bytes = bucket_cap_bytes
    first_bucket_cap_bytes = min(first_bucket_cap_bytes, bucket_cap_bytes)

    # index of the first parameter (in walk order = reverse registration)
    # of the tail bucket: the longest run of first-registered parameters
    # whose payload fits the cap
    tail_from = -1
    if last_bucket_cap_bytes is not None and len(numels) > 1:
        payload = 0
        for pidx in range(len(numels)):
            payload += numels[pidx] * itemsize
            if payload > last_bucket_cap_bytes and pidx > 0:
                break
            tail_from = pidx
        if tail_from == len(numels) - 1:
            tail_from = -1      # everything fits: no separate tail

    buckets = []
    cursor = 0            # arena cursor in elements
    cur_segments = []
    cur_start = 0
    cur_payload = 0

    def close_bucket():
        nonlocal cursor, cur_segments, cur_start, cur_payload
        if not cur_segments:
            return
        end = _round_up(cursor, bucket_align)
        buckets.append(Bucket(len(buckets), cur_start, end - cur_start,
                              tuple(cur_segments)))
        cursor = end
        cur_segments = []
        cur_start = cursor
        cur_payload = 0

    for pidx in reversed(range(len(numels))):
        n = numels[pidx]
        cap = first_bucket_cap_bytes if not buckets else bucket_cap_bytes
        if cur_segments and ((cur_payload + n * itemsize) > cap
                             or pidx == tail_from):
            close_bucket()
        start = _round_up(cursor, vec)
        cur_segments.append(Segment(pidx, start, n, groups[pidx]))
        cursor = start + n
        cur_payload += n * itemsize
    close_bucket()
    return cursor, buckets


def segment_table(bucket, vec_elems):
    """Rows ``(start_vec, end_vec, group)`` relative to the bucket start, in
    units of 16-byte vectors, for the kernels' statistic-group lookup. Gaps
    (padding) between rows belong to no group."""
    rows = []
    for seg in bucket.segments:
        rel = seg.start - bucket.start
        assert rel % vec_elems == 0
        start_vec = rel // vec_elems
        end_vec = (rel + seg.length + vec_elems - 1) // vec_elems
        rows.append((start_vec, end_vec, seg.group))
    return rows
