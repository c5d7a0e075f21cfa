#!/usr/bin/env python
"""Explicit-state model of the gradient path's synchronisation protocol
(``csrc/adl_kernels.cu``): exhaustive interleavings for small configurations,
random schedules with skewed ranks for larger ones, and the same searches on
deliberately broken variants of the protocol (which must be caught).

    python tools/protocol_model.py                 # the checks the test runs
    python tools/protocol_model.py --ranks 3 --ctas 2 --steps 4 --random 20000

What is modelled, per rank (one process per GPU), for every optimizer step:

* the compute stream produces the buckets' gradients in backward order
  (``produce``), later runs the optimizer (which reads every reduced bucket),
  then the next step's backward overwrites the gradient arena;
* the communication stream runs one kernel per bucket, in order, each as
  ``ctas`` concurrent CTAs; kernel k + 1 starts when every CTA of kernel k has
  finished and its bucket has been produced;
* **two-shot** bucket kernel (``allreduce_gns_kernel``): CTA c raises its flag
  on every peer's pad (value = launch epoch: step counter x sites-per-step +
  launch ordinal, read from the rank's own device counter), waits for CTA c of
  every peer, reads its part of slice ``rank`` from every arena, writes the sum
  back into every arena;
* **one-shot** kernel (``allreduce_oneshot_kernel``): CTA c first pushes its
  stripe of the own bucket into lane ``rank`` of every peer's staging area,
  then the same flag round, then sums the lanes locally and overwrites the own
  bucket;
* **NVLS** kernel (``allreduce_nvls_kernel``): for every foreign slice the CTA
  first reads its own copy (the per-replica statistic) and only then raises
  its flag on the slice's owner; after the flags of all peers it reads the
  own slice through the switch (``multimem.ld_reduce``: every rank's copy) and
  writes the result into every arena (``multimem.st``);
* every kernel ends with the CTA ticket; the last CTA of the step's LAST
  bucket kernel runs the **finalize**: writes this rank's statistics record
  into the exchange buffer of the step's parity, one more flag round on the
  reserved pad slot, reads every peer's record, bumps the step counter;
* the optimizer waits for the finalize (stream order).

Flags are never reset and compared as ``flag >= epoch``; the exchange record
is double-buffered by step parity; nothing else synchronises the ranks -- as
in the kernels. Memory is sequentially consistent in the model (the kernels
use ``fence.sys`` + ``st.release.sys`` / ``ld.acquire.sys`` where the model
relies on order), so this checks the PROTOCOL, not the memory-model
annotations.

A violation is a read that does not see the value the algorithm needs: a peer
gradient that is not this step's local gradient (read too early, or already
overwritten by the next backward), a staging lane of another step, a peer
record of another step, or an optimizer reading a bucket that is not fully
reduced.
"""
import argparse
import random
import sys

SITES = 1024            # ADL_SITES_PER_STEP


class Violation(Exception):
    pass


class Config(object):
    def __init__(self, ranks=2, ctas=1, steps=2, buckets=("two", "one"),
                 no_finalize_barrier=False, single_xchg=False,
                 same_epoch=False, no_start_barrier=False,
                 push_after_barrier=False, nvls_flag_first=False):
        self.ranks, self.ctas, self.steps = ranks, ctas, steps
        self.buckets = tuple(buckets)          # in backward (launch) order
        # broken variants
        self.no_finalize_barrier = no_finalize_barrier
        self.single_xchg = single_xchg
        self.same_epoch = same_epoch
        self.no_start_barrier = no_start_barrier
        self.push_after_barrier = push_after_barrier
        self.nvls_flag_first = nvls_flag_first


def build_programs(cfg):
    """``{thread id: [op, ...]}``; ops are tuples interpreted by ``step``."""
    R, C, B = cfg.ranks, cfg.ctas, len(cfg.buckets)
    programs = {}
    for r in range(R):
        comp = []
        for s in range(cfg.steps):
            for b in range(B):
                comp.append(("produce", b, s))
            comp.append(("await_finalize", s))
            comp.append(("optimizer", s))
        programs[("comp", r)] = comp
        for c in range(C):
            ops = []
            for s in range(cfg.steps):
                for b, flavour in enumerate(cfg.buckets):
                    k = s * B + b                  # kernel ordinal on the stream
                    site = 0 if cfg.same_epoch else b + 1
                    ops.append(("launch", k, b, s))
                    if flavour == "nvls":
                        # per foreign slice: this rank's own copy feeds the
                        # per-replica statistic BEFORE the owner is told it
                        # may overwrite it (multimem.st lands in every arena)
                        for d in range(1, R):
                            q = (r + d) % R
                            if cfg.nvls_flag_first:
                                ops.append(("flag", q, ("cta", c), site))
                                ops.append(("read_own_copy", q, b, s))
                            else:
                                ops.append(("read_own_copy", q, b, s))
                                ops.append(("flag", q, ("cta", c), site))
                        ops.append(("wait_peers", ("cta", c), site))
                        for q in range(R):          # multimem.ld_reduce
                            ops.append(("read_grad", q, b, s))
                        for q in range(R):          # multimem.st
                            ops.append(("write_sum", q, b, s))
                        ops.append(("ticket", b == B - 1))
                        if b == B - 1:
                            fin_site = 0 if cfg.same_epoch else B + 1
                            ops.append(("fin_record", s))
                            if not cfg.no_finalize_barrier:
                                for q in range(R):
                                    ops.append(("fin_flag", q, fin_site))
                                ops.append(("fin_wait", fin_site))
                            for q in range(R):
                                ops.append(("fin_read", q, s))
                            ops.append(("fin_publish", s))
                        ops.append(("done", k))
                        continue
                    if flavour == "one" and not cfg.push_after_barrier:
                        for q in range(R):
                            if q != r:
                                ops.append(("push", q, b, s))
                    if not cfg.no_start_barrier:
                        for q in range(R):
                            ops.append(("flag", q, ("cta", c), site))
                        ops.append(("wait", ("cta", c), site))
                    if flavour == "one" and cfg.push_after_barrier:
                        for q in range(R):
                            if q != r:
                                ops.append(("push", q, b, s))
                    if flavour == "two":
                        for q in range(R):
                            ops.append(("read_grad", q, b, s))
                        for q in range(R):
                            ops.append(("write_sum", q, b, s))
                    else:
                        for q in range(R):
                            if q != r:
                                ops.append(("read_lane", q, b, s))
                        ops.append(("read_grad", r, b, s))
                        ops.append(("write_own", b, s))
                    ops.append(("ticket", b == B - 1))
                    if b == B - 1:                 # fused finalize, last CTA only
                        fin_site = 0 if cfg.same_epoch else B + 1
                        ops.append(("fin_record", s))
                        if not cfg.no_finalize_barrier:
                            for q in range(R):
                                ops.append(("fin_flag", q, fin_site))
                            ops.append(("fin_wait", fin_site))
                        for q in range(R):
                            ops.append(("fin_read", q, s))
                        ops.append(("fin_publish", s))
                    ops.append(("done", k))
            programs[("cta", r, c)] = ops
    return programs


class State(object):
    """pcs + memory; hashable snapshot via ``key()``."""

    def __init__(self, cfg, programs):
        self.cfg = cfg
        self.programs = programs
        self.threads = sorted(programs)
        self.pc = {t: 0 for t in self.threads}
        self.last = {t: False for t in self.threads if t[0] == "cta"}
        self.mem = {}

    def copy(self):
        other = State.__new__(State)
        other.cfg, other.programs, other.threads = \
            self.cfg, self.programs, self.threads
        other.pc = dict(self.pc)
        other.last = dict(self.last)
        other.mem = dict(self.mem)
        return other

    def key(self):
        return (tuple(self.pc[t] for t in self.threads),
                tuple(self.last[t] for t in sorted(self.last)),
                frozenset(self.mem.items()))

    def get(self, loc, default=0):
        return self.mem.get(loc, default)

    def finished(self):
        return all(self.pc[t] >= len(self.programs[t]) for t in self.threads)

    # -- one op -----------------------------------------------------------

    def enabled(self, t):
        """Can thread ``t`` take its next op now?"""
        if self.pc[t] >= len(self.programs[t]):
            return False
        op = self.programs[t][self.pc[t]]
        cfg, kind = self.cfg, op[0]
        r = t[1]
        if kind == "await_finalize":
            return self.get(("step", r)) >= op[1] + 1
        if kind == "launch":
            _, k, b, s = op
            return self.get(("kdone", r)) >= k * cfg.ctas and \
                self.get(("prod", r, b)) >= s + 1
        if kind == "wait":
            epoch = self.get(("step", r)) * SITES + op[2]
            return all(self.get(("pad", r, op[1], q), -1) >= epoch
                       for q in range(cfg.ranks))
        if kind == "wait_peers":
            epoch = self.get(("step", r)) * SITES + op[2]
            return all(self.get(("pad", r, op[1], q), -1) >= epoch
                       for q in range(cfg.ranks) if q != r)
        if kind == "fin_wait":
            if not self.last[t]:
                return True
            epoch = self.get(("step", r)) * SITES + op[1]
            return all(self.get(("pad", r, "fin", q), -1) >= epoch
                       for q in range(cfg.ranks))
        return True

    def step(self, t):
        op = self.programs[t][self.pc[t]]
        self.pc[t] += 1
        cfg, kind = self.cfg, op[0]
        r = t[1]
        c = t[2] if t[0] == "cta" else None
        mem = self.mem
        if kind == "produce":
            _, b, s = op
            for part in self._parts(b):
                mem[("G", r, b, part)] = ("L", r, s)
            mem[("prod", r, b)] = s + 1
        elif kind == "optimizer":
            s = op[1]
            for b in range(len(cfg.buckets)):
                for part in self._parts(b):
                    got = self.get(("G", r, b, part), None)
                    if got != ("R", s):
                        raise Violation(
                            "optimizer of rank {} step {} reads bucket {} "
                            "part {} = {}".format(r, s, b, part, got))
        elif kind in ("launch", "wait", "wait_peers", "await_finalize"):
            pass
        elif kind == "read_own_copy":
            _, q, b, s = op
            got = self.get(("G", r, b, ("two", q, c)), None)
            if got != ("L", r, s):
                raise Violation(
                    "rank {} cta {} step {} reads its own copy of slice {} "
                    "of bucket {} = {}".format(r, c, s, q, b, got))
        elif kind == "flag":
            _, q, slot, site = op
            epoch = self.get(("step", r)) * SITES + site
            loc = ("pad", q, slot, r)
            mem[loc] = max(self.get(loc, -1), epoch)     # flags only grow
        elif kind == "push":
            _, q, b, s = op
            mine = self.get(("G", r, b, ("one", c)), None)
            mem[("stage", q, b, r, c)] = mine
        elif kind == "read_grad":
            _, q, b, s = op
            part = ("one", c) if cfg.buckets[b] == "one" else ("two", r, c)
            got = self.get(("G", q, b, part), None)
            if got != ("L", q, s):
                raise Violation(
                    "rank {} cta {} step {} reads gradient of rank {} bucket "
                    "{} = {}".format(r, c, s, q, b, got))
        elif kind == "write_sum":
            _, q, b, s = op
            mem[("G", q, b, ("two", r, c))] = ("R", s)
        elif kind == "read_lane":
            _, q, b, s = op
            got = self.get(("stage", r, b, q, c), None)
            if got != ("L", q, s):
                raise Violation(
                    "rank {} cta {} step {} reads lane of rank {} bucket {} "
                    "= {}".format(r, c, s, q, b, got))
        elif kind == "write_own":
            _, b, s = op
            mem[("G", r, b, ("one", c))] = ("R", s)
        elif kind == "ticket":
            drawn = self.get(("ticket", r))
            if drawn == cfg.ctas - 1:
                mem[("ticket", r)] = 0
                self.last[t] = bool(op[1])
            else:
                mem[("ticket", r)] = drawn + 1
                self.last[t] = False
        elif kind == "fin_record":
            if self.last[t]:
                parity = 0 if cfg.single_xchg else self.get(("step", r)) & 1
                mem[("xchg", r, parity)] = ("S", op[1])
        elif kind == "fin_flag":
            if self.last[t]:
                _, q, site = op
                epoch = self.get(("step", r)) * SITES + site
                loc = ("pad", q, "fin", r)
                mem[loc] = max(self.get(loc, -1), epoch)
        elif kind == "fin_wait":
            pass
        elif kind == "fin_read":
            if self.last[t]:
                _, q, s = op
                parity = 0 if cfg.single_xchg else self.get(("step", r)) & 1
                got = self.get(("xchg", q, parity), None)
                if got != ("S", s):
                    raise Violation(
                        "finalize of rank {} step {} reads the record of "
                        "rank {} = {}".format(r, s, q, got))
        elif kind == "fin_publish":
            if self.last[t]:
                mem[("step", r)] = op[1] + 1
                self.last[t] = False
        elif kind == "done":
            mem[("kdone", r)] = self.get(("kdone", r)) + 1
        else:
            raise AssertionError(kind)

    def _parts(self, b):
        cfg = self.cfg
        if cfg.buckets[b] in ("two", "nvls"):
            return [("two", q, c) for q in range(cfg.ranks)
                    for c in range(cfg.ctas)]
        return [("one", c) for c in range(cfg.ctas)]


def explore_all(cfg, limit=2000000):
    """Depth-first search over every interleaving (states are memoised).
    Returns the number of distinct states; raises :class:`Violation` or
    ``RuntimeError`` (deadlock / state limit)."""
    programs = build_programs(cfg)
    start = State(cfg, programs)
    seen = {start.key()}
    stack = [start]
    while stack:
        state = stack.pop()
        ready = [t for t in state.threads if state.enabled(t)]
        if not ready:
            if not state.finished():
                raise RuntimeError("deadlock: " + repr(
                    {t: state.pc[t] for t in state.threads}))
            continue
        for t in ready:
            nxt = state.copy()
            nxt.step(t)
            key = nxt.key()
            if key not in seen:
                seen.add(key)
                if len(seen) > limit:
                    raise RuntimeError("state limit reached")
                stack.append(nxt)
    return len(seen)


def explore_random(cfg, runs, seed=0):
    """Random schedules; in every run the ranks get random speeds (one of
    them is usually much slower or faster than the rest: skew is what breaks
    flag protocols)."""
    rng = random.Random(seed)
    programs = build_programs(cfg)
    for _ in range(runs):
        state = State(cfg, programs)
        speed = [rng.choice((0.02, 0.2, 1.0, 1.0, 5.0))
                 for _ in range(cfg.ranks)]
        while True:
            ready = [t for t in state.threads if state.enabled(t)]
            if not ready:
                if not state.finished():
                    raise RuntimeError("deadlock")
                break
            weights = [speed[t[1]] for t in ready]
            state.step(rng.choices(ready, weights)[0])
    return runs


BROKEN = {
    "no finalize barrier": dict(no_finalize_barrier=True),
    "same epoch for every launch of a step": dict(same_epoch=True),
    "no start barrier in the bucket kernels": dict(no_start_barrier=True),
    "one-shot pushes after the flag round": dict(push_after_barrier=True),
    "NVLS tells the owner before reading its own copy": dict(
        nvls_flag_first=True),
}


# Not in the list: a single-buffered exchange record (``single_xchg``). The
# model finds no violation for it -- a rank cannot get from one finalize to the
# next without passing the start barriers of the bucket kernels in between,
# which its peers only reach after their own finalize has read the records.
# The kernels double-buffer by step parity anyway (it costs nothing).


def check_broken(name, runs=3000, **shape):
    """A broken variant must produce a violation (exhaustively for the
    smallest shape, else within ``runs`` random schedules)."""
    if "NVLS" in name:
        shape = dict(shape, buckets=("nvls", "two"))
    cfg = Config(**dict(shape, **BROKEN[name]))
    try:
        explore_random(cfg, runs, seed=1)
    except Violation as exc:
        return str(exc)
    except RuntimeError as exc:
        return "deadlock ({})".format(exc)
    return None


def main():
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("--ranks", type=int, default=2)
    parser.add_argument("--ctas", type=int, default=1)
    parser.add_argument("--steps", type=int, default=2)
    parser.add_argument("--buckets", default="two,one")
    parser.add_argument("--random", type=int, default=0,
                        help="random schedules instead of the full search")
    args = parser.parse_args()
    cfg = Config(args.ranks, args.ctas, args.steps, args.buckets.split(","))
    if args.random:
        print("random schedules without a violation:",
              explore_random(cfg, args.random))
    else:
        print("distinct states, no violation, no deadlock:",
              explore_all(cfg))
    for name in BROKEN:
        found = check_broken(name, ranks=max(args.ranks, 2), ctas=args.ctas,
                             steps=max(args.steps, 3),
                             buckets=args.buckets.split(","))
        print("broken variant [{}]: {}".format(
            name, found or "NOT CAUGHT"))
        if not found:
            sys.exit(1)


if __name__ == "__main__":
    main()
