"""Where does a rescale spend its time?

With ``ADAPTDL_B200_RESCALE_TRACE=<directory>`` every replica appends one
JSON line per life-cycle event to ``<directory>/trace-<generation>-<rank>
.jsonl``: interpreter up, process group ready, data-parallel wrapper built
(peer mappings opened, state restored), first optimizer step done, signal
received, exit consensus, checkpoint written, exit. The single-box launcher
(``adaptdl_b200.sched.local``) merges them into its report, so a
2 -> 4 -> 8 -> 4 run comes with the breakdown signal -> consensus ->
checkpoint -> process start -> rendezvous -> remap -> first step
(BASELINE config 4; the reference restarts the same way,
``adaptdl/adaptdl/torch/data.py:321-328`` + ``checkpoint.py:106-133``, but
records none of it).

Off (the default) a mark is one dictionary lookup.
"""

import json
import os
import time

_DIR = os.environ.get("ADAPTDL_B200_RESCALE_TRACE")
_SEEN = set()
_SUSPENDED = False


def enabled():
    return bool(_DIR)


def suspend():
    """Stop recording (a warm standby interpreter is not a replica yet)."""
    global _SUSPENDED
    _SUSPENDED = True


def resume():
    """Start the life cycle afresh: the process has just become a replica."""
    global _SUSPENDED
    _SUSPENDED = False
    _SEEN.clear()


def mark(event, once=True, **fields):
    """Record ``event`` now (wall clock, comparable across processes of one
    host). ``once``: only the first occurrence per process is kept."""
    if not _DIR or _SUSPENDED:
        return
    if once:
        if event in _SEEN:
            return
        _SEEN.add(event)
    row = {"event": event, "t": time.time(),
           "rank": int(os.environ.get("ADAPTDL_REPLICA_RANK", "0") or 0),
           "replicas": int(os.environ.get("ADAPTDL_NUM_REPLICAS", "1") or 1),
           "generation": int(os.environ.get("ADAPTDL_NUM_RESTARTS", "0")
                             or 0)}
    row.update(fields)
    try:
        os.makedirs(_DIR, exist_ok=True)
        path = os.path.join(_DIR, "trace-{}-{}.jsonl".format(
            row["generation"], row["rank"]))
        with open(path, "a") as f:
            f.write(json.dumps(row) + "\n")
    except OSError:
        pass


def collect(directory):
    """All rows under ``directory`` sorted by time (launcher side)."""
    rows = []
    if not directory or not os.path.isdir(directory):
        return rows
    for name in sorted(os.listdir(directory)):
        if not name.startswith("trace-"):
            continue
        with open(os.path.join(directory, name)) as f:
            for line in f:
                line = line.strip()
                if line:
                    rows.append(json.loads(line))
    rows.sort(key=lambda r: r["t"])
    return rows


def summarize(rows):
    """Per generation: seconds between consecutive life-cycle events, taking
    for every event the LAST replica to reach it (the job moves at the pace
    of its slowest member)."""
    order = ["interpreter_up", "process_group_ready", "wrapper_ready",
             "first_step_done", "signal_received", "exit_consensus",
             "checkpoint_written", "exiting"]
    by_gen = {}
    for row in rows:
        gen = by_gen.setdefault(row["generation"], {})
        prev = gen.get(row["event"])
        if prev is None or row["t"] > prev:
            gen[row["event"]] = row["t"]
    out = {}
    for generation, events in sorted(by_gen.items()):
        phases = {}
        last_name, last_t = None, None
        for name in order:
            if name not in events:
                continue
            if last_t is not None:
                phases["{}->{}".format(last_name, name)] = round(
                    events[name] - last_t, 3)
            last_name, last_t = name, events[name]
        out[generation] = phases
    return out
