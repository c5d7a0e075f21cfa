#!/usr/bin/env python
"""Time to resume after a rescale, this framework vs the unmodified reference
(BASELINE config 4: transformer, elastic 2 -> 4 -> 8 -> 4, host side).

    python tools/rescale_bench.py --arms reference,own,own-standby \\
        --schedule 2,4,8,4 --hold 12 --out profiles/r2_elastic/rescale_bench.json

Every arm runs the SAME job (``tools/rescale_worker.py``, written against the
shared ``adaptdl.torch`` API) under the same single-box launcher
(``adaptdl_b200.sched.local.LocalElasticJob``: it only sets the ``ADAPTDL_*``
variables a scheduler pod would get, sends SIGTERM and waits for exit code
143), on this machine's CPUs with gloo:

* ``reference``   the unmodified package from ``baseline/_ref`` (cold start of
                  every generation: the only mode it has)
* ``own``         this framework, cold start
* ``own-standby`` this framework with what it adds for this path: warm standby
                  interpreters and the teardown-free exit
                  (``sched.local --standby --fast-exit``)

Per transition: SIGTERM -> every replica gone (consensus on the exit
iteration, checkpoint, interpreter teardown), start -> first optimizer step
of the next generation finished on every replica (interpreter + imports,
rendezvous, checkpoint load, first step), and their sum; per generation the
tokens/s it trained at. The job writes the marks itself, so they mean the
same thing in all arms. Default: CPU replicas (CUDA context creation,
peer-memory mapping and graph capture are then not in the numbers);
``--device cuda`` gives every replica a GPU.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from adaptdl_b200.sched.local import LocalElasticJob  # noqa: E402

WORKER = os.path.join(ROOT, "tools", "rescale_worker.py")


def arm_env(arm, marks, device="cpu"):
    env = {"OMP_NUM_THREADS": "1", "RESCALE_MARKS": marks,
           "RESCALE_DEVICE": device}
    if device != "cuda":
        env["CUDA_VISIBLE_DEVICES"] = ""
    if arm == "reference":
        env["PYTHONPATH"] = os.pathsep.join(
            [os.path.join(ROOT, "baseline", "_ref"),
             os.path.join(ROOT, "baseline", "shims")])
        # torch >= 2.6 refuses the numpy arrays the reference keeps in
        # optimizer.state["gns"] unless told otherwise
        env["TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD"] = "1"
    else:
        env["PYTHONPATH"] = ROOT
        if arm == "own-standby":
            env["ADAPTDL_B200_FAST_EXIT"] = "1"
    return env


def read_marks(marks):
    rows = []
    for name in sorted(os.listdir(marks)):
        with open(os.path.join(marks, name)) as f:
            rows += [json.loads(line) for line in f if line.strip()]
    return rows


def wait_first_step(marks, generation, replicas, timeout):
    """Wall-clock time at which the LAST replica of ``generation`` finished
    its first optimizer step."""
    deadline = time.time() + timeout
    while time.time() < deadline:
        done = {r["rank"]: r["t"] for r in read_marks(marks)
                if r["generation"] == generation
                and r["event"] == "first_step"}
        if len(done) >= replicas:
            return max(done.values())
        time.sleep(0.05)
    raise RuntimeError("generation {} never reached its first step".format(
        generation))


def hold(job, seconds):
    end = time.time() + seconds
    while time.time() < end:
        job.maintain()               # tops the standby pool up again
        if job.poll() is not None:
            raise RuntimeError("the job ended during a hold")
        time.sleep(0.1)


def run_arm(arm, schedule, hold_s, pool_warmup, timeout, device="cpu"):
    marks = tempfile.mkdtemp(prefix="rescale-marks-")
    ckpt = tempfile.mkdtemp(prefix="rescale-ckpt-")
    job = LocalElasticJob([sys.executable, WORKER], max(schedule),
                          checkpoint_dir=ckpt,
                          env=arm_env(arm, marks, device),
                          standby=(arm == "own-standby"))
    transitions, generations = [], []
    try:
        if job.standby:
            job.fill_pool()
            time.sleep(pool_warmup)
        t_start = time.time()
        job.start(schedule[0])
        t_first = wait_first_step(marks, 0, schedule[0], timeout)
        generations.append({"generation": 0, "replicas": schedule[0],
                            "start_to_first_step_s": t_first - t_start})
        for gen, replicas in enumerate(schedule[1:], start=1):
            hold(job, hold_s)
            t_signal = time.time()
            state = job.rescale(replicas, timeout=timeout)
            if state != "running":
                raise RuntimeError("rescale ended in state " + state)
            stopped = [e for e in job.events if e[1] == "stopped"][-1]
            t_gone = stopped[0]
            t_first = wait_first_step(marks, gen, replicas, timeout)
            transitions.append({
                "from": schedule[gen - 1], "to": replicas,
                "signal_to_exit_s": t_gone - t_signal,
                "exit_to_first_step_s": t_first - t_gone,
                "total_s": t_first - t_signal,
                "warm_replicas": [e for e in job.events
                                  if e[1] == "started"][-1][2].get("warm")})
            generations.append({"generation": gen, "replicas": replicas})
        hold(job, hold_s)
        job.signal_stop()
        end = time.time() + timeout
        while job.poll() is None and time.time() < end:
            time.sleep(0.05)
    finally:
        job.kill()
        job.server.close()
    rows = read_marks(marks)
    for g in generations:
        mine = [r for r in rows if r["generation"] == g["generation"]]
        first = [r for r in mine if r["event"] == "first_step"]
        prog = [r for r in mine if r["event"] == "progress"]
        if first and prog:
            # tokens all replicas trained between their first step and their
            # last progress mark
            last = {}
            for r in prog:
                if r["tokens"] >= last.get(r["rank"], (0, 0))[0]:
                    last[r["rank"]] = (r["tokens"], r["t"])
            span = max(t for _, t in last.values()) - \
                min(r["t"] for r in first)
            if span > 0:
                g["tokens_per_s"] = round(
                    sum(tokens for tokens, _ in last.values()) / span, 1)
        impl = [r for r in mine if r["event"] == "process_group"]
        if impl:
            g["impl"] = impl[0].get("impl")
            g["impl_file"] = impl[0].get("file")
        phases = {}
        for name in ("script_start", "torch_imported", "process_group",
                     "model_ready", "first_step"):
            ts = [r["t"] for r in mine if r["event"] == name]
            if ts:
                phases[name] = max(ts)
        order = [k for k in ("script_start", "torch_imported",
                             "process_group", "model_ready", "first_step")
                 if k in phases]
        g["phase_seconds"] = {
            "{}->{}".format(a, b): round(phases[b] - phases[a], 3)
            for a, b in zip(order, order[1:])}
    shutil.rmtree(marks, ignore_errors=True)
    shutil.rmtree(ckpt, ignore_errors=True)
    return {"arm": arm, "schedule": schedule, "transitions": transitions,
            "generations": generations}


def main():
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("--arms", default="reference,own,own-standby")
    parser.add_argument("--schedule", default="2,4,8,4")
    parser.add_argument("--hold", type=float, default=12.0,
                        help="seconds a generation trains before the next "
                             "rescale")
    parser.add_argument("--pool-warmup", type=float, default=25.0,
                        help="seconds the standby pool gets before the job "
                             "starts (a launcher keeps it warm all the time)")
    parser.add_argument("--timeout", type=float, default=180.0)
    parser.add_argument("--device", default="cpu", choices=["cpu", "cuda"],
                        help="cuda: one GPU per replica (NCCL; this "
                             "framework then uses its fused reducer, the "
                             "CUDA-side costs of a restart are included)")
    parser.add_argument("--out")
    args = parser.parse_args()
    schedule = [int(v) for v in args.schedule.split(",")]
    results = []
    for arm in args.arms.split(","):
        if arm == "reference" and not os.path.isdir(
                os.path.join(ROOT, "baseline", "_ref", "adaptdl")):
            print(json.dumps({"arm": arm, "unavailable":
                              "baseline/_ref not installed"}))
            continue
        result = run_arm(arm, schedule, args.hold, args.pool_warmup,
                         args.timeout, args.device)
        results.append(result)
        for t in result["transitions"]:
            print("{:12s} {}->{}: signal->exit {:.2f} s, exit->first step "
                  "{:.2f} s, total {:.2f} s".format(
                      arm, t["from"], t["to"], t["signal_to_exit_s"],
                      t["exit_to_first_step_s"], t["total_s"]), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"config": vars(args), "cpu_count": os.cpu_count(),
                       "results": results}, f, indent=1)


if __name__ == "__main__":
    main()
