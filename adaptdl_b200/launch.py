"""Start the replicas that share this container.

    python -m adaptdl_b200.launch train.py --epochs 10
    python -m adaptdl_b200.launch -m package.module --flag
    python -m adaptdl_b200.launch --replicas 8 train.py      # by hand, one box

Under the cluster scheduler a job normally gets one pod per replica. With
``spec.podPerNode: true`` it gets ONE pod per node holding all of the job's
GPUs there, and the controller sets ``ADAPTDL_LOCAL_REPLICAS`` = how many
replicas this pod hosts and ``ADAPTDL_REPLICA_RANK`` = the first of their
ranks. This launcher starts one process per local replica with consecutive
ranks and ``ADAPTDL_LOCAL_RANK`` = 0, 1, ... (the GPU ordinal), which puts
the ranks in one container -- the condition for the peer-memory gradient
reducer (``adaptdl_b200.parallel.hosts_spanned``). Without the variable it
runs the command as a single replica, so the same container command works
in both modes.

Signals: SIGTERM / SIGINT are forwarded to every replica (they checkpoint
and leave with code 143 together). Exit code: 0 if every replica returned 0,
143 if the group was preempted, otherwise the first failing replica's code
(the others are terminated as soon as one fails).
"""

import os
import signal
import subprocess
import sys
import time

EXIT_PREEMPTED = 143


def replica_environments(environ=None):
    """One environment per local replica."""
    environ = dict(os.environ if environ is None else environ)
    count = int(environ.get("ADAPTDL_LOCAL_REPLICAS") or 1)
    first = int(environ.get("ADAPTDL_REPLICA_RANK") or 0)
    if "ADAPTDL_NUM_REPLICAS" not in environ:
        # started by hand on one box (no scheduler): these replicas are the
        # whole job
        from adaptdl_b200.utils import pick_unused_port
        environ.update(ADAPTDL_NUM_REPLICAS=str(count), ADAPTDL_NUM_NODES="1",
                       ADAPTDL_MASTER_ADDR="127.0.0.1")
        environ.setdefault("ADAPTDL_MASTER_PORT", str(pick_unused_port()))
    out = []
    for local in range(count):
        env = dict(environ)
        env["ADAPTDL_REPLICA_RANK"] = str(first + local)
        env["ADAPTDL_LOCAL_RANK"] = str(local)
        env["ADAPTDL_LOCAL_REPLICAS"] = str(count)
        out.append(env)
    return out


def exit_code(codes):
    """Code of the group from the replicas' codes (see module docstring)."""
    # killed by SIGTERM before the handler was installed = preempted early
    codes = [EXIT_PREEMPTED if c == -signal.SIGTERM else c for c in codes]
    failures = [c for c in codes if c not in (0, EXIT_PREEMPTED)]
    if failures:
        return failures[0] if failures[0] > 0 else 128 - failures[0]
    if any(c == EXIT_PREEMPTED for c in codes):
        return EXIT_PREEMPTED
    return 0


def run(command, environ=None, poll=0.2):
    procs = [subprocess.Popen(command, env=env)
             for env in replica_environments(environ)]

    def forward(signum, frame):
        for proc in procs:
            if proc.poll() is None:
                proc.send_signal(signum)
    previous = {sig: signal.signal(sig, forward)
                for sig in (signal.SIGTERM, signal.SIGINT)}
    try:
        while True:
            codes = [proc.poll() for proc in procs]
            failed = [c for c in codes if c is not None and
                      c not in (0, EXIT_PREEMPTED, -signal.SIGTERM)]
            if failed or all(c is not None for c in codes):
                break
            time.sleep(poll)
        if failed:                       # one replica died: stop the rest
            for proc in procs:
                if proc.poll() is None:
                    proc.terminate()
            deadline = time.time() + 30
            for proc in procs:
                try:
                    proc.wait(max(deadline - time.time(), 0.1))
                except subprocess.TimeoutExpired:
                    proc.kill()
        codes = [proc.wait() for proc in procs]
        # the replicas we terminated report the signal, not a verdict: the
        # group's code is that of the replica that failed on its own
        return exit_code(failed or codes)
    finally:
        for sig, handler in previous.items():
            signal.signal(sig, handler)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0 if argv else 2
    environ = None
    if argv[0] == "--replicas":          # by hand: N replicas on this box
        environ = dict(os.environ, ADAPTDL_LOCAL_REPLICAS=argv[1])
        argv = argv[2:]
    if argv and argv[0] == "--":
        argv = argv[1:]
    sys.exit(run([sys.executable] + argv, environ))


if __name__ == "__main__":
    main()
