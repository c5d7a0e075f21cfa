"""Warm standby replica: an interpreter that has already paid for its imports.

A rescale is "checkpoint, exit, start again at the new size". On a B200 box
the new generation's first optimizer step is gated by things that have
nothing to do with the job: ``import torch`` (~3 s), the first touch of
``torch.optim`` (which drags in ``torch._dynamo`` / ``sympy``, another ~3.4 s)
and this package's own imports -- together about 6.5 s of a 7.5-8 s restart
(``profiles/r2_elastic/README.md``). The single-box launcher
(:mod:`adaptdl_b200.sched.local`) therefore keeps a pool of these processes:

    python -m adaptdl_b200.sched.standby [--preload mod,mod] script.py args...

does the imports, then blocks on stdin. One JSON line ``{"env": {...}}``
releases it: the generation-specific variables (replica count, rank, restart
count, rendezvous port) are applied and ``script.py`` runs as ``__main__``,
exactly as if it had been started directly. EOF on stdin means "not needed":
exit 0. A standby never touches CUDA, so it holds no device memory and no
context while the current generation is training.

The reference has no counterpart (a Kubernetes pod always cold-starts); the
contract of the released process -- environment, argv, exit code 143 on
preemption -- is the one :func:`adaptdl_b200.sched.controller.build_pod`
gives a pod.
"""

import importlib
import json
import os
import runpy
import sys

# imported before the wait: the heavy, job-independent part of a start
PRELOAD = ("torch", "torch.optim", "torch.nn", "torch.distributed",
           "torch.utils.data", "adaptdl_b200.torch")


def _preload(extra):
    from adaptdl_b200.utils import rescale_trace
    rescale_trace.suspend()        # the life cycle starts at release
    loaded = []
    for name in PRELOAD + tuple(extra):
        try:
            importlib.import_module(name)
            loaded.append(name)
        except Exception:  # noqa: BLE001 - the script will report it itself
            pass
    try:
        _warm_torch()
    except Exception:  # noqa: BLE001
        pass
    try:
        # map the native library now (dlopen only: no CUDA call is made)
        from adaptdl_b200 import _native
        _native.load(build_if_needed=False)
    except Exception:  # noqa: BLE001
        pass
    return loaded


def _warm_torch():
    """Torch defers a good part of its start-up to first use: constructing
    the first ``Optimizer`` imports ``torch._dynamo`` and ``sympy`` (3.5 s
    measured here), the first backward spins up the autograd engine. One
    throw-away CPU training step pays for all of it."""
    import torch
    layer = torch.nn.Linear(4, 4)
    for opt_class in (torch.optim.SGD, torch.optim.Adam):
        opt = opt_class(layer.parameters(), lr=0.1)
        sched = torch.optim.lr_scheduler.StepLR(opt, 1)
        opt.zero_grad()
        torch.nn.functional.mse_loss(layer(torch.zeros(2, 4)),
                                     torch.zeros(2, 4)).backward()
        torch.nn.utils.clip_grad_norm_(layer.parameters(), 1.0)
        opt.step()
        sched.step()


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    extra = ()
    if argv and argv[0] == "--preload":
        extra = tuple(m for m in argv[1].split(",") if m)
        argv = argv[2:]
    if argv and argv[0] == "--":
        argv = argv[1:]
    if not argv:
        sys.stderr.write("usage: python -m adaptdl_b200.sched.standby "
                         "[--preload mod,...] script.py [args...]\n")
        return 2
    script = argv[0]
    _preload(extra)

    line = sys.stdin.buffer.readline()
    if not line.strip():
        return 0                   # pool drained: never needed
    order = json.loads(line)
    os.environ.update({k: str(v) for k, v in order.get("env", {}).items()})
    for name in order.get("unset", ()):
        os.environ.pop(name, None)

    from adaptdl_b200 import _signal
    from adaptdl_b200.utils import rescale_trace
    _signal.set_exit_flag(False)   # anything received while idle is void
    rescale_trace.resume()
    rescale_trace.mark("interpreter_up", standby=True)

    sys.argv = [script] + argv[1:]
    # like `python script.py`: the script's directory leads sys.path
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
