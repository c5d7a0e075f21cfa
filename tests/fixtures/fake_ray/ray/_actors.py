"""In-process emulation of the Ray core calls ``adaptdl_b200.ray.aws`` makes:
``@ray.remote`` on an (async) actor class and on plain functions,
``.options(...).remote(...)``, ``ray.get`` / ``put`` / ``cancel`` /
``get_actor``. An actor lives on the asyncio loop that created it (its
methods run there, like Ray's async actors); a remote function call runs in
its own thread (a Ray task runs in a worker process -- here one task at a
time may execute a script, since tasks share this process' environment).
``ray.cancel(ref, force=False)`` delivers SIGINT to the process, which is
what a Ray worker receives."""
import asyncio
import concurrent.futures
import inspect
import os
import signal
import threading

_ACTORS = {}
_LOCAL = threading.local()
CALLS = []                # (kind, name, options) of everything started


class ObjectRef(object):
    """Awaitable on the actor's loop, ``ray.get``-able from task threads."""

    def __init__(self, future, loop=None):
        self.future = future          # concurrent.futures.Future
        self.loop = loop
        self.cancelled = False

    def __await__(self):
        return asyncio.wrap_future(self.future).__await__()


def _resolve(value):
    if isinstance(value, ObjectRef):
        return value.future.result()
    return value


class _Method(object):
    def __init__(self, handle, name):
        self._handle, self._name = handle, name

    def remote(self, *args, **kwargs):
        fn = getattr(self._handle._obj, self._name)
        loop = self._handle._loop
        future = concurrent.futures.Future()
        if loop is None:
            # threaded actor (created outside an event loop; Ray:
            # max_concurrency > 1): every call in its own thread
            def run():
                _LOCAL.ip = WORKER_IP
                try:
                    future.set_result(fn(*args, **kwargs))
                except BaseException as exc:  # noqa: BLE001
                    future.set_exception(exc)
            threading.Thread(target=run, daemon=True,
                             name="ray-actor-" + self._name).start()
            return ObjectRef(future)

        async def call():
            try:
                out = fn(*args, **kwargs)
                if inspect.isawaitable(out):
                    out = await out
                future.set_result(out)
            except BaseException as exc:  # noqa: BLE001
                future.set_exception(exc)
        try:
            running = asyncio.get_running_loop()
        except RuntimeError:
            running = None
        if running is loop:
            asyncio.ensure_future(call())
        else:
            asyncio.run_coroutine_threadsafe(call(), loop)
        return ObjectRef(future, loop)


class ActorHandle(object):
    def __init__(self, obj, loop):
        self._obj, self._loop = obj, loop

    def __getattr__(self, name):
        return _Method(self, name)


class ActorClass(object):
    def __init__(self, cls, options):
        self._cls, self._options = cls, dict(options)

    def options(self, **kwargs):
        return ActorClass(self._cls, dict(self._options, **kwargs))

    def remote(self, *args, **kwargs):
        try:
            loop = asyncio.get_running_loop()
        except RuntimeError:
            loop = None
        is_async = any(inspect.iscoroutinefunction(member) for _, member
                       in inspect.getmembers(self._cls))
        if loop is None and is_async:
            # an async actor created from synchronous code (the driver
            # script): it gets its own event-loop thread, as in Ray
            loop = asyncio.new_event_loop()
            threading.Thread(target=loop.run_forever, daemon=True,
                             name="ray-async-actor").start()
        handle = ActorHandle(self._cls(*args, **kwargs), loop)
        CALLS.append(("actor", self._cls.__name__, dict(self._options)))
        if self._options.get("name"):
            _ACTORS[self._options["name"]] = handle
        return handle


class RemoteFunction(object):
    def __init__(self, fn, options):
        self._fn, self._options = fn, dict(options)

    def options(self, **kwargs):
        return RemoteFunction(self._fn, dict(self._options, **kwargs))

    def remote(self, *args, **kwargs):
        future = concurrent.futures.Future()
        CALLS.append(("task", self._fn.__name__, dict(self._options)))
        ip = node_of(self._options)

        def run():
            _LOCAL.ip = ip
            try:
                future.set_result(self._fn(
                    *[_resolve(a) for a in args],
                    **{k: _resolve(v) for k, v in kwargs.items()}))
            except BaseException as exc:  # noqa: BLE001
                future.set_exception(exc)
        threading.Thread(target=run, daemon=True,
                         name="ray-task-" + self._fn.__name__).start()
        return ObjectRef(future)


def node_of(options):
    """The node a task was pinned to (``resources={"node:<ip>": ...}``),
    else the first worker node."""
    for key in (options.get("resources") or {}):
        if key.startswith("node:"):
            return key[len("node:"):]
    return WORKER_IP


CONTROLLER_IP = "10.0.0.254"
WORKER_IP = "127.0.0.1"


def get_node_ip_address():
    return getattr(_LOCAL, "ip", CONTROLLER_IP)


def remote(*args, **options):
    def wrap(target):
        if inspect.isclass(target):
            return ActorClass(target, options)
        return RemoteFunction(target, options)
    if len(args) == 1 and not options and callable(args[0]):
        return wrap(args[0])
    return wrap


def get(ref):
    if isinstance(ref, (list, tuple)):
        return [get(r) for r in ref]
    return _resolve(ref)


def put(value):
    future = concurrent.futures.Future()
    future.set_result(value)
    return ObjectRef(future)


def cancel(ref, force=False):
    CALLS.append(("cancel", None, {"force": force}))
    if ref.future.done() or ref.cancelled:
        return
    ref.cancelled = True
    if not force:
        os.kill(os.getpid(), signal.SIGINT)    # what a Ray worker receives


def get_actor(name):
    return _ACTORS[name]


def wait(refs, num_returns=1, timeout=None):
    refs = list(refs)
    concurrent.futures.wait([r.future for r in refs], timeout=timeout)
    done = [r for r in refs if r.future.done()]
    return done[:num_returns] if len(done) > num_returns else done, \
        [r for r in refs if not r.future.done()]


def kill(actor, no_restart=True):
    CALLS.append(("kill", type(actor._obj).__name__, {}))
