"""Small helpers shared across the package."""

import functools
import socket
import time
import traceback


def print_exc(fn):
    """Print the traceback of exceptions raised inside autograd hooks and
    engine callbacks (which would otherwise be swallowed or mangled), then
    re-raise (parity: reference ``adaptdl/adaptdl/utils.py:20-31``)."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        try:
            return fn(*args, **kwargs)
        except Exception:
            traceback.print_exc()
            raise
    return wrapper


_RECENT_PORTS = {}          # port -> time handed out (this process)


def pick_unused_port(host="127.0.0.1"):
    """Ask the kernel for a free TCP port (replaces ``portpicker``). A port
    handed out in the last minute is not handed out again by this process:
    launchers that start several jobs back to back pick each job's port
    before the previous job has bound its own."""
    now = time.time()
    for port, when in list(_RECENT_PORTS.items()):
        if now - when > 60.0:
            del _RECENT_PORTS[port]
    for _ in range(64):
        with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
            s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            s.bind((host, 0))
            port = s.getsockname()[1]
        if port not in _RECENT_PORTS:
            break
    _RECENT_PORTS[port] = now
    return port


def parse_version(text):
    """Tiny semver parser: returns (major, minor, patch) or ``None``."""
    if not text:
        return None
    core = str(text).split("+")[0].split("-")[0]
    parts = core.split(".")
    if len(parts) != 3:
        return None
    try:
        return tuple(int(p) for p in parts)
    except ValueError:
        return None
