"""Small helpers shared across the package."""

import functools
import socket
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


def pick_unused_port(host="127.0.0.1"):
    """Ask the kernel for a free TCP port (replaces ``portpicker``)."""
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind((host, 0))
        return s.getsockname()[1]


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
